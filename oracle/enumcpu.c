/*
 * enumcpu.c — CPU ORACLE for the extreme-point enumeration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (simplexmethod_b200/,
 * libenumgpu) links, loads or calls this file; only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may use it, and there only as the checker / the CPU arm.
 *
 * PARITY STATUS: pinned to the reference's own code paths; the bits of real
 * Eigen stay unpinned.  The reference's EnumerationSolver is an empty stub
 * (src/EnumerationSolver.h:3-10) and its per-basis primitives need Eigen 3.4.0
 * (CMakeLists.txt:12-17, FetchContent, not vendored, not on this box).  What
 * pins this file (tests/test_oracle_*.py, tests/test_reference_code.py):
 *   - oracle/_ref: the reference's own Canonical / Solver sources compiled from
 *     where they lie against an Eigen API stand-in (oracle/eigen_shim), and the
 *     enumeration composed from them (ref_driver.cpp): class of every basis,
 *     counters, optimum on the tiny LPs, dense (4,12)...(8,24) and windows of
 *     the headline LP (x and z to 1e-9: QR there, GE here);
 *   - exact-rational goldens for the reference's own fixtures
 *     (input_symmetric.txt, src/main.cpp:48-57, tests/test_canonical.cpp:12-22
 *     incl. its EXPECT_DOUBLE_EQ pin :52-57) and Beale's LP, tests/golden/;
 *   - an exact fractions.Fraction enumerator on random small LPs;
 *   - SciPy HiGHS optimum on the dense m=8,n=24 LPs.
 *
 * What it restates (reference file:line):
 *   gather   B.col(i) = A.col(basis[i])          Canonical.cpp:183-187,
 *                                                SimplexSolover.h:110-115
 *   solve    B x_B = b, singular bases rejected  SimplexSolover.h:124-126
 *            (the reference uses Eigen FullPivLU / ColPivHouseholderQR,
 *            Canonical.cpp:189; here: partial-pivot GE as the north star
 *            mandates, arithmetic frozen below)
 *   feasible all x_i >= -1e-9                    Canonical.cpp:165-177
 *   value    c . x restricted to basic terms     Canonical.cpp:79-87
 *   sense    IsMaximization = !minimize          Canonical.cpp:141-144
 *   order    lexicographic subsets, strict '<' keeps the lowest rank
 *
 * FROZEN PER-BASIS ARITHMETIC.  Every fma() is one IEEE-754
 * fused multiply-add, every other operation is a single correctly rounded
 * IEEE double operation; compile with -ffp-contract=off so the compiler adds
 * none of its own.  thr = eps_piv * max|A_ij| (one multiply).
 *
 *   M[r][j] = A[r + S[j]*lda] (j<m);  M[r][m] = b[r]
 *   for k = 0..m-1:
 *       p = first r >= k maximising |M[r][k]|
 *       if !(|M[p][k]| > thr)  -> SINGULAR
 *       swap rows k,p
 *       rinv[k] = 1.0 / M[k][k]
 *       for r = k+1..m-1:  l = M[r][k]*rinv[k]
 *           for j = k+1..m:  M[r][j] = fma(-l, M[k][j], M[r][j])
 *   t[i] = M[i][m]
 *   for j = m-1..0:   x[j] = t[j]*rinv[j]                 (column sweep)
 *       for i = 0..j-1:  t[i] = fma(-M[i][j], x[j], t[i])
 *   if any !(x[j] >= -eps_feas) -> INFEASIBLE            (NaN is infeasible)
 *   z = 0;  for j = m-1..0:  z = fma(c[S[j]], x[j], z)
 *   key = maximize ? -z : z
 *   best is replaced iff key < best_key, or key == best_key and rank < best_rank
 *
 * ENUMGPU_PIVOT_RELATIVE (options.pivot_rule; the Eigen-like rule, cf.
 * FullPivLU::isInvertible at SimplexSolover.h:124-126): in the loop above the
 * test becomes  if !(|M[p][k]| > 0) -> SINGULAR  and the elimination runs to the
 * end, tracking pmax = max_k |pivot_k|, pmin = min_k |pivot_k|; afterwards
 *   if !(pmin > eps_rel * pmax) -> SINGULAR      (eps_rel = eps_piv, default m*2^-52)
 */
#define _GNU_SOURCE
#include "enumcpu.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---------------------------------------------------------------- binomials */

static uint64_t g_binom[ENUMGPU_MAX_N + 1][ENUMGPU_MAX_N + 1];
static pthread_once_t g_binom_once = PTHREAD_ONCE_INIT;

static void binom_init(void)
{
    /* Pascal's triangle; entries that would pass 2^63 saturate to 0 = "overflow" */
    for (int n = 0; n <= ENUMGPU_MAX_N; ++n) {
        for (int k = 0; k <= ENUMGPU_MAX_N; ++k) g_binom[n][k] = 0;
        g_binom[n][0] = 1;
        for (int k = 1; k <= n; ++k) {
            uint64_t a = g_binom[n - 1][k - 1], b2 = (k <= n - 1) ? g_binom[n - 1][k] : 0;
            int a_bad = (a == 0), b_bad = (k <= n - 1) && (b2 == 0);
            if (a_bad || b_bad || a + b2 < a || (a + b2) >> 63) g_binom[n][k] = 0;
            else g_binom[n][k] = a + b2;
        }
    }
}

uint64_t enumcpu_binomial(int32_t n, int32_t k)
{
    pthread_once(&g_binom_once, binom_init);
    if (n < 0 || k < 0 || k > n || n > ENUMGPU_MAX_N) return 0;
    return g_binom[n][k];
}

/* rank(S) = C(n,m) - 1 - sum_i C(n-1-S[i], m-i)    */
uint64_t enumcpu_rank(int32_t n, int32_t m, const int32_t* S)
{
    pthread_once(&g_binom_once, binom_init);
    if (m < 1 || n < m || n > ENUMGPU_MAX_N) return UINT64_MAX;
    uint64_t acc = 0;
    for (int i = 0; i < m; ++i) {
        if (S[i] < 0 || S[i] >= n || (i && S[i] <= S[i - 1])) return UINT64_MAX;
        int top = n - 1 - S[i], kk = m - i;
        if (top >= kk) acc += g_binom[top][kk];
    }
    return g_binom[n][m] - 1 - acc;
}

int enumcpu_unrank(int32_t n, int32_t m, uint64_t r, int32_t* S)
{
    pthread_once(&g_binom_once, binom_init);
    if (m < 1 || n < m || n > ENUMGPU_MAX_N || g_binom[n][m] == 0 || r >= g_binom[n][m]) return -1;
    int v = 0;
    for (int i = 0; i < m; ++i) {
        for (;;) {
            uint64_t cnt = g_binom[n - 1 - v][m - 1 - i]; /* subsets whose i-th element is v */
            if (cnt <= r) { r -= cnt; ++v; } else break;
        }
        S[i] = v++;
    }
    return 0;
}

/* successor in lexicographic order; returns 0 when S was the last subset */
static int next_subset(int n, int m, int32_t* S)
{
    int i = m - 1;
    while (i >= 0 && S[i] == n - m + i) --i;
    if (i < 0) return 0;
    ++S[i];
    for (int j = i + 1; j < m; ++j) S[j] = S[j - 1] + 1;
    return 1;
}

/* ------------------------------------------------------- one basis (frozen) */

int enumcpu_eval_basis(const enumgpu_problem* p, double eps_feas, double thr,
                       const int32_t* S, double* x, double* z_out)
{
    return enumcpu_eval_basis_rule(p, eps_feas, ENUMGPU_PIVOT_ABSOLUTE, thr, S, x, z_out);
}

/* rule ABSOLUTE: tol = thr = eps_piv * max|A|;  rule RELATIVE: tol = eps_rel */
int enumcpu_eval_basis_rule(const enumgpu_problem* p, double eps_feas, int rule, double tol,
                            const int32_t* S, double* x, double* z_out)
{
    const double thr = (rule == ENUMGPU_PIVOT_RELATIVE) ? 0.0 : tol;
    double pmax = 0.0, pmin = INFINITY;
    const int m = p->m, lda = p->lda;
    double M[ENUMGPU_MAX_M][ENUMGPU_MAX_M + 1];
    double rinv[ENUMGPU_MAX_M], t[ENUMGPU_MAX_M];

    for (int j = 0; j < m; ++j) {
        const double* col = p->A_colmajor + (size_t)S[j] * lda;
        for (int r = 0; r < m; ++r) M[r][j] = col[r];
    }
    for (int r = 0; r < m; ++r) M[r][m] = p->b[r];

    for (int k = 0; k < m; ++k) {
        int piv = k;
        double best = fabs(M[k][k]);
        for (int r = k + 1; r < m; ++r) {
            double v = fabs(M[r][k]);
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > thr)) return ENUMCPU_SINGULAR;
        if (best > pmax) pmax = best;
        if (best < pmin) pmin = best;
        if (piv != k)
            for (int j = k; j <= m; ++j) { double s = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = s; }
        rinv[k] = 1.0 / M[k][k];
        for (int r = k + 1; r < m; ++r) {
            double l = M[r][k] * rinv[k];
            for (int j = k + 1; j <= m; ++j) M[r][j] = fma(-l, M[k][j], M[r][j]);
        }
    }
    if (rule == ENUMGPU_PIVOT_RELATIVE && !(pmin > tol * pmax)) return ENUMCPU_SINGULAR;
    for (int i = 0; i < m; ++i) t[i] = M[i][m];
    for (int j = m - 1; j >= 0; --j) {
        x[j] = t[j] * rinv[j];
        for (int i = 0; i < j; ++i) t[i] = fma(-M[i][j], x[j], t[i]);
    }
    int feasible = 1;
    for (int j = 0; j < m; ++j)
        if (!(x[j] >= -eps_feas)) feasible = 0;
    double z = 0.0;
    for (int j = m - 1; j >= 0; --j) z = fma(p->c[S[j]], x[j], z);
    *z_out = z;
    return feasible ? ENUMCPU_FEASIBLE : ENUMCPU_INFEASIBLE;
}

double enumcpu_scale(const enumgpu_problem* p)
{
    double s = 0.0;
    for (int j = 0; j < p->n; ++j)
        for (int i = 0; i < p->m; ++i) {
            double v = fabs(p->A_colmajor[i + (size_t)j * p->lda]);
            if (v > s) s = v;
        }
    return s;
}

/* --------------------------------------------------------------- range scan */

typedef struct {
    const enumgpu_problem* p;
    double eps_feas, thr;      /* thr: absolute threshold, or eps_rel under the relative rule */
    int rule;
    uint64_t begin, end;
    /* optional listing of the ranks of one class (shared by all threads) */
    int list_cls;              /* -1: none */
    uint64_t* list_out;
    uint64_t list_cap;
    uint64_t* list_count;      /* atomic */
    /* out */
    double best_key;
    uint64_t best_rank, n_sing, n_infeas, n_feas;
    uint8_t* status_out; /* optional, indexed by rank - status_base */
    uint64_t status_base;
} scan_job;

static void* scan_range(void* arg)
{
    scan_job* J = (scan_job*)arg;
    const enumgpu_problem* p = J->p;
    int32_t S[ENUMGPU_MAX_M];
    double x[ENUMGPU_MAX_M], z;
    J->best_key = INFINITY; J->best_rank = UINT64_MAX;
    J->n_sing = J->n_infeas = J->n_feas = 0;
    if (J->begin >= J->end) return NULL;
    enumcpu_unrank(p->n, p->m, J->begin, S);
    for (uint64_t r = J->begin; r < J->end; ++r) {
        int st = enumcpu_eval_basis_rule(p, J->eps_feas, J->rule, J->thr, S, x, &z);
        if (J->status_out) J->status_out[r - J->status_base] = (uint8_t)st;
        if (st == J->list_cls) {
            uint64_t pos = __atomic_fetch_add(J->list_count, 1, __ATOMIC_RELAXED);
            if (pos < J->list_cap) J->list_out[pos] = r;
        }
        if (st == ENUMCPU_SINGULAR) ++J->n_sing;
        else if (st == ENUMCPU_INFEASIBLE) ++J->n_infeas;
        else {
            ++J->n_feas;
            double key = p->maximize ? -z : z;
            if (key < J->best_key || (key == J->best_key && r < J->best_rank)) {
                J->best_key = key; J->best_rank = r;
            }
        }
        next_subset(p->n, p->m, S);
    }
    return NULL;
}

static int check_problem(const enumgpu_problem* p)
{
    if (!p || !p->A_colmajor || !p->b || !p->c) return ENUMGPU_ERR_ARG;
    if (p->m < 1 || p->m > ENUMGPU_MAX_M || p->n < p->m || p->n > ENUMGPU_MAX_N || p->lda < p->m)
        return ENUMGPU_ERR_ARG;
    for (int j = 0; j < p->n; ++j) {
        if (!isfinite(p->c[j])) return ENUMGPU_ERR_NONFINITE;
        for (int i = 0; i < p->m; ++i)
            if (!isfinite(p->A_colmajor[i + (size_t)j * p->lda])) return ENUMGPU_ERR_NONFINITE;
    }
    for (int i = 0; i < p->m; ++i)
        if (!isfinite(p->b[i])) return ENUMGPU_ERR_NONFINITE;
    return 0;
}

static int cmp_u64(const void* a, const void* b)
{
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}

int enumcpu_solve_ex(const enumgpu_problem* p, const enumgpu_options* o, int n_threads,
                     uint8_t* status_out, enumgpu_result* out)
{
    return enumcpu_solve_list(p, o, n_threads, status_out, -1, NULL, 0, NULL, out);
}

int enumcpu_solve_list(const enumgpu_problem* p, const enumgpu_options* o, int n_threads, uint8_t* status_out,
                       int list_cls, uint64_t* list_out, uint64_t list_cap, uint64_t* n_listed, enumgpu_result* out)
{
    memset(out, 0, sizeof *out);
    int rc = check_problem(p);
    if (rc) { out->status = rc; return rc; }
    uint64_t total = enumcpu_binomial(p->n, p->m);
    if (total == 0) { out->status = ENUMGPU_ERR_RANGE; return out->status; }
    const int rule = o ? o->pivot_rule : ENUMGPU_PIVOT_ABSOLUTE;
    if (rule != ENUMGPU_PIVOT_ABSOLUTE && rule != ENUMGPU_PIVOT_RELATIVE) { out->status = ENUMGPU_ERR_ARG; return out->status; }
    double eps_feas = (o && o->eps_feas >= 0) ? o->eps_feas : 1e-9;
    double eps_piv  = (o && o->eps_piv  >= 0) ? o->eps_piv
                    : (rule == ENUMGPU_PIVOT_RELATIVE ? (double)p->m * 0x1p-52 : 1e-9);
    uint64_t begin = o ? o->rank_begin : 0, end = o ? o->rank_end : 0;
    if (begin == 0 && end == 0) end = total;
    if (begin > end || end > total) { out->status = ENUMGPU_ERR_RANGE; return out->status; }
    double thr = (rule == ENUMGPU_PIVOT_RELATIVE) ? eps_piv : eps_piv * enumcpu_scale(p);
    uint64_t listed = 0;

    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    uint64_t span = end - begin;
    if ((uint64_t)n_threads > span) n_threads = span ? (int)span : 1;

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    scan_job* jobs = (scan_job*)calloc((size_t)n_threads, sizeof(scan_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) {
        jobs[i].p = p; jobs[i].eps_feas = eps_feas; jobs[i].thr = thr; jobs[i].rule = rule;
        jobs[i].list_cls = list_out || n_listed ? list_cls : -1; jobs[i].list_out = list_out; jobs[i].list_cap = list_out ? list_cap : 0;
        jobs[i].list_count = &listed;
        jobs[i].begin = begin + span / (uint64_t)n_threads * (uint64_t)i
                      + ((uint64_t)i < span % (uint64_t)n_threads ? (uint64_t)i : span % (uint64_t)n_threads);
        jobs[i].status_out = status_out; jobs[i].status_base = begin;
    }
    for (int i = 0; i < n_threads; ++i) jobs[i].end = (i + 1 < n_threads) ? jobs[i + 1].begin : end;
    for (int i = 1; i < n_threads; ++i) pthread_create(&th[i], NULL, scan_range, &jobs[i]);
    scan_range(&jobs[0]);
    for (int i = 1; i < n_threads; ++i) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);

    double best_key = INFINITY; uint64_t best_rank = UINT64_MAX;
    for (int i = 0; i < n_threads; ++i) {
        out->n_singular += jobs[i].n_sing; out->n_infeasible += jobs[i].n_infeas;
        out->n_feasible += jobs[i].n_feas;
        if (jobs[i].best_key < best_key || (jobs[i].best_key == best_key && jobs[i].best_rank < best_rank)) {
            best_key = jobs[i].best_key; best_rank = jobs[i].best_rank;
        }
    }
    free(jobs); free(th);
    if (n_listed) *n_listed = listed;
    if (list_out) qsort(list_out, (size_t)(listed < list_cap ? listed : list_cap), sizeof(uint64_t), cmp_u64);

    out->m = p->m;
    out->n_bases = span;
    out->key = best_key;
    out->best_rank = best_rank;
    out->kernel_ms = (double)(t1.tv_sec - t0.tv_sec) * 1e3 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-6;
    out->algo_used = ENUMGPU_ALGO_INDEPENDENT;
    if (best_rank == UINT64_MAX) { out->status = ENUMGPU_NO_FEASIBLE; out->objective = NAN; return out->status; }
    enumcpu_unrank(p->n, p->m, best_rank, out->basis);
    double z;
    enumcpu_eval_basis_rule(p, eps_feas, rule, thr, out->basis, out->x_B, &z);
    out->objective = z;
    out->status = ENUMGPU_OK;
    return out->status;
}

int enumcpu_solve(const enumgpu_problem* p, const enumgpu_options* o, enumgpu_result* out)
{
    return enumcpu_solve_ex(p, o, 1, NULL, out);
}
