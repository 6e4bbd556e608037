// enum_common.cuh — shared device-side definitions of libenumgpu (sm_100a).
//
// The per-basis arithmetic implemented by every kernel here is the frozen
// definition of DESIGN.md §3 (restated in oracle/enumcpu.c for the checker):
// partial-pivot GE with one reciprocal per pivot, explicit FMAs, column-sweep
// back substitution, feasibility x >= -eps, objective accumulated from the
// last basic variable to the first.  Reference semantics: gather
// Canonical.cpp:183-187, singular rejected SimplexSolover.h:124-126, feasible
// Canonical.cpp:165-177, objective Canonical.cpp:79-87.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "../../include/enumgpu.h"

namespace enumgpu {

constexpr int kMaxM = ENUMGPU_MAX_M;
constexpr int kMaxN = ENUMGPU_MAX_N;
constexpr int kBinomRows = kMaxN + 1;   // top index 0..64
constexpr int kBinomCols = kMaxM + 1;   // k index 0..16

// Everything a kernel needs besides the matrix data.  Passed by value.
struct LaunchParams {
    const double* A;        // device, column-major
    const double* b;
    const double* c;
    const uint64_t* binom;  // device, [kBinomRows][kBinomCols], C(top,k)
    int32_t  m, n, lda, maximize;
    double   eps_feas;
    double   thr;           // eps_piv * max|A_ij|
    uint64_t rank_begin, rank_end;
    uint32_t chunk;         // ranks per thread (independent kernel)
    uint32_t shard_index;   // interleaved sharding: this launch owns the work
    uint32_t shard_count;   // windows whose index is shard_index mod shard_count
    // optional listing of the feasible bases (enumgpu_list_feasible): ranks are appended in
    // arbitrary order; *list_count counts all of them, only the first list_cap are stored
    unsigned long long* list_count;
    uint64_t* list_ranks;
    uint64_t  list_cap;
};

__device__ __forceinline__ void list_append(unsigned long long* count, uint64_t* ranks, uint64_t cap, uint64_t rank)
{
    const unsigned long long pos = atomicAdd(count, 1ull);
    if (pos < cap) ranks[pos] = rank;
}

// (key, rank) pair + counters: what every thread / warp / block / GPU reduces.
struct Best {
    double   key;
    uint64_t rank;
};

__device__ __forceinline__ bool better(double k1, uint64_t r1, double k2, uint64_t r2)
{
    return (k1 < k2) || (k1 == k2 && r1 < r2);
}

struct BlockPartial {          // one per block, reduced by the finalize kernel
    double   key;
    uint64_t rank;
    uint64_t n_sing, n_infeas, n_feas;
};

// Compile-time loop: f(integral_constant<int,I>) for I in [B, E).  '#pragma
// unroll' is only a request — nvcc leaves inner loops of large unrolled nests
// rolled, which turns register arrays into local memory.  This cannot.
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}
// same, descending: I = E-1 ... B
template <int B, int E, class F>
__device__ __forceinline__ void static_rfor(F&& f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, E - 1>{});
        static_rfor<B, E - 1>(f);
    }
}

// IEEE double ops that must not be contracted or reassociated.
__device__ __forceinline__ double fnma(double a, double b, double c) { return __fma_rn(-a, b, c); }

// ---------------------------------------------------------------------------
// Warp + block reduction of (key, rank) by lexicographic min and of counters
// by sum.  Result valid in thread 0 of the block.
template <int kThreads, class Count>     // Count: uint32_t or uint64_t per-thread counters
__device__ __forceinline__ void block_reduce(double& key, uint64_t& rank,
                                             Count& ns, Count& ni, Count& nf,
                                             BlockPartial* out_slot, int n_warps = kThreads / 32)
{
    constexpr int kWarps = kThreads / 32;   // capacity; n_warps <= kWarps are live
    __shared__ double   s_key[kWarps];
    __shared__ uint64_t s_rank[kWarps];
    __shared__ uint64_t s_cnt[kWarps][3];
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double   k2 = __shfl_down_sync(full, key, off);
        uint64_t r2 = __shfl_down_sync(full, rank, off);
        if (better(k2, r2, key, rank)) { key = k2; rank = r2; }
        ns += __shfl_down_sync(full, ns, off);
        ni += __shfl_down_sync(full, ni, off);
        nf += __shfl_down_sync(full, nf, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_key[warp] = key; s_rank[warp] = rank; s_cnt[warp][0] = ns; s_cnt[warp][1] = ni; s_cnt[warp][2] = nf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bk = s_key[0]; uint64_t br = s_rank[0];
        uint64_t cs = s_cnt[0][0], ci = s_cnt[0][1], cf = s_cnt[0][2];
        for (int w = 1; w < n_warps; ++w) {
            if (better(s_key[w], s_rank[w], bk, br)) { bk = s_key[w]; br = s_rank[w]; }
            cs += s_cnt[w][0]; ci += s_cnt[w][1]; cf += s_cnt[w][2];
        }
        out_slot->key = bk; out_slot->rank = br;
        out_slot->n_sing = cs; out_slot->n_infeas = ci; out_slot->n_feas = cf;
    }
}

// Unrank r into S[0..m) using a binomial table in (shared) memory.
__device__ __forceinline__ void unrank_lex(const uint64_t* __restrict__ binom, int n, int m,
                                           uint64_t r, int* S)
{
    int v = 0;
    for (int i = 0; i < m; ++i) {
        for (;;) {
            uint64_t cnt = binom[(n - 1 - v) * kBinomCols + (m - 1 - i)];
            if (cnt <= r) { r -= cnt; ++v; } else break;
        }
        S[i] = v++;
    }
}

// One basis with run-time m (local-memory arrays): used once per enqueue, by
// one thread, to materialise x_B / objective of the winning basis on the device.
__device__ inline int eval_basis_generic(const double* A, int lda, const double* b, const double* c, int m,
                                  const int* S, double thr, double eps_feas, double* x, double* z_out)
{
    double Mx[kMaxM][kMaxM + 1];
    double rinv[kMaxM];
    for (int j = 0; j < m; ++j)
        for (int r = 0; r < m; ++r) Mx[r][j] = A[r + (size_t)S[j] * lda];
    for (int r = 0; r < m; ++r) Mx[r][m] = b[r];
    for (int k = 0; k < m; ++k) {
        int p = k;
        double best = fabs(Mx[k][k]);
        for (int r = k + 1; r < m; ++r) {
            const double v = fabs(Mx[r][k]);
            if (v > best) { best = v; p = r; }
        }
        if (!(best > thr)) return 2;
        if (p != k)
            for (int j = k; j <= m; ++j) { const double t = Mx[k][j]; Mx[k][j] = Mx[p][j]; Mx[p][j] = t; }
        rinv[k] = __drcp_rn(Mx[k][k]);
        for (int r = k + 1; r < m; ++r) {
            const double l = __dmul_rn(Mx[r][k], rinv[k]);
            for (int j = k + 1; j <= m; ++j) Mx[r][j] = fnma(l, Mx[k][j], Mx[r][j]);
        }
    }
    bool infeasible = false;
    double z = 0.0;
    for (int j = m - 1; j >= 0; --j) {
        x[j] = __dmul_rn(Mx[j][m], rinv[j]);
        infeasible |= !(x[j] >= -eps_feas);
        for (int i = 0; i < j; ++i) Mx[i][m] = fnma(Mx[i][j], x[j], Mx[i][m]);
        z = __fma_rn(c[S[j]], x[j], z);
    }
    *z_out = z;
    return infeasible ? 1 : 0;
}

}  // namespace enumgpu
