/*
 * enumcpu.h — interface of the CPU oracle (TEST INFRASTRUCTURE ONLY; see the
 * header of enumcpu.c).  Same problem/options/result structs as the product
 * ABI (include/enumgpu.h) so tests can diff results field by field.
 */
#ifndef ENUMCPU_H_
#define ENUMCPU_H_

#include "../include/enumgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ENUMCPU_FEASIBLE   0
#define ENUMCPU_INFEASIBLE 1
#define ENUMCPU_SINGULAR   2

uint64_t enumcpu_binomial(int32_t n, int32_t k);
uint64_t enumcpu_rank(int32_t n, int32_t m, const int32_t* subset);
int      enumcpu_unrank(int32_t n, int32_t m, uint64_t rank, int32_t* subset);

/* max |A_ij| — the pivot-threshold scale */
double   enumcpu_scale(const enumgpu_problem* p);

/* frozen arithmetic for ONE basis S (sorted); x has m entries; thr = eps_piv*scale.
 * returns ENUMCPU_FEASIBLE / INFEASIBLE / SINGULAR (x, z undefined if singular) */
int      enumcpu_eval_basis(const enumgpu_problem* p, double eps_feas, double thr,
                            const int32_t* S, double* x, double* z);

/* same with the singularity rule of enumgpu_options.pivot_rule: tol = eps_piv*scale (ENUMGPU_PIVOT_ABSOLUTE)
 * or the relative threshold eps_rel (ENUMGPU_PIVOT_RELATIVE: min|pivot| > eps_rel * max|pivot|) */
int      enumcpu_eval_basis_rule(const enumgpu_problem* p, double eps_feas, int rule, double tol,
                                 const int32_t* S, double* x, double* z);

/* single-threaded enumeration of the rank range */
int      enumcpu_solve(const enumgpu_problem* p, const enumgpu_options* o, enumgpu_result* out);

/* n_threads pthreads over contiguous sub-ranges; status_out (may be NULL) gets
 * one ENUMCPU_* byte per rank of the range; out->kernel_ms = wall time */
int      enumcpu_solve_ex(const enumgpu_problem* p, const enumgpu_options* o, int n_threads,
                          uint8_t* status_out, enumgpu_result* out);

/* enumcpu_solve_ex that also lists the ranks (ascending) of the bases of class list_cls (ENUMCPU_*; -1 = none):
 * list_out (may be NULL) gets at most list_cap of them, *n_listed (may be NULL) their full number */
int      enumcpu_solve_list(const enumgpu_problem* p, const enumgpu_options* o, int n_threads, uint8_t* status_out,
                            int list_cls, uint64_t* list_out, uint64_t list_cap, uint64_t* n_listed,
                            enumgpu_result* out);

#ifdef __cplusplus
}
#endif
#endif
