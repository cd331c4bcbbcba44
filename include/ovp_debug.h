/* TEST / TUNING HOOKS of the B200-native ov_plane hot path — NOT part of the drop-in ABI (include/ovp.h).
 *
 * These entry points exist only in ov_plane_b200/lib/libovp_debug.so (the product sources compiled with -DOVP_DEBUG); the
 * product library libovp.so does not export them.  They let tools/microbench*.py measure single kernels and let
 * tests/test_gpu_cholfused.py unit-test chol_fused_kernel against NumPy on arbitrary matrices. */
#ifndef OVP_DEBUG_H
#define OVP_DEBUG_H
#include "ovp.h"
#ifdef __cplusplus
extern "C" {
#endif

/* dependent-chain latencies (cycles per operation, one warp) of DFMA, rsqrt, 1/x, sqrt, shuffles, shared-memory loads, DMMA */
int ovp_debug_fp64_latency(ovp_ctx *ctx, double *out8);
/* fused Cholesky on a synthetic SPD n x n system (+ mrows x n right-hand side): out[0] = us per (fill + factor), out[1] = us per
 * fill, out[2..] = per-CTA globaltimer stamps of the last run */
int ovp_debug_chol_fused(ovp_ctx *ctx, int n, int mrows, int iters, double *out, int out_cap);
/* factor the lower triangle of a host matrix A (n x n, column-major) over its leading npiv columns with pivot tolerance tol and,
 * when M is given, solve Y = M L^-T (mrows x npiv) and w = L^-1 z */
int ovp_debug_chol_solve(ovp_ctx *ctx, const double *A, int n, int npiv, double tol, const double *M, int mrows, const double *z,
                         double *L_out, double *Y_out, double *w_out);
/* variants of the 16-column in-warp pivot chain (tools/microbench_potrf.py) */
int ovp_debug_potrf_variants(ovp_ctx *ctx, int variant, int reps, double *out16);
int ovp_debug_potrf_cond(ovp_ctx *ctx, int nthreads, int nchain, int smem_bytes, int reps, double *out16, int mode);

#ifdef __cplusplus
}
#endif
#endif
