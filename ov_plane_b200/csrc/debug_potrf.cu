// Micro-benchmarks of the serial pivot chain of the 16-column in-warp Cholesky panel (tools/microbench_potrf.py).  Debug only.
#include "ovp_internal.h"
namespace ovp {
#define DP_LD 68
// variant: 0 smem broadcast (production form), 1 = 0 without the panel / pivinv stores, 2 chain only through smem,
//          3 pivot through ONE shuffle + off-chain smem broadcast for the updates, 4 shuffle pivot chain only,
//          5 = 3 with the l(.,j) line read as double2
__global__ void __launch_bounds__(32) potrf_variant_kernel(int variant, double *out, int reps) {
  __shared__ double a[32 * DP_LD];
  __shared__ __align__(16) double bcast[96];
  __shared__ double thr[16], pivinv[16];
  const int lane = threadIdx.x;
  for (int rep = 0; rep < reps; rep++) {
    for (int c = 0; c < 16; c++)
      a[lane * DP_LD + c] = (lane == c) ? 4.0 + 0.01 * c : 0.05 / (1.0 + abs(lane - c)) + 0.001 * ((lane * 3 + c) % 7);
    if (lane < 16)
      thr[lane] = 0.0;
    __syncwarp();
    const int row = lane;
    const bool lower = lane >= 16;
    double r[16];
#pragma unroll
    for (int c = 0; c < 16; c++)
      r[c] = (lower || c <= lane) ? a[row * DP_LD + c] : 0.0;
    double *lb = bcast;
    lb[lane] = 0.0;
    lb[32 + lane] = 0.0;
    double mydiag = (lane < 16) ? a[row * DP_LD + lane] : 0.0;
    double dcur = __shfl_sync(0xffffffffu, mydiag, 0);
    __syncwarp();
    long long t0 = clock64();
    if (variant <= 2) {
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const double d = dcur;
        const bool ok = (d > thr[j]) && (d > 0.0);
        const double invp = ok ? rsqrt(d) : 0.0;
        const double l = r[0] * invp;
        mydiag = fma(-l, l, mydiag);
        double *lbj = lb + (j & 1) * 32, *dbj = lb + 64 + (j & 1) * 16;
        if (lane < 16) {
          lbj[lane] = l;
          dbj[lane] = mydiag;
        }
        if (variant == 0) {
          if (lower || lane >= j)
            a[row * DP_LD + j] = l;
          if (lane == j)
            pivinv[j] = invp;
        }
        __syncwarp();
        dcur = dbj[(j + 1) & 15];
        if (variant != 2) {
          const double *lq = lbj + j;
#pragma unroll
          for (int k = 1; k < 16; k++)
            r[k - 1] = fma(-l, lq[k], r[k]);
          r[15] = 0.0;
        } else {
          r[0] = r[0] * 0.999 + dcur * 1e-3;
        }
      }
    } else if (variant == 6) {
      // pivot and the next column's own entry through ONE shuffle; the other 14 updates through shared memory, same iteration
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const double d = dcur;
        const bool ok = (d > thr[j]) && (d > 0.0);
        const double rs = rsqrt(ok ? d : 1.0);
        const double invp = ok ? rs : 0.0;
        const double l = r[0] * invp;
        const double l1 = __shfl_sync(0xffffffffu, l, (j + 1) & 31);
        dcur = fma(-l1, l1, ediag);
        const double r0n = fma(-l, l1, r[1]);
        mydiag = fma(-l, l, mydiag);
        ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
        double *lbj = lb + (j & 1) * 32;
        if (lane < 16)
          lbj[lane] = l;
        if (lower || lane >= j)
          a[row * DP_LD + j] = l;
        if (lane == j)
          pivinv[j] = invp;
        __syncwarp();
        const double *lq = lbj + j;
#pragma unroll
        for (int k = 2; k < 16; k++)
          r[k - 1] = fma(-l, lq[k], r[k]);
        r[0] = r0n;
        r[15] = 0.0;
      }
    } else if (variant == 7) {
      // as 6, software-pipelined by hand: the next TWO entries are updated eagerly through shuffles, the shared-memory
      // updates of column j are issued after the chain part of column j+1 (one warp issues in order: a stalled consumer of an
      // LDS would otherwise hold back the next pivot)
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
      double lprev = 0.0;       // l of the previous column (this lane)
      const double *lqprev = lb; // zero line (pads) for the first iteration
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const double d = dcur;
        const bool ok = (d > thr[j]) && (d > 0.0);
        const double rs = rsqrt(ok ? d : 1.0);
        const double invp = ok ? rs : 0.0;
        const double l = r[0] * invp;
        const double l1 = __shfl_sync(0xffffffffu, l, (j + 1) & 31);
        const double l2 = __shfl_sync(0xffffffffu, l, (j + 2) & 31);
        dcur = fma(-l1, l1, ediag);
        mydiag = fma(-l, l, mydiag);
        ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
        // deferred updates of the PREVIOUS column on entries k >= 3 (relative to its own origin j-1) = r[2..] now
#pragma unroll
        for (int k = 3; k < 16; k++)
          r[k - 1] = fma(-lprev, lqprev[k], r[k - 1]);
        // eager updates of this column on the next two entries, then shift
        const double r0n = fma(-l, l1, r[1]);
        const double r1n = fma(-l, l2, r[2]);
        double *lbj = lb + (j & 1) * 32;
        if (lane < 16)
          lbj[lane] = l;
        if (lower || lane >= j)
          a[row * DP_LD + j] = l;
        if (lane == j)
          pivinv[j] = invp;
#pragma unroll
        for (int k = 3; k < 16; k++)
          r[k - 2] = r[k - 1]; // shift (entries k >= 3 of this column are still to be updated: deferred)
        // after the shift: r[0] = r0n, r[1] = r1n, r[2..13] = old r[3..14] (not yet updated with this column), r[14], r[15] = 0
        r[0] = r0n;
        r[1] = r1n;
        r[14] = 0.0;
        r[15] = 0.0;
        __syncwarp();
        lprev = l;
        lqprev = lbj + j; // entry k of this column sits, after the shift, in r[k - 1]
      }
    } else if (variant == 8) {
      double q[16];
#pragma unroll
      for (int c = 0; c < 16; c++)
        q[c] = r[c];
      double e0 = q[0], e1 = q[1];
      lb[64 + lane] = 0.0;
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
      double lprev = 0.0;
      const double *lqprev = lb + 64;
      __syncwarp();
      t0 = clock64();
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const double d = dcur;
        const double u1 = __shfl_sync(0xffffffffu, e0, (j + 1) & 31);
        const double u2 = __shfl_sync(0xffffffffu, e0, (j + 2) & 31);
        const bool ok = (d > thr[j]) && (d > 0.0);
        const double rs = rsqrt(d);
        const double invp = ok ? rs : 0.0;
        const double l = e0 * invp, l1 = u1 * invp, l2 = u2 * invp;
        dcur = fma(-l1, l1, ediag);
        mydiag = fma(-l, l, mydiag);
        ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
        __syncwarp();
        const double x2 = fma(-lprev, lqprev[3], q[2]);
#pragma unroll
        for (int k = 2; k < 15; k++)
          q[k] = fma(-lprev, lqprev[k + 2], q[k + 1]);
        q[15] = 0.0;
        const double e0n = fma(-l, l1, e1);
        e1 = fma(-l, l2, x2);
        double *lbj = lb + (j & 1) * 32;
        if (lane < 16)
          lbj[lane] = l;
        if (lower || lane >= j)
          a[row * DP_LD + j] = l;
        if (lane == j)
          pivinv[j] = invp;
        e0 = e0n;
        lprev = l;
        lqprev = lbj + j;
      }
#pragma unroll
      for (int c = 0; c < 16; c++)
        r[c] = q[c] + e0 + e1;
    } else {
      // pivot chain: rsqrt -> scale -> ONE shuffle -> fma ; everything else off the chain
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1); // a(1,1) before column 0
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const double d = dcur;
        const bool ok = (d > thr[j]) && (d > 0.0);
        const double invp = ok ? rsqrt(d) : 0.0;
        const double l = r[0] * invp;
        const double l1 = __shfl_sync(0xffffffffu, l, (j + 1) & 31);
        dcur = fma(-l1, l1, ediag); // next pivot, on every lane
        mydiag = fma(-l, l, mydiag);
        ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31); // a(j+2,j+2) after column j; column j+1's term is added next round
        // note: ediag must include column j+1's update too; it is applied below through l2
        if (variant != 4) {
          double *lbj = lb + (j & 1) * 32;
          if (lane < 16)
            lbj[lane] = l;
          if (lower || lane >= j)
            a[row * DP_LD + j] = l;
          if (lane == j)
            pivinv[j] = invp;
          __syncwarp();
          const double *lq = lbj + j;
          if (variant == 5) {
            // double2 reads: j even -> lq+1 is odd; handle by reading from lq (aligned when j even) or lq+1 (aligned when j odd)
            if (j & 1) {
#pragma unroll
              for (int k = 1; k < 16; k += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(lq + k);
                r[k - 1] = fma(-l, v.x, r[k]);
                if (k + 1 < 16)
                  r[k] = fma(-l, v.y, r[k + 1]);
              }
            } else {
              r[0] = fma(-l, lq[1], r[1]);
#pragma unroll
              for (int k = 2; k < 16; k += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(lq + k);
                r[k - 1] = fma(-l, v.x, r[k]);
                if (k + 1 < 16)
                  r[k] = fma(-l, v.y, r[k + 1]);
              }
            }
          } else {
#pragma unroll
            for (int k = 1; k < 16; k++)
              r[k - 1] = fma(-l, lq[k], r[k]);
          }
          r[15] = 0.0;
        } else {
          r[0] = r[0] * 0.999 + l1 * 1e-3;
        }
      }
    }
    long long t1 = clock64();
    double sink = dcur + mydiag;
#pragma unroll
    for (int c = 0; c < 16; c++)
      sink += r[c];
    if (lane == 0) {
      out[rep] = (double)(t1 - t0);
      out[8 + rep] = sink + a[5 * DP_LD + 3] + pivinv[3];
    }
    __syncwarp();
  }
}
} // namespace ovp
extern "C" int ovp_debug_potrf_variants(ovp_ctx *h, int variant, int reps, double *out16) {
  using namespace ovp;
  Ctx *c = &h->c;
  if (reps > 8)
    reps = 8;
  potrf_variant_kernel<<<1, 32, 0, c->stream>>>(variant, c->dscal + 192, reps);
  OVP_CUDA(cudaStreamSynchronize(c->stream));
  OVP_CUDA(cudaMemcpy(out16, c->dscal + 192, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  return OVP_OK;
}

namespace ovp {
// the production chain (variant 8) under production conditions: CTA of `blockDim.x` threads, `nchain` warps run the chain on
// their own 32 x 16 panels (dynamic shared memory), the other warps wait at the barrier
template <int SYNCPOS, int PRED> __global__ void potrf_cond_kernel(int nchain, double *out, int reps, int use_generic) {
  extern __shared__ double dsm[];
  __shared__ __align__(16) double bcast[8 * 96];
  __shared__ double thr[16], pivinv[8 * 16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 16)
    thr[tid] = 0.0;
  for (int rep = 0; rep < reps; rep++) {
    long long t0 = 0, t1 = 0;
    __syncthreads();
    if (warp < nchain) {
      double *a = dsm + warp * 32 * DP_LD;
      for (int c = 0; c < 16; c++)
        a[lane * DP_LD + c] = (lane == c) ? 4.0 + 0.01 * c : 0.05 / (1.0 + abs(lane - c)) + 0.001 * ((lane * 3 + c) % 7);
      __syncwarp();
      const int row = lane;
      const bool lower = lane >= 16;
      double q[16];
#pragma unroll
      for (int c = 0; c < 16; c++)
        q[c] = (lower || c <= lane) ? a[row * DP_LD + c] : 0.0;
      double e0 = q[0], e1 = q[1];
      double *lb = bcast + warp * 96;
      lb[lane] = 0.0;
      lb[32 + lane] = 0.0;
      lb[64 + lane] = 0.0;
      double mydiag = (lane < 16) ? a[row * DP_LD + lane] : 0.0;
      double dcur = __shfl_sync(0xffffffffu, mydiag, 0);
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
      double lprev = 0.0;
      const double *lqprev = lb + 64;
      __syncwarp();
      t0 = clock64();
#pragma unroll 1
      for (int j = 0; j < 16; j++) {
        const double d = dcur;
        const double u1 = __shfl_sync(0xffffffffu, e0, (j + 1) & 31);
        const double u2 = __shfl_sync(0xffffffffu, e0, (j + 2) & 31);
        const bool ok = (d > thr[j]) && (d > 0.0);
        const double rs = rsqrt(d);
        const double invp = ok ? rs : 0.0;
        const double l = e0 * invp, l1 = u1 * invp, l2 = u2 * invp;
        dcur = fma(-l1, l1, ediag);
        mydiag = fma(-l, l, mydiag);
        ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
        if (SYNCPOS == 0 || SYNCPOS == 2)
          __syncwarp();
        const double x2 = fma(-lprev, lqprev[3], q[2]);
#pragma unroll
        for (int k = 2; k < 15; k++)
          q[k] = fma(-lprev, lqprev[k + 2], q[k + 1]);
        q[15] = 0.0;
        const double e0n = fma(-l, l1, e1);
        e1 = fma(-l, l2, x2);
        double *lbj = lb + (j & 1) * 32;
        if (lane < 16)
          lbj[lane] = l;
        if (PRED == 0) {
          if (lower || (lane >= j && warp == 0))
            a[row * DP_LD + j] = l;
          if (warp == 0 && lane == j)
            pivinv[j] = invp;
        } else {
          if (lower || lane >= j)
            a[row * DP_LD + j] = l;
          if (lane == j)
            pivinv[warp * 16 + j] = invp;
        }
        e0 = e0n;
        lprev = l;
        lqprev = lbj + j;
        if (SYNCPOS == 1 || SYNCPOS == 2)
          __syncwarp();
      }
      t1 = clock64();
      double sink = dcur + mydiag + e0 + e1;
#pragma unroll
      for (int c = 0; c < 16; c++)
        sink += q[c];
      if (lane == 0 && warp == 0) {
        out[rep] = (double)(t1 - t0);
        out[8 + rep] = sink;
      }
    }
    __syncthreads();
  }
}
} // namespace ovp
extern "C" int ovp_debug_potrf_cond(ovp_ctx *h, int nthreads, int nchain, int smem_bytes, int reps, double *out16, int mode) {
  using namespace ovp;
  Ctx *c = &h->c;
  if (reps > 8)
    reps = 8;
  auto k = potrf_cond_kernel<0, 0>;
  if (mode == 1) k = potrf_cond_kernel<1, 0>;
  if (mode == 2) k = potrf_cond_kernel<2, 0>;
  if (mode == 3) k = potrf_cond_kernel<0, 1>;
  if (mode == 4) k = potrf_cond_kernel<1, 1>;
  OVP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k<<<1, nthreads, smem_bytes, c->stream>>>(nchain, c->dscal + 192, reps, 0);
  OVP_CUDA(cudaStreamSynchronize(c->stream));
  OVP_CUDA(cudaMemcpy(out16, c->dscal + 192, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  return OVP_OK;
}
