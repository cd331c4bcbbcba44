// Lab bench for the pivot chain of cf_potrf64 (csrc/cholfused.cu): the column loop of one 16-column panel, copied, with parts switched
// off by template flags, timed with clock64 inside the kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_lab chain_lab.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CF_B 64
#define CF_LD 68
#define CF_AT(r, c) ((c) * CF_LD + (r))
__device__ __forceinline__ unsigned cf_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double cf_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cf_sts_if(unsigned addr, double v, int pred) {
  asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p st.shared.f64 [%0], %1; }" ::"r"(addr), "d"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ double fast_rsqrt(double d) { // MUFU.RSQ64H seed + one third-order step, no special-case branch
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double t = y0 * y0;
  const double e = fma(-t, d, 1.0);
  const double c = fma(e, 0.375, 0.5);
  const double t2 = y0 * e;
  return fma(c, t2, y0);
}
enum { F_DEFER = 1, F_STORE = 2, F_FASTRSQ = 4, F_SHFL = 8, F_THR = 16 };
template <int FL> __global__ void __launch_bounds__(256, 1) lab(double *out, long long *cyc, int nwarps, int reps) {
  extern __shared__ double sm[];
  double *a = sm, *x = a + CF_B * CF_LD, *thr = x + 1280, *pivinv = thr + CF_B, *bcast = pivinv + CF_B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  long long tot = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = tid; i < CF_B * CF_LD; i += 256) {
      int r = i % CF_LD, c = i / CF_LD;
      a[i] = (r == c) ? 70.0 + r : 1.0 / (1 + ((r * 7 + c * 3) % 11));
    }
    if (tid < CF_B)
      thr[tid] = 1e-11 * (70.0 + tid);
    __syncthreads();
    const int c0 = 0, nbp = 16, vw = nwarps - 1;
    if (warp_u <= vw) {
      const bool virt = warp_u == vw;
      const int row = (lane < 16) ? c0 + lane : c0 + 16 * warp + lane;
      const bool lower = lane >= 16;
      const bool rok = virt ? true : row < CF_B;
      double q[16];
#pragma unroll
      for (int c = 0; c < 16; c++)
        q[c] = (virt && lower) ? ((c == lane - 16) ? 1.0 : 0.0) : ((rok && (lower || c <= lane)) ? a[CF_AT(row, c0 + c)] : 0.0);
      double e0 = q[0], e1 = q[1];
      double *lb = bcast + warp * 96;
      lb[lane] = 0.0;
      lb[32 + lane] = 0.0;
      lb[64 + lane] = 0.0;
      double mydiag = (lane < 16) ? a[CF_AT(row, c0 + lane)] : 0.0;
      double dcur = __shfl_sync(0xffffffffu, mydiag, 0);
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
      double lprev = 0.0;
      unsigned thr_a = cf_saddr(thr + c0), piv_a = cf_saddr(pivinv + c0), lb_a = cf_saddr(lb);
      asm volatile("" : "+r"(thr_a), "+r"(piv_a), "+r"(lb_a));
      const unsigned row_a = (virt && lower) ? cf_saddr(x + (lane - 16) * 20) : cf_saddr(a + CF_AT(row < CF_B ? row : 0, c0));
      const unsigned row_s = (virt && lower) ? 8u : (unsigned)(CF_LD * 8);
      const unsigned lst_a = lb_a + 8 * lane;
      unsigned lq_a = lb_a + 64 * 8;
      int lane_r = lane;
      asm volatile("" : "+r"(lane_r));
      const int p_lo16 = lane < 16, p_row = lower && (virt || row < CF_B), p_diag = !lower && warp == 0, p_w0 = warp == 0;
      int bad = 0;
      asm volatile("bar.sync 1, %0;" ::"r"((vw + 1) * 32) : "memory");
      const long long t0 = clock64();
#pragma unroll 1
      for (int j = 0; j < nbp; j++) {
        const double d = dcur;
        const double thrj = (FL & F_THR) ? cf_lds(thr_a + 8 * j) : 1e-9;
        double u1, u2;
        if (FL & F_SHFL) {
          u1 = __shfl_sync(0xffffffffu, e0, (j + 1) & 31);
          u2 = __shfl_sync(0xffffffffu, e0, (j + 2) & 31);
        } else {
          u1 = e0 * 0.5;
          u2 = e0 * 0.25;
        }
        double lq[17];
        if (FL & F_DEFER) {
#pragma unroll
          for (int m = 3; m < 17; m++)
            lq[m] = cf_lds(lq_a + 8 * m);
        }
        const bool ok = (d > thrj) && (d > 0.0);
        const double rs = (FL & F_FASTRSQ) ? fast_rsqrt(d) : rsqrt(d);
        const double invp = ok ? rs : 0.0;
        const double l = e0 * invp, l1 = u1 * invp, l2 = u2 * invp;
        dcur = fma(-l1, l1, ediag);
        mydiag = fma(-l, l, mydiag);
        if (FL & F_SHFL)
          ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
        else
          ediag = mydiag + 60.0;
        bad |= !ok;
        double x2 = q[2];
        if (FL & F_DEFER) {
          x2 = fma(-lprev, lq[3], q[2]);
#pragma unroll
          for (int k = 2; k < 15; k++)
            q[k] = fma(-lprev, lq[k + 2], q[k + 1]);
          q[15] = 0.0;
        } else {
#pragma unroll
          for (int k = 2; k < 15; k++)
            q[k] = q[k + 1];
        }
        const double e0n = fma(-l, l1, e1);
        e1 = fma(-l, l2, x2);
        const unsigned par = (j & 1) * 256;
        if (FL & F_STORE) {
          cf_sts_if(lst_a + par, l, p_lo16);
          cf_sts_if(row_a + j * row_s, l, p_row | (p_diag & (lane_r >= j)));
          cf_sts_if(piv_a + 8 * j, invp, p_w0 & (lane_r == j));
        }
        e0 = e0n;
        lprev = l;
        lq_a = lb_a + par + 8 * j;
        __syncwarp();
      }
      const long long t1 = clock64();
      tot += t1 - t0;
      if (lane == 0 && rep == reps - 1)
        cyc[warp] = tot / reps;
      out[tid] = e0 + e1 + dcur + bad + q[5];
    }
    __syncthreads();
  }
}


struct ChainCtx {
  unsigned thr_a, piv_a, lb_a, row_a, row_s, lst_a;
  int lane_r, p_lo16, p_row, p_diag, p_w0;
};
template <int j> __device__ __forceinline__ void chain_step(double (&q)[16], double &dcur, double &ediag, double &mydiag, double &lprev, int &bad, const ChainCtx &c) {
  const double d = dcur;
  const double thrj = cf_lds(c.thr_a + 8 * j);
  const double e0 = q[j];
  const double u1 = __shfl_sync(0xffffffffu, e0, (j + 1) & 31);
  const double u2 = __shfl_sync(0xffffffffu, e0, (j + 2) & 31);
  const unsigned lqb = c.lb_a + ((j + 1) & 1) * 256; // line of column j-1
  double lq[16];
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (k >= j + 2 && j > 0)
      lq[k] = cf_lds(lqb + 8 * k);
  const bool ok = d > thrj;
  const double rs = fast_rsqrt(d);
  const double invp = ok ? rs : 0.0;
  const double l = e0 * invp, l1 = u1 * invp, l2 = u2 * invp;
  dcur = fma(-l1, l1, ediag);
  mydiag = fma(-l, l, mydiag);
  ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
  bad |= !ok;
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (k >= j + 2 && j > 0)
      q[k] = fma(-lprev, lq[k], q[k]);
  if (j + 1 < 16)
    q[(j + 1) & 15] = fma(-l, l1, q[(j + 1) & 15]);
  if (j + 2 < 16)
    q[(j + 2) & 15] = fma(-l, l2, q[(j + 2) & 15]);
  const unsigned par = (j & 1) * 256;
  cf_sts_if(c.lst_a + par, l, c.p_lo16);
  cf_sts_if(c.row_a + j * c.row_s, l, c.p_row | (c.p_diag & (c.lane_r >= j)));
  cf_sts_if(c.piv_a + 8 * j, invp, c.p_w0 & (c.lane_r == j));
  lprev = l;
  __syncwarp();
}
// Variant U: the column loop fully unrolled with static register indices: the deferred update touches only the columns that still exist
// (105 instead of 224 DFMAs per panel), one DSETP per column (thr' = max(thr, 0)), branch-free rsqrt.
template <int NB2> __global__ void __launch_bounds__(256, 1) labU(double *out, long long *cyc, int nwarps, int reps) {
  extern __shared__ double sm[];
  double *a = sm, *x = a + CF_B * CF_LD, *thr = x + 1280, *pivinv = thr + CF_B, *bcast = pivinv + CF_B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  long long tot = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = tid; i < CF_B * CF_LD; i += 256) {
      int r = i % CF_LD, c = i / CF_LD;
      a[i] = (r == c) ? 70.0 + r : 1.0 / (1 + ((r * 7 + c * 3) % 11));
    }
    if (tid < CF_B)
      thr[tid] = 1e-11 * (70.0 + tid);
    __syncthreads();
    const int c0 = 0, vw = nwarps - 1;
    if (warp_u <= vw) {
      const bool virt = warp_u == vw;
      const int row = (lane < 16) ? c0 + lane : c0 + 16 * warp + lane;
      const bool lower = lane >= 16;
      const bool rok = virt ? true : row < CF_B;
      double q[16];
#pragma unroll
      for (int c = 0; c < 16; c++)
        q[c] = (virt && lower) ? ((c == lane - 16) ? 1.0 : 0.0) : ((rok && (lower || c <= lane)) ? a[CF_AT(row, c0 + c)] : 0.0);
      double *lb = bcast + warp * 96;
      lb[lane] = 0.0;
      lb[32 + lane] = 0.0;
      lb[64 + lane] = 0.0;
      double mydiag = (lane < 16) ? a[CF_AT(row, c0 + lane)] : 0.0;
      double dcur = __shfl_sync(0xffffffffu, mydiag, 0);
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
      double lprev = 0.0;
      unsigned thr_a = cf_saddr(thr + c0), piv_a = cf_saddr(pivinv + c0), lb_a = cf_saddr(lb);
      asm volatile("" : "+r"(thr_a), "+r"(piv_a), "+r"(lb_a));
      const unsigned row_a = (virt && lower) ? cf_saddr(x + (lane - 16) * 20) : cf_saddr(a + CF_AT(row < CF_B ? row : 0, c0));
      const unsigned row_s = (virt && lower) ? 8u : (unsigned)(CF_LD * 8);
      const unsigned lst_a = lb_a + 8 * lane;
      int lane_r = lane;
      asm volatile("" : "+r"(lane_r));
      const int p_lo16 = lane < 16, p_row = lower && (virt || row < CF_B), p_diag = !lower && warp == 0, p_w0 = warp == 0;
      int bad = 0;
      asm volatile("bar.sync 1, %0;" ::"r"((vw + 1) * 32) : "memory");
      const long long t0 = clock64();
      ChainCtx cc{thr_a, piv_a, lb_a, row_a, row_s, lst_a, lane_r, p_lo16, p_row, p_diag, p_w0};
#define ST(J) chain_step<J>(q, dcur, ediag, mydiag, lprev, bad, cc);
      ST(0) ST(1) ST(2) ST(3) ST(4) ST(5) ST(6) ST(7) ST(8) ST(9) ST(10) ST(11) ST(12) ST(13) ST(14) ST(15)
#undef ST
      const long long t1 = clock64();
      tot += t1 - t0;
      if (lane == 0 && rep == reps - 1)
        cyc[warp] = tot / reps;
      {
        // filler: NB2 further instructions of straight-line code between two runs of the chain (what the trailing update, barriers and
        // the next panel's set-up are in the kernel) - does the chain still sit in the instruction cache next time round?
        unsigned fx = tid, fy = (unsigned)bad;
#pragma unroll
        for (int i = 0; i < NB2 / 2; i++) {
          asm volatile("add.u32 %0, %0, %1;" : "+r"(fx) : "r"(fy));
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(fy) : "r"(fx));
        }
        bad += (int)(fx + fy == 12345u);
      }
      out[tid] = q[15] + dcur + bad + q[5];
    }
    __syncthreads();
  }
}
template <int NB2> void runU(const char *name) {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 256 * 8);
  cudaMalloc(&cyc, 8 * 8);
  size_t smem = (CF_B * CF_LD + 1280 + 2 * CF_B + 8 * 96) * 8;
  cudaFuncSetAttribute(labU<NB2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int nw : {1, 4, 5}) {
    labU<NB2><<<1, 256, smem>>>(out, cyc, nw, 20);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("%-44s warps %d: %6.1f cycles / column (warp 0), %6.1f (last warp)   %s\n", name, nw, h[0] / 16.0, h[nw - 1] / 16.0, cudaGetErrorString(cudaGetLastError()));
  }
}
template <int FL> void run(const char *name) {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 256 * 8);
  cudaMalloc(&cyc, 8 * 8);
  size_t smem = (CF_B * CF_LD + 1280 + 2 * CF_B + 8 * 96) * 8;
  cudaFuncSetAttribute(lab<FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int nw : {1, 4, 5}) {
    lab<FL><<<1, 256, smem>>>(out, cyc, nw, 20);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("%-44s warps %d: %6.1f cycles / column (warp 0), %6.1f (last warp)   %s\n", name, nw, h[0] / 16.0, h[nw - 1] / 16.0, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
  cudaFree(cyc);
}
int main() {
  run<F_DEFER | F_STORE | F_SHFL | F_THR>("as in cholfused.cu");
  runU<0>("unrolled, exact-width update, fast rsqrt");
  runU<128>("  + 128 filler instructions between runs");
  runU<256>("  + 256 filler instructions between runs");
  runU<512>("  + 512 filler instructions between runs");
  runU<1024>("  + 1024 filler instructions between runs");
  runU<2048>("  + 2048 filler instructions between runs");
  runU<4096>("  + 4096 filler instructions between runs");
  run<F_DEFER | F_STORE | F_SHFL | F_THR | F_FASTRSQ>("branch-free rsqrt");
  run<F_STORE | F_SHFL | F_THR>("no deferred update");
  run<F_DEFER | F_SHFL | F_THR>("no stores");
  run<F_DEFER | F_STORE | F_THR>("no shuffles");
  run<F_DEFER | F_STORE | F_SHFL>("no threshold load");
  run<F_SHFL>("chain + shuffles only");
  run<0>("chain only");
  run<F_FASTRSQ>("chain only, branch-free rsqrt");
  return 0;
}
