// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) tile GEMM used by every dense contraction of the path:
// M = P[:,ids] H^T, S = H M[ids,:] + R, the Cholesky trailing updates, the triangular-inverse merges, Y = M L^-T and
// P -= Y Y^T (StateHelper.cpp:142-171 restructured).  tcgen05 has no f64 kind, and the path needs fp64
// (DESIGN.md "precision"), so the legacy-shaped DMMA instruction is the tensor-core instruction available for it.
//
// Operands are strided views (element (i,k) = p[i*si + K(k)*sk], optional gather K(k) = kidx[k] along the contraction
// dimension) so that transposes and the P[:, ids] / M[ids, :] gathers of EKFUpdate need no copies and no divergent code:
// the tile loaders are straight-line, all global loads of a k-step are issued before any is consumed, and the next k-step
// is prefetched into registers while the current one is in the tensor pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovp {

// Address of a shared-memory array, pinned in a register.  nvcc 12.9 for sm_100a treats the address of every shared array (static, and
// the dynamic block) as a constant it may REMATERIALISE at each use as (SR_CgaCtaId << 24) + offset: one S2R (tens of cycles, and the
// consumer waits on it) in front of every inner loop that touches shared memory.  The opaque asm stops that; going through
// shared -> generic keeps the address space known, so accesses through the returned pointer are still LDS / STS.
template <typename T> __device__ __forceinline__ T *pin_shared(T *p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("" : "+r"(a));
  return (T *)__cvta_shared_to_generic((size_t)a);
}

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Host-side logical matrix view: element (i,j) = p[row(i) + col(j)*ld] (after optional transpose), optional gather indices.
struct MatView {
  const double *p;
  int ld;
  const int *ridx; // physical row index per logical row (nullptr = identity)
  const int *cidx; // physical col index per logical col (nullptr = identity)
  int trans;       // 1: logical (i,j) reads physical (j,i)
  __device__ __forceinline__ double at(int i, int j) const {
    if (trans) {
      int t = i;
      i = j;
      j = t;
    }
    int r = ridx ? ridx[i] : i;
    int c = cidx ? cidx[j] : j;
    return p[(size_t)c * (size_t)ld + (size_t)r];
  }
};
inline MatView mv(const double *p, int ld, int trans = 0, const int *ridx = nullptr, const int *cidx = nullptr) {
  MatView v;
  v.p = p;
  v.ld = ld;
  v.ridx = ridx;
  v.cidx = cidx;
  v.trans = trans;
  return v;
}

// Device-side strided operand: element (x, k) = p[x * sx + K(k) * sk]   (x = row of A or column of B)
struct SView {
  const double *p;
  long long sx, sk;
  const int *kidx;
};

enum { TRI_FULL = 0, TRI_LOWER = 1, TRI_LOWER_MIRROR = 2 };

// C[i + j*ldc] = alpha * sum_k A(i,k) B(k,j) + beta * C + (i==j ? diag_add[i] or diag_const : 0)
struct GemmProblem {
  int M, N, K;
  SView A; // M x K
  SView B; // K x N (x = column j)
  double *C;
  int ldc;
  double alpha, beta;
  const double *diag_add; // optional, length >= min(M,N)
  double diag_const;
  int tri;     // TRI_*: LOWER computes tiles with tile_row >= tile_col only; MIRROR additionally writes C(j,i)
  int a_kfast; // 1: A is contiguous along k in memory (tile loader walks k fastest)
  int b_kfast;
  int ktri;    // structural zeros along k: 1 = B(k, j) == 0 for k < j (B lower trapezoidal: a Cholesky factor used as H^T), 2 = A(i, k) == 0
               // for k < i (its transpose on the left).  A tile starts its k loop at its own first column / row instead of 0: the skipped
               // products are exact zeros, so the result is bit-identical and M = P[:, ids] L, S = L^T M[ids, :] cost half / a third.
};
#define OVP_GEMM_MAX_BATCH 8
struct GemmBatch {
  GemmProblem p[OVP_GEMM_MAX_BATCH];
  int n;
  const int *flag; // optional device flag: when non-null and *flag == 0 the whole launch is a no-op
};

#define OVP_GT 64  // large tile edge
#define OVP_GK 16  // k step
#define OVP_GLD 68 // smem leading dim of the 64-tile (68 mod 16 == 4: conflict-free DMMA fragment reads)

// Tile loaders.  TILE x 16 elements per operand per k-step, 128 threads => NL = TILE / 8 elements per thread.
template <int TILE, bool GATHER>
__device__ __forceinline__ void load_tile_regs(double (&r)[TILE / 8], const SView &v, int x0, int X, int k0, int K, int kfast, int tid) {
#pragma unroll
  for (int t = 0; t < TILE / 8; t++) {
    int xx, kk;
    if (kfast) {
      kk = tid & 15;
      xx = (tid >> 4) + 8 * t;
    } else {
      xx = tid & (TILE - 1);
      kk = tid / TILE + (128 / TILE) * t;
    }
    int gx = x0 + xx, gk = k0 + kk;
    bool ok = (gx < X) && (gk < K);
    long long kphys = gk;
    if (GATHER)
      kphys = ok ? (long long)v.kidx[gk] : 0;
    r[t] = ok ? v.p[(long long)gx * v.sx + kphys * v.sk] : 0.0;
  }
}
// Staging layout in shared memory: element (k, x) at k * sk + x * sx.  An operand walked x-fastest is stored [k][x] (sk = TILE + 4,
// sx = 1); one walked k-fastest (k contiguous in global memory) is stored [x][k] with stride 20 (sk = 1, sx = 20): either way the
// staging stores of a half warp and the DMMA fragment reads (8 x, 4 k) fall on 16 distinct 8-byte banks (TILE + 4 and 20 are 4 mod 16).
#define OVP_GKS 20
template <int TILE>
__device__ __forceinline__ void store_tile_smem(const double (&r)[TILE / 8], double *sm, int kfast, int tid) {
#pragma unroll
  for (int t = 0; t < TILE / 8; t++) {
    int xx, kk;
    if (kfast) {
      kk = tid & 15;
      xx = (tid >> 4) + 8 * t;
      sm[xx * OVP_GKS + kk] = r[t];
    } else {
      xx = tid & (TILE - 1);
      kk = tid / TILE + (128 / TILE) * t;
      sm[kk * (TILE + 4) + xx] = r[t];
    }
  }
}

// (Round 2, measured and not kept: gather indices of the tile's k range in shared memory + a second register stage (loads two k-steps ahead):
// ncu attributes 40 % of the gathered products' stall samples to the long scoreboard of the prefetch (index load, then element load), yet the
// variant measured 0.703 vs 0.685 ms per step for the 27 products and 22.8 vs 24.7 TFLOP/s on the 2048^3 self-test (240 registers on the 64-tile).)
// (Round 2, measured and not kept: per-thread tile map hoisted out of the k loop and pinned (row pointers, shared-memory slots), all loads of a
// k-step unconditional with clamped indices: 350 -> 210 SASS instructions per k-step, no branch regions, 126 registers - and 1.03 instead of
// 0.675 ms per step for the 27 products.  The loop is not bound by its instruction count.)
// (Round 2, measured and not kept: a 4-stage cp.async operand pipeline instead of the register prefetch of the next k-step - 0.72 vs 0.71 ms
// per step for the 27 products.  The k-step is bound by the FP64 pipe, not by the loads: 16 DMMAs per warp x ~17 cycles of issue each per
// sub-partition; 240 tiles of 32 x 32 on 148 SMs are 2 rounds of 30 k-steps = 17 K cycles against 13 K at perfect balance.)
// TILE = 64: 4 warps x (32x32) ; TILE = 32: 4 warps x (16x16) — the small tile spreads mid-size problems over all 148 SMs
template <int TILE, bool GA, bool GB> __global__ void __launch_bounds__(128) gemm_f64_kernel(GemmBatch batch) {
  if (batch.flag && *batch.flag == 0)
    return;
  const GemmProblem &pb = batch.p[blockIdx.z];
  const int tm = blockIdx.y, tn = blockIdx.x;
  if (tm * TILE >= pb.M || tn * TILE >= pb.N)
    return;
  if (pb.tri != TRI_FULL && tm < tn)
    return;
  constexpr int WT = TILE / 2;  // warp tile edge
  constexpr int NM = WT / 8;    // mma tiles per warp per dimension
  constexpr int SMT = (OVP_GK * (TILE + 4) > TILE * OVP_GKS) ? OVP_GK * (TILE + 4) : TILE * OVP_GKS;
  __shared__ double As[SMT];
  __shared__ double Bs[SMT];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = tm * TILE, n0 = tn * TILE;
  const int M = pb.M, N = pb.N, K = pb.K;
  const SView va = pb.A, vb = pb.B;
  const int akf = pb.a_kfast, bkf = pb.b_kfast;
  const int ask = akf ? 1 : TILE + 4, asx = akf ? OVP_GKS : 1, bsk = bkf ? 1 : TILE + 4, bsx = bkf ? OVP_GKS : 1;
  double acc[NM][NM][2];
#pragma unroll
  for (int i = 0; i < NM; i++)
#pragma unroll
    for (int j = 0; j < NM; j++)
      acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[TILE / 8], rb[TILE / 8];
  const int kbeg = (pb.ktri == 1) ? (n0 & ~(OVP_GK - 1)) : ((pb.ktri == 2) ? (m0 & ~(OVP_GK - 1)) : 0);
  load_tile_regs<TILE, GA>(ra, va, m0, M, kbeg, K, akf, tid);
  load_tile_regs<TILE, GB>(rb, vb, n0, N, kbeg, K, bkf, tid);
  // epilogue operands that do not depend on the product are fetched now, off the critical path
  const double alpha = pb.alpha, beta = pb.beta;
  double *Cp = pb.C;
  const int ldc = pb.ldc, tri = pb.tri;
  double cin[NM][NM][2];
  if (TILE == 32) {
#pragma unroll
  for (int i = 0; i < NM; i++)
#pragma unroll
    for (int j = 0; j < NM; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int gi = m0 + wm * WT + i * 8 + (lane >> 2);
        int gj = n0 + wn * WT + j * 8 + (lane & 3) * 2 + h;
        cin[i][j][h] = (beta != 0.0 && gi < M && gj < N) ? Cp[(size_t)gj * ldc + gi] : 0.0;
      }
  }
  for (int k0 = kbeg; k0 < K; k0 += OVP_GK) {
    store_tile_smem<TILE>(ra, As, akf, tid);
    store_tile_smem<TILE>(rb, Bs, bkf, tid);
    __syncthreads();
    if (k0 + OVP_GK < K) { // prefetch the next k-step while this one is in the tensor pipe
      load_tile_regs<TILE, GA>(ra, va, m0, M, k0 + OVP_GK, K, akf, tid);
      load_tile_regs<TILE, GB>(rb, vb, n0, N, k0 + OVP_GK, K, bkf, tid);
    }
#pragma unroll
    for (int kk = 0; kk < OVP_GK; kk += 4) {
      double a[NM], b[NM];
#pragma unroll
      for (int i = 0; i < NM; i++)
        a[i] = As[(kk + (lane & 3)) * ask + (wm * WT + i * 8 + (lane >> 2)) * asx];
#pragma unroll
      for (int j = 0; j < NM; j++)
        b[j] = Bs[(kk + (lane & 3)) * bsk + (wn * WT + j * 8 + (lane >> 2)) * bsx];
#pragma unroll
      for (int i = 0; i < NM; i++)
#pragma unroll
        for (int j = 0; j < NM; j++)
          dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  if (TILE != 32) { // large tile: fetch C after the main loop (all loads in flight together), registers are free now
#pragma unroll
  for (int i = 0; i < NM; i++)
#pragma unroll
    for (int j = 0; j < NM; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int gi = m0 + wm * WT + i * 8 + (lane >> 2);
        int gj = n0 + wn * WT + j * 8 + (lane & 3) * 2 + h;
        cin[i][j][h] = (beta != 0.0 && gi < M && gj < N) ? Cp[(size_t)gj * ldc + gi] : 0.0;
      }
  }
#pragma unroll
  for (int i = 0; i < NM; i++)
#pragma unroll
    for (int j = 0; j < NM; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int gi = m0 + wm * WT + i * 8 + (lane >> 2);
        int gj = n0 + wn * WT + j * 8 + (lane & 3) * 2 + h;
        if (gi < M && gj < N) {
          double v = alpha * acc[i][j][h] + beta * cin[i][j][h];
          if (gi == gj)
            v += pb.diag_add ? pb.diag_add[gi] : pb.diag_const;
          if (tri == TRI_FULL) {
            Cp[(size_t)gj * ldc + gi] = v;
          } else if (gi >= gj) { // lower part of the (diagonal) tile
            Cp[(size_t)gj * ldc + gi] = v;
            if (tri == TRI_LOWER_MIRROR && gi != gj && gi < N && gj < M)
              Cp[(size_t)gi * ldc + gj] = v;
          }
        }
      }
}

// A: logical M x K view; gathers are supported along K only
inline SView sview_A(const MatView &v) {
  SView s;
  s.p = v.p;
  if (!v.trans) {
    s.sx = 1;
    s.sk = v.ld;
    s.kidx = v.cidx;
  } else {
    s.sx = v.ld;
    s.sk = 1;
    s.kidx = v.ridx;
  }
  return s;
}
// B: logical K x N view
inline SView sview_B(const MatView &v) {
  SView s;
  s.p = v.p;
  if (!v.trans) {
    s.sk = 1;
    s.sx = v.ld;
    s.kidx = v.ridx;
  } else {
    s.sx = 1;
    s.sk = v.ld;
    s.kidx = v.cidx;
  }
  return s;
}

inline GemmProblem make_problem(int M, int N, int K, MatView A, MatView B, double *C, int ldc, double alpha = 1.0, double beta = 0.0) {
  GemmProblem p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.A = sview_A(A);
  p.B = sview_B(B);
  p.C = C;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  p.diag_add = nullptr;
  p.diag_const = 0.0;
  p.tri = TRI_FULL;
  p.ktri = 0;
  // loader walk: along whichever logical direction is contiguous in memory
  p.a_kfast = (p.A.sk == 1) ? 1 : 0;
  p.b_kfast = (p.B.sk == 1) ? 1 : 0;
  return p;
}

} // namespace ovp
