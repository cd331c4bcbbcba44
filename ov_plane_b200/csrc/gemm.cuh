// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) tile GEMM used by every dense contraction of the path:
// M = P[:,ids] H^T, S = H M[ids,:] + R, the Cholesky trailing updates, the triangular-inverse merges, Y = M L^-T and
// P -= Y Y^T (StateHelper.cpp:142-171 restructured).  tcgen05 has no f64 kind, and the path needs fp64
// (DESIGN.md "precision"), so the legacy-shaped DMMA instruction is the tensor-core instruction available for it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovp {

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Logical matrix view: element (i,j) = p[row(i) + col(j)*ld] (after optional transpose), with optional gather indices.
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

enum { TRI_FULL = 0, TRI_LOWER = 1, TRI_LOWER_MIRROR = 2 };

// C[i + j*ldc] = alpha * sum_k A(i,k) B(k,j) + beta * C + (i==j ? diag_add[i] or diag_const : 0)
struct GemmProblem {
  int M, N, K;
  MatView A; // M x K
  MatView B; // K x N
  double *C;
  int ldc;
  double alpha, beta;
  const double *diag_add; // optional, length >= min(M,N)
  double diag_const;
  int tri;     // TRI_*: LOWER computes tiles with tile_row >= tile_col only; MIRROR additionally writes C(j,i)
  int a_kfast; // 1: A is contiguous along k in memory (tile loader walks k fastest)
  int b_kfast;
};
#define OVP_GEMM_MAX_BATCH 8
struct GemmBatch {
  GemmProblem p[OVP_GEMM_MAX_BATCH];
  int n;
  const int *flag; // optional device flag: when non-null and *flag == 0 the whole launch is a no-op
};

#define OVP_GT 64  // tile edge
#define OVP_GK 16  // k step
#define OVP_GLD 68 // smem leading dim (68 mod 16 == 4: conflict-free DMMA fragment reads)

__global__ void __launch_bounds__(128) gemm_f64_kernel(GemmBatch batch) {
  if (batch.flag && *batch.flag == 0)
    return;
  const GemmProblem &pb = batch.p[blockIdx.z];
  const int tm = blockIdx.y, tn = blockIdx.x;
  if (tm * OVP_GT >= pb.M || tn * OVP_GT >= pb.N)
    return;
  if (pb.tri != TRI_FULL && tm < tn)
    return;
  __shared__ double As[OVP_GK][OVP_GLD];
  __shared__ double Bs[OVP_GK][OVP_GLD];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = tm * OVP_GT, n0 = tn * OVP_GT;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int k0 = 0; k0 < pb.K; k0 += OVP_GK) {
#pragma unroll
    for (int t = 0; t < 8; t++) {
      int e = tid + t * 128;
      int ii, kk;
      if (pb.a_kfast) {
        kk = e & 15;
        ii = e >> 4;
      } else {
        ii = e & 63;
        kk = e >> 6;
      }
      int gi = m0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < pb.M && gk < pb.K) ? pb.A.at(gi, gk) : 0.0;
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
      int e = tid + t * 128;
      int jj, kk;
      if (pb.b_kfast) {
        kk = e & 15;
        jj = e >> 4;
      } else {
        jj = e & 63;
        kk = e >> 6;
      }
      int gj = n0 + jj, gk = k0 + kk;
      Bs[kk][jj] = (gj < pb.N && gk < pb.K) ? pb.B.at(gk, gj) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < OVP_GK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++)
        a[i] = As[kk + (lane & 3)][wm * 32 + i * 8 + (lane >> 2)];
#pragma unroll
      for (int j = 0; j < 4; j++)
        b[j] = Bs[kk + (lane & 3)][wn * 32 + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
          dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int gi = m0 + wm * 32 + i * 8 + (lane >> 2);
        int gj = n0 + wn * 32 + j * 8 + (lane & 3) * 2 + h;
        if (gi < pb.M && gj < pb.N) {
          double v = pb.alpha * acc[i][j][h];
          if (pb.beta != 0.0)
            v += pb.beta * pb.C[(size_t)gj * pb.ldc + gi];
          if (gi == gj)
            v += pb.diag_add ? pb.diag_add[gi] : pb.diag_const;
          if (pb.tri == TRI_FULL) {
            pb.C[(size_t)gj * pb.ldc + gi] = v;
          } else if (gi >= gj) { // lower part of the (diagonal) tile
            pb.C[(size_t)gj * pb.ldc + gi] = v;
            if (pb.tri == TRI_LOWER_MIRROR && gi != gj && gi < pb.N && gj < pb.M)
              pb.C[(size_t)gi * pb.ldc + gj] = v;
          }
        }
      }
}

inline GemmProblem make_problem(int M, int N, int K, MatView A, MatView B, double *C, int ldc, double alpha = 1.0, double beta = 0.0) {
  GemmProblem p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.A = A;
  p.B = B;
  p.C = C;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  p.diag_add = nullptr;
  p.diag_const = 0.0;
  p.tri = TRI_FULL;
  // default loader hints: a view is contiguous along its logical rows unless transposed
  p.a_kfast = A.trans ? 1 : 0;
  p.b_kfast = B.trans ? 0 : 1;
  return p;
}

} // namespace ovp
