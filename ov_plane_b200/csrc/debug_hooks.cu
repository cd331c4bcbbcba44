// TEST / TUNING HOOKS — compiled only into libovp_debug.so (-DOVP_DEBUG), never into the product library libovp.so.
// Declared in include/ovp_debug.h.  Used by tools/microbench*.py (latency and phase measurements of the fused Cholesky) and by
// tests/test_gpu_cholfused.py (unit test of chol_fused_kernel against NumPy).
#include "ovp_internal.h"
using namespace ovp;

// ---- micro-benchmarks of single kernels (tools/microbench.py) ------------------------------------------------------------
namespace ovp {
// dependent-chain latencies of fp64 operations on this GPU (cycles per op), one warp
__global__ void fp64_latency_kernel(double *out, double seed) {
  double x = seed + threadIdx.x * 1e-9;
  long long t0, t1;
  const int N = 256;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    x = fma(x, 1.0000001, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0)
    out[0] = (double)(t1 - t0) / N;
  double y = x;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++)
    y = rsqrt(y) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0)
    out[1] = (double)(t1 - t0) / N;
  double z = y;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++)
    z = 1.0 / z + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0)
    out[2] = (double)(t1 - t0) / N;
  double w = z;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++)
    w = sqrt(w) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0)
    out[3] = (double)(t1 - t0) / N;
  // shuffle of a double, dependent
  double s = w;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0)
    out[4] = (double)(t1 - t0) / N;
  // shared-memory dependent load chain
  __shared__ double sh[64];
  sh[threadIdx.x] = (double)((threadIdx.x * 7 + 3) & 31);
  __syncwarp();
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    idx = (int)sh[idx];
  t1 = clock64();
  if (threadIdx.x == 0)
    out[5] = (double)(t1 - t0) / N;
  // float rsqrt + 2 Newton steps in fp64 (candidate fast path)
  double q = s + 2.0 + idx;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) {
    double y0 = (double)rsqrtf((float)q);
    y0 = y0 * fma(-0.5 * q * y0, y0, 1.5);
    y0 = y0 * fma(-0.5 * q * y0, y0, 1.5);
    q = y0 + 1.5;
  }
  t1 = clock64();
  if (threadIdx.x == 0) {
    out[6] = (double)(t1 - t0) / N;
    out[7] = x + y + z + w + s + q;
  }
  // DMMA m8n8k4: dependent chain (latency) and 8 independent accumulators (issue rate of one warp)
  double c0 = 0.0, c1 = 0.0, aa = 1.0 + 1e-9 * threadIdx.x, bb = 1.0 - 1e-9 * threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    dmma_m8n8k4(c0, c1, aa, bb);
  t1 = clock64();
  if (threadIdx.x == 0)
    out[8] = (double)(t1 - t0) / N;
  double e[8][2];
#pragma unroll
  for (int k = 0; k < 8; k++)
    e[k][0] = e[k][1] = 0.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N / 8; i++)
#pragma unroll
    for (int k = 0; k < 8; k++)
      dmma_m8n8k4(e[k][0], e[k][1], aa, bb);
  t1 = clock64();
  double sum = c0 + c1;
#pragma unroll
  for (int k = 0; k < 8; k++)
    sum += e[k][0] + e[k][1];
  if (threadIdx.x == 0) {
    out[9] = (double)(t1 - t0) / N;
    out[10] = sum;
  }
}
} // namespace ovp
extern "C" int ovp_debug_fp64_latency(ovp_ctx *h, double *out8) {
  Ctx *c = ovp::enter(h);
  fp64_latency_kernel<<<1, 32, 0, c->stream>>>(c->dscal + 160, 1.2345);
  OVP_CUDA(cudaStreamSynchronize(c->stream));
  OVP_CUDA(cudaMemcpy(out8, c->dscal + 160, 10 * sizeof(double), cudaMemcpyDeviceToHost));
  return OVP_OK;
}


namespace ovp {
__global__ void spd_fill_kernel(double *A, int ld, int n) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n)
    return;
  int i = idx % n, j = idx / n;
  double v = (i == j) ? 4.0 + 0.001 * i : 0.3 / (1.0 + abs(i - j)) + 0.01 * ((i * 7 + j * 3) % 5);
  if (i < j)
    v = 0.0;
  else if (i != j)
    v = 0.3 / (1.0 + (i - j)) * 0.1 + 0.001 * (((i + j) * 7) % 5);
  A[(size_t)j * ld + i] = v;
}
} // namespace ovp
// microbenchmark of the fused Cholesky on a synthetic SPD n x n system (+ optional mrows x n right-hand side):
// out[0] = us per (fill + factor), out[1] = us per fill alone, out[2..] = globaltimer stamps (ns, relative to the earliest) of the
// last run, 16 per CTA
extern "C" int ovp_debug_chol_fused(ovp_ctx *h, int n, int mrows, int iters, double *out, int out_cap) {
  Ctx *c = ovp::enter(h);
  if (n > c->wsS.cap || mrows > c->Nmax)
    return fail(c, OVP_ERR_CAPACITY, "debug_chol_fused: too large");
  const int T = (n + 63) / 64;
  const int ncta = T * (T + 1) / 2 + (mrows ? (mrows + 1 + 15) / 16 : 0);
  long long *dbg = nullptr;
  OVP_CUDA(cudaMalloc(&dbg, ((size_t)ncta * 16 + 64) * sizeof(long long)));
  OVP_CUDA(cudaMemset(dbg, 0, ((size_t)ncta * 16 + 64) * sizeof(long long)));
  float ms;
  for (int variant = 0; variant < 2; variant++) {
    for (int rep = 0; rep < 2; rep++) {
      if (rep == 1)
        cudaEventRecord(c->ev[4], c->stream);
      for (int it = 0; it < iters; it++) {
        spd_fill_kernel<<<(n * n + 255) / 256, 256, 0, c->stream>>>(c->wsS.S, c->wsS.cap, n);
        if (variant == 0) {
          int st = chol_fused(c, c->wsS.S, c->wsS.cap, n, n, 0.0, mrows ? c->dM : nullptr, c->Nmax, mrows, mrows ? c->dvec + c->Rcap : nullptr, 1,
                              c->dY, c->Nmax, c->dvec, -1.0, nullptr, nullptr, dbg);
          if (st)
            return st;
        }
      }
      if (rep == 1)
        cudaEventRecord(c->ev[5], c->stream);
      OVP_CUDA(cudaStreamSynchronize(c->stream));
    }
    cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
    out[variant] = 1e3 * ms / iters;
  }
  std::vector<long long> ts((size_t)ncta * 16 + 64);
  OVP_CUDA(cudaMemcpy(ts.data(), dbg, ts.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  cudaFree(dbg);
  long long t0 = LLONG_MAX;
  const size_t nper = (size_t)ncta * 16;
  for (size_t i = 0; i < nper; i++)
    if (ts[i] > 0 && ((i & 15) < 8 || (i & 15) == 15)) // slots 8..14 are clock64 stamps, relative to slot 8 of the same CTA
      t0 = std::min(t0, ts[i]);
  out[2] = ncta;
  for (size_t i = 0; i < ts.size() && (int)(3 + i) < out_cap; i++) {
    if (i >= nper) { // spine phase stamps (clock64), relative to the first one
      out[3 + i] = ts[i] > 0 ? (double)(ts[i] - ts[nper]) : -1.0;
      continue;
    }
    const bool cyc = (i & 15) >= 8 && (i & 15) < 15;
    out[3 + i] = ts[i] > 0 ? (double)(ts[i] - (cyc ? ts[(i & ~(size_t)15) + 8] : t0)) : -1.0;
  }
  int info = 0;
  OVP_CUDA(cudaMemcpy(&info, c->dflags + 1, sizeof(int), cudaMemcpyDeviceToHost));
  if (info) {
    cudaMemset(c->dflags + 1, 0, sizeof(int));
    return fail(c, OVP_ERR_NOT_POSITIVE_DEFINITE, "debug_chol_fused: test matrix not positive definite");
  }
  return OVP_OK;
}

// Test hook for the fused Cholesky (tests/test_gpu_cholfused.py): factor a host matrix (lower triangle of A, n x n, column-major)
// over its leading npiv columns with pivot tolerance tol and, when M is given, solve Y = M L^-T (mrows x npiv) and w = L^-1 z.
// Not part of the ABI in include/ovp.h.
extern "C" int ovp_debug_chol_solve(ovp_ctx *h, const double *A, int n, int npiv, double tol, const double *M, int mrows, const double *z,
                                    double *L_out, double *Y_out, double *w_out) {
  Ctx *c = ovp::enter(h);
  if (n > c->wsS.cap || mrows > c->Nmax || npiv > n)
    return fail(c, OVP_ERR_CAPACITY, "debug_chol_solve: too large");
  const int ld = c->wsS.cap;
  OVP_CUDA(cudaMemsetAsync(c->wsS.S, 0, (size_t)ld * ld * sizeof(double), c->stream));
  OVP_CUDA(cudaMemcpy2DAsync(c->wsS.S, (size_t)ld * sizeof(double), A, (size_t)n * sizeof(double), (size_t)n * sizeof(double), n,
                             cudaMemcpyHostToDevice, c->stream));
  if (M) {
    OVP_CUDA(cudaMemcpy2DAsync(c->dM, (size_t)c->Nmax * sizeof(double), M, (size_t)mrows * sizeof(double), (size_t)mrows * sizeof(double),
                               npiv, cudaMemcpyHostToDevice, c->stream));
    OVP_CUDA(cudaMemcpyAsync(c->dvec + c->Rcap, z, (size_t)npiv * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  int st = chol_fused(c, c->wsS.S, ld, n, npiv, tol, M ? c->dM : nullptr, c->Nmax, mrows, M ? c->dvec + c->Rcap : nullptr, 1, c->dY, c->Nmax, c->dvec, -1.0,
                      c->dscal + 8, nullptr);
  if (st)
    return st;
  OVP_CUDA(cudaMemcpy2DAsync(L_out, (size_t)n * sizeof(double), c->wsS.S, (size_t)ld * sizeof(double), (size_t)n * sizeof(double), n,
                             cudaMemcpyDeviceToHost, c->stream));
  if (M) {
    OVP_CUDA(cudaMemcpy2DAsync(Y_out, (size_t)mrows * sizeof(double), c->dY, (size_t)c->Nmax * sizeof(double), (size_t)mrows * sizeof(double),
                               npiv, cudaMemcpyDeviceToHost, c->stream));
    OVP_CUDA(cudaMemcpyAsync(w_out, c->dvec, (size_t)npiv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  OVP_CUDA(cudaStreamSynchronize(c->stream));
  int info = 0;
  OVP_CUDA(cudaMemcpy(&info, c->dflags + 1, sizeof(int), cudaMemcpyDeviceToHost));
  if (info) {
    cudaMemset(c->dflags + 1, 0, sizeof(int));
    return fail(c, OVP_ERR_NOT_POSITIVE_DEFINITE, "debug_chol_solve: matrix not positive definite (strict mode)");
  }
  return OVP_OK;
}
