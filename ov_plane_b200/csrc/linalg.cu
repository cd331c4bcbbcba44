// Dense fp64 building blocks: batched DMMA GEMM launcher, blocked (rank-tolerant) Cholesky with diagonal-block and full
// triangular inverses, gemv / reductions.  Replaces Eigen's LLT + solveInPlace(I) + dense products of
// StateHelper::EKFUpdate (StateHelper.cpp:156-171) and the Givens triangularisation of measurement_compress_inplace
// (UpdaterHelper.cpp:548-579) by a Q-less Cholesky-QR.
#include "ovp_internal.h"
#include <algorithm>
#include <cstdarg>
#include <cstdio>

namespace ovp {

int fail(Ctx *c, int status, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  c->last_error = buf;
  return status;
}

void launch_gemm(Ctx *c, const GemmBatch &b) {
  int tm = 0, tn = 0;
  for (int i = 0; i < b.n; i++) {
    tm = std::max(tm, (b.p[i].M + OVP_GT - 1) / OVP_GT);
    tn = std::max(tn, (b.p[i].N + OVP_GT - 1) / OVP_GT);
  }
  if (tm == 0 || tn == 0 || b.n == 0)
    return;
  dim3 grid(tn, tm, b.n);
  double work = 0;
  for (int i = 0; i < b.n; i++)
    work += (b.p[i].tri == TRI_FULL ? 2.0 : 1.0) * (double)b.p[i].M * b.p[i].N * b.p[i].K;
  prof_begin(c, PROF_GEMM, work);
  const bool ga = b.p[0].A.kidx != nullptr, gb = b.p[0].B.kidx != nullptr;
  if (ga && gb)
    gemm_f64_kernel<true, true><<<grid, 128, 0, c->stream>>>(b);
  else if (ga)
    gemm_f64_kernel<true, false><<<grid, 128, 0, c->stream>>>(b);
  else if (gb)
    gemm_f64_kernel<false, true><<<grid, 128, 0, c->stream>>>(b);
  else
    gemm_f64_kernel<false, false><<<grid, 128, 0, c->stream>>>(b);
  c->launches++;
  prof_end(c);
}

static cudaEvent_t prof_event(Ctx *c) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  return c->ev_pool[c->ev_used++];
}
void prof_begin(Ctx *c, int id, double work) {
  if (!c->profiling)
    return;
  c->prof_pending = prof_event(c);
  c->prof_pending_id = id;
  c->prof_pending_work = work;
  cudaEventRecord(c->prof_pending, c->stream);
}
void prof_end(Ctx *c) {
  if (!c->profiling || !c->prof_pending)
    return;
  Ctx::ProfRec r;
  r.id = c->prof_pending_id;
  r.e0 = c->prof_pending;
  r.e1 = prof_event(c);
  r.work = c->prof_pending_work;
  cudaEventRecord(r.e1, c->stream);
  c->prof_recs.push_back(r);
  c->prof_pending = nullptr;
}
void launch_gemm1(Ctx *c, const GemmProblem &p, const int *flag) {
  if (p.M <= 0 || p.N <= 0)
    return;
  GemmBatch b;
  b.n = 1;
  b.p[0] = p;
  b.flag = flag;
  launch_gemm(c, b);
}

// -------------------------------------------------------------------------------------------------------------------
__global__ void fill_kernel(double *p, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    p[i] = v;
}
void launch_fill(Ctx *c, double *p, size_t n, double v) {
  if (n == 0)
    return;
  if (v == 0.0) {
    cudaMemsetAsync(p, 0, n * sizeof(double), c->stream);
    c->launches++;
    return;
  }
  int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  fill_kernel<<<blocks, 256, 0, c->stream>>>(p, n, v);
  c->launches++;
}

__global__ void gemv_kernel(int M, int K, MatView A, const double *x, double *y, const int *flag) {
  if (flag && *flag == 0)
    return;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= M)
    return;
  double s = 0.0;
  for (int k = lane; k < K; k += 32)
    s += A.at(warp, k) * x[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0)
    y[warp] = s;
}
void launch_gemv(Ctx *c, int M, int K, MatView A, const double *x, double *y, const int *flag) {
  if (M <= 0)
    return;
  int threads = 128;
  int blocks = (M * 32 + threads - 1) / threads;
  gemv_kernel<<<blocks, threads, 0, c->stream>>>(M, K, A, x, y, flag);
  c->launches++;
}

__global__ void sumsq_kernel(const double *x, int n, double *out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    s += x[i] * x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0)
    sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0)
      out[0] = s;
  }
}
void launch_sumsq(Ctx *c, const double *x, int n, double *out) {
  sumsq_kernel<<<1, 256, 0, c->stream>>>(x, n, out);
  c->launches++;
}

// -------------------------------------------------------------------------------------------------------------------
// Diagonal 64x64 block: Cholesky (with the zero-pivot rule) + its triangular inverse, one CTA, all in shared memory.
// -------------------------------------------------------------------------------------------------------------------
#define DB 64
#define DLD 65
__global__ void __launch_bounds__(256) potrf_diag_kernel(double *A, int ld, int bs, const double *diag0, double tol, double *Linv,
                                                         int ldi, int *info) {
  extern __shared__ double sm[];
  double(*a)[DLD] = (double(*)[DLD])sm;                  // working copy (lower), later scratch of the inverse merges
  double(*x)[DLD] = (double(*)[DLD])(sm + DB * DLD);     // inverse
  double(*l)[DLD] = (double(*)[DLD])(sm + 2 * DB * DLD); // factor L
  __shared__ double pivinv[DB];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < DB * DB; idx += 256) {
    int i = idx & 63, j = idx >> 6;
    double v;
    if (i < bs && j < bs)
      v = (i >= j) ? A[(size_t)j * ld + i] : 0.0;
    else
      v = (i == j) ? 1.0 : 0.0;
    a[i][j] = v;
    x[i][j] = 0.0;
    l[i][j] = 0.0;
  }
  // right-looking factorisation, ONE barrier per column: column j of `a` is read-only during step j (the scaled column goes
  // to `l`), the rank-1 update uses the unscaled column times 1/d.  Thread layout: row i = tid & 63, column group tid >> 6.
  const int ti = tid & 63, tg = tid >> 6;
  for (int j = 0; j < DB; j++) {
    __syncthreads();
    const double d = a[j][j];
    bool ok = true;
    if (j < bs)
      ok = (d > tol * diag0[j]) && (d > 0.0);
    const double p = ok ? sqrt(d) : 0.0;
    const double invp = ok ? 1.0 / p : 0.0;
    const double invd = ok ? 1.0 / d : 0.0;
    if (tg == 0 && ti >= j)
      l[ti][j] = (ti == j) ? p : a[ti][j] * invp;
    if (tid == 0) {
      pivinv[j] = invp;
      if (!ok && tol == 0.0 && info)
        atomicExch(info, 1); // strict mode: not positive definite
    }
    const double aij = a[ti][j] * invd;
    for (int k = j + 1 + tg; k < DB; k += 4)
      if (ti >= k)
        a[ti][k] -= aij * a[k][j];
  }
  __syncthreads();
  // write L back (lower incl. diagonal); strictly-upper part of the block is zeroed
  for (int idx = tid; idx < DB * DB; idx += 256) {
    int i = idx & 63, j = idx >> 6;
    if (i < bs && j < bs)
      A[(size_t)j * ld + i] = (i >= j) ? l[i][j] : 0.0;
  }
  // ---- inverse by recursive doubling: 8x8 base blocks, then merges at 8, 16, 32 (zero-pivot rows / columns stay zero) ----
  if (tid < 64) {
    int blk = tid >> 3, cc = tid & 7;
    int o = blk * 8;
    int c = o + cc;
    if (pivinv[c] != 0.0) {
      x[c][c] = pivinv[c];
      for (int i = c + 1; i < o + 8; i++) {
        double s = 0.0;
        for (int k = c; k < i; k++)
          s += l[i][k] * x[k][c];
        x[i][c] = -s * pivinv[i];
      }
    }
  }
  __syncthreads();
  for (int s = 8, sh = 3; s < DB; s *= 2, sh++) {
    // work item = (pair, i, j) with i fastest; s*s items per pair, npairs*s*s = 64*s/2 items in total
    const int nitems = (DB / (2 * s)) * s * s;
    // T = L21 * X11 for every pair (T lives in `a`)
    for (int idx = tid; idx < nitems; idx += 256) {
      int i = idx & (s - 1);
      int j = (idx >> sh) & (s - 1);
      int o = (idx >> (2 * sh)) * 2 * s;
      double acc = 0.0;
      for (int k = j; k < s; k++) // X11 is lower triangular: X11[k][j] = 0 for k < j
        acc += l[o + s + i][o + k] * x[o + k][o + j];
      a[o + s + i][o + j] = acc;
    }
    __syncthreads();
    // X21 = -X22 * T
    for (int idx = tid; idx < nitems; idx += 256) {
      int i = idx & (s - 1);
      int j = (idx >> sh) & (s - 1);
      int o = (idx >> (2 * sh)) * 2 * s;
      double acc = 0.0;
      for (int k = 0; k <= i; k++) // X22 lower triangular
        acc += x[o + s + i][o + s + k] * a[o + s + k][o + j];
      x[o + s + i][o + j] = -acc;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < DB * DB; idx += 256) {
    int i = idx & 63, j = idx >> 6;
    if (i < bs && j < bs)
      Linv[(size_t)j * ldi + i] = x[i][j];
  }
}

__global__ void save_diag_kernel(const double *A, int ld, int n, double *d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    d[i] = A[(size_t)i * ld + i];
}

int ws_alloc(Ctx *c, DenseWs &ws, int cap) {
  ws.cap = cap;
  size_t e = (size_t)cap * cap;
  OVP_CUDA(cudaMalloc(&ws.S, e * sizeof(double)));
  OVP_CUDA(cudaMalloc(&ws.Linv, e * sizeof(double)));
  OVP_CUDA(cudaMalloc(&ws.T, e * sizeof(double)));
  OVP_CUDA(cudaMalloc(&ws.diag0, (size_t)cap * sizeof(double)));
  OVP_CUDA(cudaMemset(ws.S, 0, e * sizeof(double)));
  OVP_CUDA(cudaMemset(ws.Linv, 0, e * sizeof(double)));
  OVP_CUDA(cudaMemset(ws.T, 0, e * sizeof(double)));
  return OVP_OK;
}
void ws_free(DenseWs &ws) {
  cudaFree(ws.S);
  cudaFree(ws.Linv);
  cudaFree(ws.T);
  cudaFree(ws.diag0);
  ws = DenseWs();
}

static bool g_potrf_attr_set = false;

int chol_partial(Ctx *c, DenseWs &ws, double *A, int ld, int n, int npiv, double tol, bool want_inverse) {
  if (npiv > ws.cap || n > ld)
    return fail(c, OVP_ERR_CAPACITY, "chol_partial: system %d exceeds workspace %d", npiv, ws.cap);
  if (npiv <= 0)
    return OVP_OK;
  const size_t smem = 3 * DB * DLD * sizeof(double);
  if (!g_potrf_attr_set) {
    OVP_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_potrf_attr_set = true;
  }
  save_diag_kernel<<<(npiv + 127) / 128, 128, 0, c->stream>>>(A, ld, npiv, ws.diag0);
  c->launches++;
  if (want_inverse)
    launch_fill(c, ws.Linv, (size_t)ws.cap * npiv, 0.0);
  int *info = c->dflags + 1; // dflags[1]: not-positive-definite indicator
  for (int j0 = 0; j0 < npiv; j0 += DB) {
    int bs = std::min(DB, npiv - j0);
    prof_begin(c, PROF_POTRF, (double)bs * bs * bs / 3.0);
    potrf_diag_kernel<<<1, 256, smem, c->stream>>>(A + (size_t)j0 * ld + j0, ld, bs, ws.diag0 + j0, tol,
                                                   ws.Linv + (size_t)j0 * ws.cap + j0, ws.cap, info);
    c->launches++;
    prof_end(c);
    int r0 = j0 + bs;
    int nb = n - r0;
    if (nb > 0) {
      // L21 = A21 * Linv11^T (in place: one tile column, every CTA owns its rows)
      GemmProblem p = make_problem(nb, bs, bs, mv(A + (size_t)j0 * ld + r0, ld), mv(ws.Linv + (size_t)j0 * ws.cap + j0, ws.cap, 1),
                                   A + (size_t)j0 * ld + r0, ld);
      launch_gemm1(c, p);
      int nc = npiv - r0;
      if (nc > 0) {
        // A22 -= L21 L21^T for columns < npiv, lower tiles
        GemmProblem q = make_problem(nb, nc, bs, mv(A + (size_t)j0 * ld + r0, ld), mv(A + (size_t)j0 * ld + r0, ld, 1),
                                     A + (size_t)r0 * ld + r0, ld, -1.0, 1.0);
        q.tri = TRI_LOWER;
        launch_gemm1(c, q);
      }
    }
  }
  if (want_inverse) {
    for (int s = DB; s < npiv; s *= 2) {
      std::vector<int> as;
      for (int a = 0; a + s < npiv; a += 2 * s)
        as.push_back(a);
      for (size_t b0 = 0; b0 < as.size(); b0 += OVP_GEMM_MAX_BATCH) {
        GemmBatch b1, b2;
        b1.flag = b2.flag = nullptr;
        b1.n = b2.n = 0;
        for (size_t i = b0; i < std::min(as.size(), b0 + OVP_GEMM_MAX_BATCH); i++) {
          int a = as[i];
          int M2 = std::min(s, npiv - (a + s));
          // T[a+s.., a..] = L[a+s.., a..a+s] * Linv[a..a+s, a..a+s]
          b1.p[b1.n++] = make_problem(M2, s, s, mv(A + (size_t)a * ld + (a + s), ld), mv(ws.Linv + (size_t)a * ws.cap + a, ws.cap),
                                      ws.T + (size_t)a * ws.cap + (a + s), ws.cap);
          // Linv[a+s.., a..] = -Linv[a+s.., a+s..] * T
          b2.p[b2.n++] = make_problem(M2, s, M2, mv(ws.Linv + (size_t)(a + s) * ws.cap + (a + s), ws.cap),
                                      mv(ws.T + (size_t)a * ws.cap + (a + s), ws.cap), ws.Linv + (size_t)a * ws.cap + (a + s), ws.cap,
                                      -1.0, 0.0);
        }
        launch_gemm(c, b1);
        launch_gemm(c, b2);
      }
    }
  }
  return OVP_OK;
}

} // namespace ovp
