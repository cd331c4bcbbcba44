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

template <int TILE> static void launch_gemm_tile(Ctx *c, const GemmBatch &b, dim3 grid) {
  const bool ga = b.p[0].A.kidx != nullptr, gb = b.p[0].B.kidx != nullptr;
  if (ga && gb)
    gemm_f64_kernel<TILE, true, true><<<grid, 128, 0, c->stream>>>(b);
  else if (ga)
    gemm_f64_kernel<TILE, true, false><<<grid, 128, 0, c->stream>>>(b);
  else if (gb)
    gemm_f64_kernel<TILE, false, true><<<grid, 128, 0, c->stream>>>(b);
  else
    gemm_f64_kernel<TILE, false, false><<<grid, 128, 0, c->stream>>>(b);
}
void launch_gemm(Ctx *c, const GemmBatch &b) {
  int tm = 0, tn = 0;
  long long tiles64 = 0;
  double work = 0;
  for (int i = 0; i < b.n; i++) {
    int a = (b.p[i].M + 63) / 64, bb = (b.p[i].N + 63) / 64;
    tm = std::max(tm, a);
    tn = std::max(tn, bb);
    tiles64 += (b.p[i].tri == TRI_FULL) ? (long long)a * bb : (long long)a * (a + 1) / 2;
    work += (b.p[i].tri == TRI_FULL ? 2.0 : 1.0) * (double)b.p[i].M * b.p[i].N * b.p[i].K;
  }
  if (tm == 0 || tn == 0 || b.n == 0)
    return;
  prof_begin(c, PROF_GEMM, work);
  if (tiles64 >= 148 || c->force_tile64) {
    launch_gemm_tile<64>(c, b, dim3(tn, tm, b.n));
  } else { // fewer 64-tiles than SMs: 32-tiles quadruple the CTA count and quarter the per-CTA tensor work
    int tm32 = 0, tn32 = 0;
    for (int i = 0; i < b.n; i++) {
      tm32 = std::max(tm32, (b.p[i].M + 31) / 32);
      tn32 = std::max(tn32, (b.p[i].N + 31) / 32);
    }
    launch_gemm_tile<32>(c, b, dim3(tn32, tm32, b.n));
  }
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
// Diagonal 64x64 block, one CTA, shared memory: right-looking Cholesky blocked by 8 columns (8x8 diagonal factor by the
// first 8 lanes of warp 0 with shuffles, row-parallel triangular solve, rank-8 trailing update: 3 barriers per 8 columns),
// then the triangular inverse by recursive doubling (8x8 bases, merges at 8/16/32).  Rank-tolerant pivots: a pivot
// <= tol * (original diagonal) (or <= 0) zeroes that row/column of L and of L^-1 (semi-definite Gram matrices).
#define POTRF_TS(k)                                                                                                         \
  if (tstamps && tid == 0)                                                                                                   \
    tstamps[k] = clock64();
__global__ void __launch_bounds__(256) potrf_diag_kernel(double *A, int ld, int bs, const double *diag0, double tol, double *Linv,
                                                         int ldi, int *info, long long *tstamps = nullptr) {
  extern __shared__ double sm[];
  double(*a)[DLD] = (double(*)[DLD])sm;                  // block, factored in place
  double(*x)[DLD] = (double(*)[DLD])(sm + DB * DLD);     // inverse
  double(*t)[DLD] = (double(*)[DLD])(sm + 2 * DB * DLD); // scratch of the inverse merges
  __shared__ double pivinv[DB];
  __shared__ double thr[DB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < DB * DB; idx += 256) {
    int i = idx & 63, j = idx >> 6;
    double v;
    if (i < bs && j < bs)
      v = (i >= j) ? A[(size_t)j * ld + i] : 0.0;
    else
      v = (i == j) ? 1.0 : 0.0;
    a[i][j] = v;
    x[i][j] = 0.0;
  }
  if (tid < DB)
    thr[tid] = (tid < bs) ? tol * diag0[tid] : 0.0;
  POTRF_TS(0)
  __syncthreads();
  POTRF_TS(1)
  // Blocked by 8 columns with one panel of look-ahead: while warp 0 runs the serial pivot chain of the 8x8 diagonal block of
  // panel p (lane = row, registers + shuffles), warps 1..7 finish the rank-8 trailing update of panel p-1 on the columns
  // beyond panel p ("part B"); the columns of panel p itself were updated first ("part A") by everyone.
  for (int c0 = 0; c0 < DB; c0 += 8) {
    if (warp == 0) {
      // (1) 8x8 diagonal factor
      double r[8];
      const int li = lane & 7;
#pragma unroll
      for (int c = 0; c < 8; c++)
        r[c] = (c <= li) ? a[c0 + li][c0 + c] : 0.0;
      double my_pinv = 0.0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double d = __shfl_sync(0xffffffffu, r[j], j);
        const bool ok = (d > thr[c0 + j]) && (d > 0.0);
        if (tol == 0.0 && !ok && info && lane == 0 && c0 + j < bs)
          atomicExch(info, 1); // strict mode: not positive definite
        const double invp = ok ? rsqrt(d) : 0.0;
        const double l = r[j] * invp;
        if (li >= j)
          r[j] = l;
        if (li == j)
          my_pinv = invp;
#pragma unroll
        for (int k = j + 1; k < 8; k++) {
          const double lk = __shfl_sync(0xffffffffu, l, k);
          if (li >= k)
            r[k] = fma(-l, lk, r[k]);
        }
      }
      if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; c++)
          if (c <= lane)
            a[c0 + lane][c0 + c] = r[c];
        pivinv[c0 + lane] = my_pinv;
      }
    } else if (c0 >= 8) {
      // part B of the previous panel (columns >= c0 + 8), 224 threads: row = wt & 63 would leave holes, so use a flat map
      const int pc = c0 - 8; // previous panel's first column
      const int wt = tid - 32;
      const int nrow = DB - (c0 + 8); // rows (and columns) still to update beyond the current panel
      if (nrow > 0) {
        // items: (row i in [c0+8, 64), column group g of 2 columns k in [c0+8, i])
        for (int it = wt; it < nrow * ((nrow + 1) / 2); it += 224) {
          const int i = c0 + 8 + it % nrow;
          const int k = c0 + 8 + 2 * (it / nrow);
          if (k > i)
            continue;
          double s0 = a[i][k], s1 = (k + 1 <= i) ? a[i][k + 1] : 0.0;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const double lij = a[i][pc + j];
            s0 = fma(-lij, a[k][pc + j], s0);
            s1 = fma(-lij, a[(k + 1 < DB) ? k + 1 : k][pc + j], s1);
          }
          a[i][k] = s0;
          if (k + 1 <= i)
            a[i][k + 1] = s1;
        }
      }
    }
    __syncthreads();
    // (2) rows below: L21 = A21 * L11^-T, one thread per row, outer-product form
    if (tid < DB - c0 - 8) {
      const int i = c0 + 8 + tid;
      double sr[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        sr[j] = a[i][c0 + j];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double xj = sr[j] * pivinv[c0 + j];
        sr[j] = xj;
#pragma unroll
        for (int k = j + 1; k < 8; k++)
          sr[k] = fma(-xj, a[c0 + k][c0 + j], sr[k]);
      }
#pragma unroll
      for (int j = 0; j < 8; j++)
        a[i][c0 + j] = sr[j];
    }
    __syncthreads();
    // (3) part A: rank-8 update of the NEXT panel's columns [c0+8, c0+16) for all rows below (lower part only)
    if (c0 + 8 < DB) {
      const int i = c0 + 8 + (tid >> 2);
      const int kk = (tid & 3) * 2;
      if (i < DB) {
        const int k = c0 + 8 + kk;
        if (k <= i) {
          double s0 = a[i][k], s1 = (k + 1 <= i) ? a[i][k + 1] : 0.0;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const double lij = a[i][c0 + j];
            s0 = fma(-lij, a[k][c0 + j], s0);
            s1 = fma(-lij, a[k + 1][c0 + j], s1);
          }
          a[i][k] = s0;
          if (k + 1 <= i)
            a[i][k + 1] = s1;
        }
      }
    }
    __syncthreads();
  }
  POTRF_TS(2)
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
          s += a[i][k] * x[k][c];
        x[i][c] = -s * pivinv[i];
      }
    }
  }
  __syncthreads();
  POTRF_TS(3)
  for (int s = 8, sh = 3; s < DB; s *= 2, sh++) {
    // work item = (pair, row i, group of 4 columns): 4 independent accumulators share every load of the left operand
    const int nitems = (DB / (2 * s)) * s * (s / 4);
    for (int idx = tid; idx < nitems; idx += 256) {
      int i = idx & (s - 1);
      int jg = ((idx >> sh) & (s / 4 - 1)) * 4;
      int o = (idx / (s * (s / 4))) * 2 * s;
      double acc[4] = {0, 0, 0, 0};
      for (int k = jg; k < s; k++) { // X11 lower triangular: X11[k][j] = 0 for k < j (zeros are stored, so k >= jg suffices)
        const double av = a[o + s + i][o + k];
#pragma unroll
        for (int q = 0; q < 4; q++)
          acc[q] = fma(av, x[o + k][o + jg + q], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        t[o + s + i][o + jg + q] = acc[q];
    }
    __syncthreads();
    for (int idx = tid; idx < nitems; idx += 256) {
      int i = idx & (s - 1);
      int jg = ((idx >> sh) & (s / 4 - 1)) * 4;
      int o = (idx / (s * (s / 4))) * 2 * s;
      double acc[4] = {0, 0, 0, 0};
      for (int k = 0; k <= i; k++) { // X22 lower triangular
        const double xv = x[o + s + i][o + s + k];
#pragma unroll
        for (int q = 0; q < 4; q++)
          acc[q] = fma(xv, t[o + s + k][o + jg + q], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
        x[o + s + i][o + jg + q] = -acc[q];
    }
    __syncthreads();
  }
  POTRF_TS(4)
  for (int idx = tid; idx < DB * DB; idx += 256) {
    int i = idx & 63, j = idx >> 6;
    if (i < bs && j < bs) {
      A[(size_t)j * ld + i] = (i >= j) ? a[i][j] : 0.0;
      Linv[(size_t)j * ldi + i] = x[i][j];
    }
  }
  POTRF_TS(5)
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
  if (c->use_fused_chol && !want_inverse)
    return chol_fused(c, A, ld, n, npiv, tol, nullptr, 0, 0, nullptr, nullptr, 0, nullptr);
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
      // L21 = A21 * Linv11^T IN PLACE: legal only with ONE tile column (every CTA then owns the rows it reads and writes), so
      // this launch must use the 64-wide tile (bs <= 64); with 32-wide tiles the two column-CTAs of a row block would overwrite
      // columns the other one is still reading.
      GemmProblem p = make_problem(nb, bs, bs, mv(A + (size_t)j0 * ld + r0, ld), mv(ws.Linv + (size_t)j0 * ws.cap + j0, ws.cap, 1),
                                   A + (size_t)j0 * ld + r0, ld);
      const bool keep = c->force_tile64;
      c->force_tile64 = true;
      launch_gemm1(c, p);
      c->force_tile64 = keep;
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

// stand-alone Householder left-nullspace projection on a global-memory matrix (col-major, ld): reflectors from the first
// nref columns applied to all ncols columns.  One CTA.
__global__ void __launch_bounds__(256) householder_cols_kernel(double *A, int ld, int rows, int ncols, int nref) {
  extern __shared__ double vbuf[];
  __shared__ double s_beta;
  const int tid = threadIdx.x;
  for (int j = 0; j < nref; j++) {
    if (tid < 32) {
      double s = 0.0;
      for (int i = j + tid; i < rows; i += 32) {
        double v = A[(size_t)j * ld + i];
        s += v * v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
      double x0 = A[(size_t)j * ld + j];
      double nrm = sqrt(s);
      double alpha = (x0 > 0.0) ? -nrm : nrm;
      double v0 = x0 - alpha;
      double vtv = s - x0 * x0 + v0 * v0;
      for (int i = j + tid; i < rows; i += 32) {
        vbuf[i] = (i == j) ? v0 : A[(size_t)j * ld + i];
        if (nrm > 0.0)
          A[(size_t)j * ld + i] = (i == j) ? alpha : 0.0; // the reflected column itself: [alpha; 0]
      }
      if (tid == 0)
        s_beta = (vtv > 0.0 && nrm > 0.0) ? 2.0 / vtv : 0.0;
    }
    __syncthreads();
    const double beta = s_beta;
    for (int cidx = j + 1 + tid; cidx < ncols; cidx += 256) {
      double *col = A + (size_t)cidx * ld;
      double s = 0.0;
      for (int i = j; i < rows; i++)
        s += vbuf[i] * col[i];
      s *= beta;
      for (int i = j; i < rows; i++)
        col[i] -= s * vbuf[i];
    }
    __syncthreads();
  }
}


} // namespace ovp
