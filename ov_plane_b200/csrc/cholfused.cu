// One-launch blocked Cholesky (+ optional triangular solve of a tall right-hand side) for the 64..1024-wide dense systems of the
// update chain: G = L L^T of the stacked Gram matrix (measurement compression) and S = L L^T, Y = M L^-T, w = L^-1 z of the
// innovation system (StateHelper.cpp:142-171 restructured).
//
// The multi-kernel version spent the whole update in launch boundaries: a 470-wide system is 8 diagonal blocks, each followed by
// a panel solve and a trailing update (24 dependent launches) plus the inverse merges.  Here every 64x64 tile of the lower
// triangle gets ONE CTA that keeps its tile in shared memory for the whole factorisation and talks to the other tiles through
// release/acquire flags in global memory (data stays in L2):
//   tile (i,j):  for k < j:  wait L(i,k), L(j,k);  tile -= L(i,k) L(j,k)^T              (DMMA, operands staged in smem)
//                i == j:     in-smem Cholesky (zero-pivot rule) + triangular inverse;   publish L(j,j), Linv(j)
//                i >  j:     wait Linv(j);  tile = tile * Linv(j)^T;                    publish L(i,j)
//   row block r of M (16 rows, optional):  for k: wait Linv(k): Y_k = M_k Linv(k)^T; for j > k: wait L(j,k): M_j -= Y_k L(j,k)^T
// CTAs only ever wait on CTAs with a smaller block index (tiles are numbered column by column, row blocks come last), and the
// hardware dispatches blocks in index order, so a waiting CTA never holds the SM its producer needs.
#include "ovp_internal.h"

namespace ovp {

#define CF_B 64
#define CF_LD 68 // 68 mod 16 == 4: DMMA fragment reads (8 rows x 4 k) hit 16 distinct 8-byte banks per half warp
#define CF_RB 16
// Tiles travel between CTAs as whole shared-memory images (64 columns x CF_LD doubles, padding included) through exchange slots in
// global memory (L2 resident): ONE cp.async.bulk (TMA bulk copy) per tile and direction, completion on an mbarrier on the loading
// side, bulk-group wait + release flag on the storing side.  The 16x16 inverses of a diagonal tile's four diagonal blocks travel
// as a compact image (4 blocks x 16 columns x CF_XLD doubles).
#define CF_SLOT (CF_B * CF_LD)            // doubles per tile slot
#define CF_SLOT_BYTES (CF_SLOT * 8)       // 34 816 B
#define CF_XLD 20                         // 20 mod 16 == 4: same bank property as CF_LD for the DMMA fragment reads of X
#define CF_XSZ (4 * 16 * CF_XLD)          // 1280 doubles
#define CF_XBYTES (CF_XSZ * 8)            // 10 240 B
#define CF_XAT(b, r, c) ((b) * (16 * CF_XLD) + (c) * CF_XLD + (r)) // element (16b + r, 16b + c) of the inverse of diagonal block b

struct CholFusedArgs {
  double *A;
  int ld, n, npiv;
  double tol;
  double *LinvD; // Tp compact inverse images (CF_XSZ doubles each, stride CF_B * CF_B)
  double *xch;   // exchange slots (CF_SLOT doubles each): L(i,k) at (k * T + i), U(i,i-1) at (T * T + i), U(j,j) at (T * T + T + j)
  double *diag0; // original diagonal (Tp * 64), written by the tile CTAs for the spine
  int *flags;    // [Tp] D, [T * Tp] P (i * Tp + k), [T] U(j,j), [T] U(i,i-1)
  int *ctrl;     // [0] epoch, [1] finished-CTA counter
  int *info;
  int T, Tp, ntile;
  const double *M; // optional tall right-hand side (mrows x npiv, ld ldm); row `mrows` of the virtual matrix is z
  int ldm, mrows;
  const double *z; // element k at z[k * zstride]
  int zstride;
  double *Y;
  int ldy;
  double *w;
  double gate_thresh; // chi2 = |w|^2 and the gate flag (thresh < 0 or chi2 <= thresh) are produced by the CTA that owns the z row
  double *chi2;
  int *gate_flag;
  int nrb, mstride;
  long long *dbg; // optional: 16 globaltimer stamps per CTA (tools/microbench.py)
  int prefactored;  // 1: second launch of the two-launch fallback - only row-block CTAs, the factor tiles already sit in the exchange slots
};

__device__ __forceinline__ long long cf_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#ifdef OVP_DEBUG
#define CF_TS(slot)                                                                                                          \
  if (p.dbg && threadIdx.x == 0)                                                                                             \
    p.dbg[(size_t)blockIdx.x * 16 + (slot)] = cf_gtime();
#else
#define CF_TS(slot) ;
#endif
// shared-memory access with explicit 32-bit addresses: inside the pivot loop the compiler otherwise re-derives the address of a
// shared array from SR_CgaCtaId (S2R / S2UR, >100 cycles each) every iteration
__device__ __forceinline__ unsigned cf_saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double cf_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double cf_lds_if(unsigned addr, int pred) {
  double v = 0.0;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.shared.f64 %0, [%1];\n\t}" : "+d"(v) : "r"(addr), "r"(pred));
  return v;
}
__device__ __forceinline__ void cf_sts_if(unsigned addr, double v, int pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr), "d"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ int cf_ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cf_st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ void cf_wait(const int *flag, int e) {
  if (threadIdx.x == 0)
    while (cf_ld_acquire(flag) != e) {
    }
  __syncthreads();
}
// all threads' global writes -> visible before the flag
__device__ __forceinline__ void cf_signal(int *flag, int e) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    cf_st_release(flag, e);
  }
}

// ---- TMA bulk copies + mbarrier (async proxy) ---------------------------------------------------------------------------------
__device__ __forceinline__ void cf_mbar_init(unsigned mb, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void cf_mbar_expect_tx(unsigned mb, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned cf_mbar_try(unsigned mb, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(mb), "r"(parity)
               : "memory");
  return ok;
}
__device__ __forceinline__ void cf_mbar_wait(unsigned mb, unsigned parity) {
  while (!cf_mbar_try(mb, parity)) {
  }
}
// global -> shared, completes `bytes` on the mbarrier (bytes multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void cf_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mb) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mb)
               : "memory");
}
// shared -> global, joins the thread's current bulk async-group
__device__ __forceinline__ void cf_bulk_s2g(void *dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cf_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cf_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); } // writes performed
// generic-proxy shared-memory writes -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void cf_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic-proxy acquire of a flag -> ordered before async-proxy reads of the data it guards
__device__ __forceinline__ void cf_fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Publish a shared-memory image through an exchange slot: called by ONE thread after a CTA barrier that follows the last
// generic-proxy write (each writer ran cf_fence_async_smem before the barrier).  Returns with the bulk group committed.
__device__ __forceinline__ void cf_publish_begin(double *slot, unsigned src, unsigned bytes) {
  cf_bulk_s2g(slot, src, bytes);
  cf_bulk_commit();
}
// ... and the release of its flag once the bulk writes have been performed
__device__ __forceinline__ void cf_publish_end(int *flag, int e) {
  cf_bulk_wait_all();
  __threadfence();
  cf_st_release(flag, e);
}
// One thread: wait for the producer's flag, then start the bulk load(s) of the guarded slot(s) onto the mbarrier
__device__ __forceinline__ void cf_acquire(const int *flag, int e) {
  while (cf_ld_acquire(flag) != e) {
  }
}

// Tiles sit in shared memory COLUMN-major like the global matrix (element (r, c) at c * CF_LD + r): global <-> shared copies
// are 16-byte chunks of two consecutive rows (a warp moves one full 512-byte column), and the DMMA fragments (8 rows x 4 k)
// still hit 16 distinct 8-byte banks per half warp because CF_LD mod 16 == 4.
#define CF_AT(r, c) ((c) * CF_LD + (r))
__device__ __forceinline__ void cf_chunk(int q, int &r2, int &c) {
  c = q >> 5;
  r2 = (q & 31) * 2;
}

// tile of a column-major matrix (g = its top-left element, ld even, 16-byte aligned) -> smem; entries outside (rv, cv), and above
// the diagonal when lower_only, are zero
__device__ __noinline__ void cf_load_tile(double *s, const double *g, int ld, int rv, int cv, bool lower_only) {
  double2 v[8]; // all loads of a thread are in flight before the first shared-memory store
#pragma unroll
  for (int q = 0; q < 8; q++) {
    int r2, c;
    cf_chunk(threadIdx.x + 256 * q, r2, c);
    const bool v0 = r2 < rv && c < cv && (!lower_only || r2 >= c), v1 = r2 + 1 < rv && c < cv && (!lower_only || r2 + 1 >= c);
    const double *src = g + (size_t)c * ld + r2;
    if (v0 && v1) {
      v[q] = __ldcg(reinterpret_cast<const double2 *>(src));
    } else {
      v[q].x = v0 ? __ldcg(src) : 0.0;
      v[q].y = v1 ? __ldcg(src + 1) : 0.0;
    }
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    int r2, c;
    cf_chunk(threadIdx.x + 256 * q, r2, c);
    *reinterpret_cast<double2 *>(s + CF_AT(r2, c)) = v[q];
  }
}
// asynchronous variant (cp.async): the copy runs while the CTA computes; invalid entries are zero-filled (src-size 0)
__device__ __noinline__ void cf_cpasync_tile(double *s, const double *g, int ld, int rv, int cv, bool lower_only) {
#pragma unroll 2
  for (int q = 0; q < 8; q++) {
    int r2, c;
    cf_chunk(threadIdx.x + 256 * q, r2, c);
    const bool v0 = r2 < rv && c < cv && (!lower_only || r2 >= c), v1 = r2 + 1 < rv && c < cv && (!lower_only || r2 + 1 >= c);
    const double *src = g + (size_t)c * ld + r2;
    const unsigned dst = cf_saddr(s + CF_AT(r2, c));
    if (v0 && v1) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    } else { // edge / diagonal-crossing chunk: through registers with L2-only loads (an 8-byte cp.async would have to be .ca, and a
             // line left in this SM's L1 by an earlier launch on the same buffers must never be served)
      double2 v;
      v.x = v0 ? __ldcg(src) : 0.0;
      v.y = v1 ? __ldcg(src + 1) : 0.0;
      *reinterpret_cast<double2 *>(s + CF_AT(r2, c)) = v;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __noinline__ void cf_store_tile(const double *s, double *g, int ld, int rv, int cv, bool lower_zero_upper) {
  double2 v[8]; // all shared-memory reads first, then the stores back to back
#pragma unroll
  for (int q = 0; q < 8; q++) {
    int r2, c;
    cf_chunk(threadIdx.x + 256 * q, r2, c);
    v[q] = *reinterpret_cast<const double2 *>(s + CF_AT(r2, c));
    if (lower_zero_upper && r2 < c)
      v[q].x = 0.0;
    if (lower_zero_upper && r2 + 1 < c)
      v[q].y = 0.0;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    int r2, c;
    cf_chunk(threadIdx.x + 256 * q, r2, c);
    double *dst = g + (size_t)c * ld + r2;
    if (lower_zero_upper && r2 + 1 < c)
      continue; // strictly upper chunk of a diagonal tile: zero in global memory already (never written by anyone)
    if (c < cv) {
      if (r2 + 1 < rv)
        __stcg(reinterpret_cast<double2 *>(dst), v[q]);
      else if (r2 < rv)
        __stcg(dst, v[q].x);
    }
  }
}

// C (64x64) = (acc_c ? C : 0) + alpha * A (64 x 64) * B^T, all tiles in smem (column-major: X(row, k) at k * CF_LD + row).
// MODE 0: full.  MODE 1: B is lower triangular (B(j, k) = 0 for k > j): column block jb only needs k < 8 jb + 8; the 8 warps take
// 32 rows x the column-block PAIR {cp, 7 - cp}, which balances the triangle (72 instead of 128 DMMA per warp).  MODE 2: only the
// lower triangle of C is wanted (symmetric update of a diagonal tile): the 36 lower 8x8 blocks are dealt 5/4 per warp.
// When C aliases A or B the caller passes inplace = true (barrier between the last read and the first write).
template <int MODE> __device__ __noinline__ void cf_mma_64(double *C, const double *A, const double *B, double alpha, bool acc_c, bool inplace) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (MODE != 2) {
    const int rb = (warp & 1) * 32, cp = warp >> 1;
    const int cb0 = 8 * cp, cb1 = 8 * (7 - cp);
    const int kend1 = (MODE == 1) ? cb1 + 8 : CF_B, kend0 = (MODE == 1) ? cb0 + 8 : CF_B;
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++)
        acc[i][j][0] = acc[i][j][1] = 0.0;
    const double *pa = A + t * CF_LD + rb + g;
    const double *pb0 = B + t * CF_LD + cb0 + g, *pb1 = B + t * CF_LD + cb1 + g;
#pragma unroll 2
    for (int k4 = 0; k4 < kend1; k4 += 4) {
      double af[4];
#pragma unroll
      for (int i = 0; i < 4; i++)
        af[i] = pa[k4 * CF_LD + 8 * i];
      const double b1v = pb1[k4 * CF_LD];
#pragma unroll
      for (int i = 0; i < 4; i++)
        dmma_m8n8k4(acc[i][1][0], acc[i][1][1], af[i], b1v);
      if (k4 < kend0) {
        const double b0v = pb0[k4 * CF_LD];
#pragma unroll
        for (int i = 0; i < 4; i++)
          dmma_m8n8k4(acc[i][0][0], acc[i][0][1], af[i], b0v);
      }
    }
    if (inplace)
      __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double *pc = C + CF_AT(rb + 8 * i + g, (j ? cb1 : cb0) + 2 * t + h);
          *pc = (acc_c ? *pc : 0.0) + alpha * acc[i][j][h];
        }
  } else {
    // row-block pair {rp, 7 - rp} owns 9 lower blocks; the even warp of the pair takes (7-rp, 0..4), the odd one the other 4
    const int rp = warp >> 1, odd = warp & 1;
    int bi[5], bj[5];
    const int nb = odd ? 4 : 5;
#pragma unroll
    for (int q = 0; q < 5; q++) {
      if (!odd) {
        bi[q] = 7 - rp;
        bj[q] = q;
      } else if (q < 3 - rp) {
        bi[q] = 7 - rp;
        bj[q] = 5 + q;
      } else {
        bi[q] = rp;
        bj[q] = q - (3 - rp);
      }
    }
    double acc[5][2];
#pragma unroll
    for (int q = 0; q < 5; q++)
      acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 2
    for (int k4 = 0; k4 < CF_B; k4 += 4) {
      const double *pk = A + (k4 + t) * CF_LD + g, *qk = B + (k4 + t) * CF_LD + g;
#pragma unroll
      for (int q = 0; q < 5; q++)
        if (q < nb)
          dmma_m8n8k4(acc[q][0], acc[q][1], pk[8 * bi[q]], qk[8 * bj[q]]);
    }
    if (inplace)
      __syncthreads();
#pragma unroll
    for (int q = 0; q < 5; q++)
      if (q < nb)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double *pc = C + CF_AT(8 * bi[q] + g, 8 * bj[q] + 2 * t + h);
          *pc = (acc_c ? *pc : 0.0) + alpha * acc[q][h];
        }
  }
}

// Batched small products on 8x8 output blocks, round-robin over the 8 warps:
//   C_b[mb*8 x nb*8] = beta * C_b + alpha * A_b[. x K] * op(B_b),  b = 0..nbatch-1, operand b at pointer + b * bstride
// NN: B is K x N (B(k, col)); otherwise N x K (B(col, k)).  lower: skip blocks above the block diagonal.  Pointers address element (0,0).
template <bool NN>
__device__ __noinline__ void cf_mma_blocks(double *C, const double *A, const double *B, int mb, int nb, int K, double alpha, double beta,
                                              bool lower, int nbatch, int bstride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int per = mb * nb;
  for (int blk = warp; blk < per * nbatch; blk += 8) {
    const int b = blk / per, q = blk - b * per;
    const int bi = q % mb, bj = q / mb;
    if (lower && bj > bi)
      continue;
    const double *pa = A + b * bstride + CF_AT(8 * bi + g, t);
    const double *pb = NN ? (B + b * bstride + CF_AT(t, 8 * bj + g)) : (B + b * bstride + CF_AT(8 * bj + g, t));
    double c0 = 0.0, c1 = 0.0;
    for (int k4 = 0; k4 < K; k4 += 4) {
      const double a = pa[k4 * CF_LD];
      const double bb = NN ? pb[k4] : pb[k4 * CF_LD];
      dmma_m8n8k4(c0, c1, a, bb);
    }
    double *pc = C + b * bstride + CF_AT(8 * bi + g, 8 * bj + 2 * t);
    pc[0] = (beta != 0.0 ? beta * pc[0] : 0.0) + alpha * c0;
    pc[CF_LD] = (beta != 0.0 ? beta * pc[CF_LD] : 0.0) + alpha * c1;
  }
}

// rank-16 update of the lower triangle of the region rows/cols [c0+16, 64) of tile a with its columns [c0, c0+16):
// 8x8 blocks of the lower triangle, up to three per warp, all operand fragments loaded before the first DMMA
__device__ __noinline__ void cf_trail16(double *a, int c0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mb = (CF_B - c0 - 16) / 8;
  const int nblk = mb * (mb + 1) / 2;
  const int o = c0 + 16;
  double fa[3][4], fb[3][4], acc[3][2];
  int bi[3], bj[3];
#pragma unroll
  for (int s = 0; s < 3; s++) {
    int q = warp + 8 * s;
    bi[s] = -1;
    bj[s] = 0;
    if (q < nblk) {
      int i = 0;
      while (q >= i + 1) {
        q -= i + 1;
        i++;
      }
      bi[s] = i;
      bj[s] = q;
    }
    acc[s][0] = acc[s][1] = 0.0;
    if (bi[s] >= 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        fa[s][k] = a[CF_AT(o + 8 * bi[s] + g, c0 + 4 * k + t)];
        fb[s][k] = a[CF_AT(o + 8 * bj[s] + g, c0 + 4 * k + t)];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int s = 0; s < 3; s++)
      if (bi[s] >= 0)
        dmma_m8n8k4(acc[s][0], acc[s][1], fa[s][k], fb[s][k]);
#pragma unroll
  for (int s = 0; s < 3; s++)
    if (bi[s] >= 0) {
      double *pc = a + CF_AT(o + 8 * bi[s] + g, o + 8 * bj[s] + 2 * t);
      pc[0] -= acc[s][0];
      pc[CF_LD] -= acc[s][1];
    }
}

// In-smem Cholesky of the leading bs columns of a 64-row tile (rows below the pivot block are solved along, columns >= bs
// receive the Schur complement).  Blocked by 16 columns.  The serial pivot chain of a 16x16 diagonal block runs in registers
// with lane = row and shuffles; lanes 16..31 of the same warp carry 16 rows BELOW the block through the same chain, and every
// further group of 16 rows below gets its own warp that repeats the diagonal block redundantly in its lanes 0..15 - so the
// whole 16-column panel is finished when the chain is, with no cross-warp traffic.  Chain per column: rsqrt -> scale ->
// shuffle -> fma (the next pivot is rebuilt on every lane from a value shuffled one column earlier).  The trailing update
// inside the tile is DMMA.  Zero-pivot rule: pivot <= thr (= tol * original diagonal) or <= 0 -> column of zeros, pivinv = 0
// (rank-deficient Gram matrices); strict (tol == 0) flags *info instead (S must be positive definite).
// 1 / sqrt(d) for the pivot chain: MUFU.RSQ64H seed + one third-order step - the arithmetic of CUDA's rsqrt(double) without its
// special-case branch (zero / denormal / inf / nan arguments), which splits the pivot loop into basic blocks.  A pivot that fails the
// threshold test never uses the value.
__device__ __forceinline__ double cf_rsqrt(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double t = y0 * y0;
  const double e = fma(-t, d, 1.0);
  const double c = fma(e, 0.375, 0.5);
  const double t2 = y0 * e;
  return fma(c, t2, y0);
}

// One column of the pivot chain with the column index (within a half panel of 8) as a template parameter: all register indices are
// static, and the 8 steps form ONE straight-line block that ptxas schedules across columns (the next pivot's rsqrt starts while this
// column's updates and stores are still being issued).  Measured on B200 (tools/lab/chain_lab.cu): 213 cycles per column for the rolled
// loop, 92 for 16 columns of straight-line code.  An FP64 instruction occupies the issue port of its SM sub-partition for ~4 cycles, so
// the chain is bound by its DFMA COUNT as much as by the rsqrt -> multiply -> fma dependency (~90 cycles): the deferred rank-1 update
// touches only the columns of the current half panel that still exist (168 DFMAs per 16 columns; rolled: 224; 16 columns unrolled: 105).
// Why 8 and not 16 columns: inside the kernel the 16-column block (12 KB of code, 255 registers) and the 8-column block (7 KB, 240
// registers) give the same end-to-end time (321.2 vs 321.0 updates/s, A/B on B200 with -DCF_CHAIN_COLS=16); neither reaches the lab
// figure in situ (115-140 cycles per column: the warp sampler shows short-scoreboard stalls on the threshold load and the shuffles that
// ptxas places late, next to their consumers, in the kernel's register allocation).  The smaller one is the default.
#ifndef CF_CHAIN_COLS
#define CF_CHAIN_COLS 8 // columns of straight-line code per loop iteration of the pivot chain: 8 (7 KB) or 16 (12 KB; A/B: -DCF_CHAIN_COLS=16)
#endif
struct CfChain {
  unsigned thr_a, piv_a, lb_a, row_a, row_s, lst_a;
  int lane_r, p_lo16, p_row, p_diag, p_w0;
};
template <int j>
__device__ __forceinline__ void cf_chain_step(double (&q)[16], int jb, double &dcur, double &ediag, double &mydiag, double &lprev, int &bad,
                                              const CfChain &c) {
  const int jg = jb + j; // column within the 16-column panel
  const double d = dcur;
  const double thrj = cf_lds(c.thr_a + 8 * jg);
  const double e0 = q[j]; // a(row, jg): final
  const double u1 = __shfl_sync(0xffffffffu, e0, (jg + 1) & 31);
  const double u2 = __shfl_sync(0xffffffffu, e0, (jg + 2) & 31);
  // l(., jg-1) of the diagonal block's 16 rows (jb is even: the parity of jg-1 is that of j+1); entries 16..31 of a line are zeros (what the
  // second half panel reads for its dead columns), and before the first column the line itself is zero
  const unsigned lqb = c.lb_a + ((j + 1) & 1) * 256 + 8 * jb;
  double lq[16];
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (k >= j + 2)
      lq[k] = cf_lds(lqb + 8 * k);
  const bool ok = d > thrj; // thr >= 0
  const double rs = cf_rsqrt(d); // speculative: a rejected pivot discards it
  const double invp = ok ? rs : 0.0;
  const double l = e0 * invp, l1 = u1 * invp, l2 = u2 * invp;
  dcur = fma(-l1, l1, ediag);
  mydiag = fma(-l, l, mydiag);
  ediag = __shfl_sync(0xffffffffu, mydiag, (jg + 2) & 31);
  bad |= !ok;
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (k >= j + 2)
      q[k] = fma(-lprev, lq[k], q[k]); // the update with column jg-1, deferred behind this column's pivot
  if (j + 1 < 16)
    q[(j + 1) & 15] = fma(-l, l1, q[(j + 1) & 15]);
  if (j + 2 < 16)
    q[(j + 2) & 15] = fma(-l, l2, q[(j + 2) & 15]);
  cf_sts_if(c.lst_a + (j & 1) * 256, l, c.p_lo16);
  cf_sts_if(c.row_a + jg * c.row_s, l, c.p_row | (c.p_diag & (c.lane_r >= jg)));
  cf_sts_if(c.piv_a + 8 * jg, invp, c.p_w0 & (c.lane_r == jg));
  lprev = l;
  __syncwarp(); // l(., jg) line complete for the next step's deferred update
}

struct CfPrefetch { // the next step's two tiles: their flags are polled by the spine's idle warp during the last pivot chain, which then
                    // starts the two bulk loads (exchange slot -> shared memory) onto `mbar`
  const int *flag0, *flag1;
  double *s0, *s1;
  const double *g0, *g1;
  unsigned mbar;
};
__device__ void cf_potrf64(double *a, double *x, int bs, const double *thr, bool strict, double *pivinv, int *info, double *bcast,
                           const CfPrefetch &pf, int epoch, long long *dbgp = nullptr) {
#ifdef OVP_DEBUG // phase stamps of the spine (tools/microbench_chol.py); the product library carries none: every stamp is a test, a branch
                 // and a few instructions of a code path whose size matters (see cf_chain_step)
#define PT(slot)                                                                                                             \
  if (dbgp && threadIdx.x == 0)                                                                                              \
    dbgp[slot] = clock64();
#else
#define PT(slot) ;
#endif
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  PT(34)
  if (tid < CF_B)
    pivinv[tid] = 0.0;
  if (bs < CF_B) // partial tile: the chain writes only bs rows of the diagonal blocks of x (compact image, CF_XAT)
    for (int idx = tid; idx < CF_XSZ; idx += 256)
      x[idx] = 0.0;
  __syncthreads();
  PT(35)
#pragma unroll 1 // one copy of the unrolled pivot chain (1.4 K instructions)
  for (int c0 = 0; c0 < bs; c0 += 16) {
    const int nbp = min(16, bs - c0);
    const int vw = (CF_B - c0 - 16) / 16; // warps 0..vw-1 carry the rows below the block, warp vw carries the identity (below)
    if (warp_u == 7) { // the I/O warp: during the last panel it polls the flags of the next step's two tiles and starts their bulk loads
      if (c0 + 16 >= bs) { // last panel: poll the flags of the next step's tiles, then start their bulk loads
        if (lane == 0 && (pf.flag0 || pf.flag1)) {
          if (pf.flag0)
            cf_acquire(pf.flag0, epoch);
          if (pf.flag1)
            cf_acquire(pf.flag1, epoch);
          cf_fence_async_all();
          cf_mbar_expect_tx(pf.mbar, (pf.flag0 ? CF_SLOT_BYTES : 0) + (pf.flag1 ? CF_SLOT_BYTES : 0));
          if (pf.flag0)
            cf_bulk_g2s(cf_saddr(pf.s0), pf.g0, CF_SLOT_BYTES, pf.mbar);
          if (pf.flag1)
            cf_bulk_g2s(cf_saddr(pf.s1), pf.g1, CF_SLOT_BYTES, pf.mbar);
        }
      }
    } else if (warp_u <= vw) {
      // warp_u is the warp index broadcast from lane 0 by a shuffle: ptxas then knows the branch is warp-uniform, emits the
      // shuffles below without divergence checks and keeps the loop body one basic block it can schedule as a whole.  Lanes
      // 0..15 of every chain warp repeat the diagonal block; lanes 16..31 of warp w < vw carry rows c0+16+16w.. of the tile;
      // lanes 16..31 of warp vw carry the rows of the 16x16 IDENTITY: what the chain solves for them is E L11^-T, i.e. the
      // inverse of the diagonal block comes out of the same chain for free (row i, column j -> Linv(j, i), stored into x).
      const bool virt = warp_u == vw;
      const int row = (lane < 16) ? c0 + lane : c0 + 16 * warp + lane;
      const bool lower = lane >= 16;
      const bool rok = virt ? true : row < CF_B;
      // Entry k of a row, relative to the current column j: e0 = a(row, j) (final), e1 = a(row, j+1) (updated through column
      // j-1), q[k] = a(row, j+k), k >= 2 (updated through column j-2: the update with column j-1 is DEFERRED into this
      // iteration, behind the pivot chain - one warp issues in order, and a consumer of a shared-memory load must not sit in
      // front of the next pivot).  Static register indices with a ROLLED column loop: the array shifts by one per column inside
      // the deferred update.  Pivot chain per column: rsqrt -> select -> 2 multiplies -> fma; the operands from other lanes
      // (u1, u2, next diagonal) are shuffled before they are needed.
#ifdef OVP_DEBUG
      if (c0 == 0) PT(50)
#endif
      double q[16];
      {
        // predicated loads, no branches: written as a conditional expression this compiled into 16 divergent branch / reconvergence
        // blocks (lanes 0..15 and 16..31 differ) - 1.2 K cycles per panel, as much as the 16 pivots themselves
        const int idn = virt && lower;
        const unsigned qa = cf_saddr(a + CF_AT(rok ? row : 0, c0));
#pragma unroll
        for (int c = 0; c < 16; c++) {
          const double v = cf_lds_if(qa + c * (CF_LD * 8), !idn && rok && (lower || c <= lane));
          q[c] = (idn && c == lane - 16) ? 1.0 : v;
        }
      }
      double e0 = q[0], e1 = q[1];
#ifdef OVP_DEBUG
      if (c0 == 0 && dbgp && tid == 0) dbgp[51] = clock64() + (long long)(e0 + e1 == 12345.678);
#endif
      double *lb = bcast + warp * 96; // [0,32) and [32,64): l(., j) by column parity (16 values + 16 zeros); [64,96): zeros
      lb[lane] = 0.0;
      lb[32 + lane] = 0.0;
      lb[64 + lane] = 0.0;
      double mydiag = (lane < 16) ? a[CF_AT(row, c0 + lane)] : 0.0;
      double dcur = __shfl_sync(0xffffffffu, mydiag, 0);
      double ediag = __shfl_sync(0xffffffffu, mydiag, 1);
      double lprev = 0.0;
#ifdef OVP_DEBUG
      if (c0 == 0 && dbgp && tid == 0) dbgp[52] = clock64() + (long long)(dcur + ediag == 12345.678);
#endif
      // loop-invariant addresses and predicates, pinned in registers
      unsigned thr_a = cf_saddr(thr + c0), piv_a = cf_saddr(pivinv + c0), lb_a = cf_saddr(lb);
      // opaque to the compiler: otherwise it REMATERIALISES the address of a static shared array inside the loop as
      // (SR_CgaCtaId << 24) + offset, i.e. one S2R per column in front of the threshold load that gates the pivot
      asm volatile("" : "+r"(thr_a), "+r"(piv_a), "+r"(lb_a));
      // where a lane stores its finished entry of column j: real rows a(row, c0 + j); identity row i: x(c0 + j, c0 + i)
      const unsigned row_a = (virt && lower) ? cf_saddr(x + CF_XAT(c0 >> 4, 0, lane - 16)) : cf_saddr(a + CF_AT(row < CF_B ? row : 0, c0));
      const unsigned row_s = (virt && lower) ? 8u : (unsigned)(CF_LD * 8);
      const unsigned lst_a = lb_a + 8 * lane;
      unsigned lq_a = lb_a + 64 * 8;
      int lane_r = lane;
      asm volatile("" : "+r"(lane_r));
      const int p_lo16 = lane < 16, p_row = lower && (virt || row < CF_B), p_diag = !lower && warp == 0, p_w0 = warp == 0;
      int bad = 0;
      // every chain warp has read the unfactored diagonal block (and its own rows) from the tile: only now may warp 0 start to
      // overwrite it.  Named barrier over the vw + 1 chain warps - without it the read races with warp 0's first store whenever a
      // warp is delayed by a few hundred cycles (seen only with several cooperative launches sharing the GPU).
#ifdef OVP_DEBUG
      if (dbgp && lane == 0 && c0 <= 16)
        dbgp[40 + 4 * (c0 >> 4) + warp] = clock64(); // arrival of each chain warp at the named barrier (panels 0 and 1)
#endif
      asm volatile("bar.sync 1, %0;" ::"r"((vw + 1) * 32) : "memory");
      PT(30 + (c0 >> 4))
      if (nbp == 16) {
        const CfChain cc{thr_a, piv_a, lb_a, row_a, row_s, lst_a, lane_r, p_lo16, p_row, p_diag, p_w0};
#define CF_ST(J) cf_chain_step<J>(q, jb, dcur, ediag, mydiag, lprev, bad, cc);
#if CF_CHAIN_COLS == 16
        {
          const int jb = 0;
          CF_ST(0) CF_ST(1) CF_ST(2) CF_ST(3) CF_ST(4) CF_ST(5) CF_ST(6) CF_ST(7)
          CF_ST(8) CF_ST(9) CF_ST(10) CF_ST(11) CF_ST(12) CF_ST(13) CF_ST(14) CF_ST(15)
        }
#else
#pragma unroll 1
        for (int jb = 0; jb < 16; jb += 8) {
          CF_ST(0) CF_ST(1) CF_ST(2) CF_ST(3) CF_ST(4) CF_ST(5) CF_ST(6) CF_ST(7)
#pragma unroll
          for (int k = 0; k < 8; k++) { // the second half panel continues with static indices 0..7
            q[k] = q[k + 8];
            q[k + 8] = 0.0;
          }
        }
#endif
#undef CF_ST
      } else {
        // partial panel (the last one of a system whose size is not a multiple of 16): the rolled form of the same step; the register
        // array shifts by one per column so that its indices stay static
#pragma unroll 1
        for (int j = 0; j < nbp; j++) {
          const double d = dcur;
          const double thrj = cf_lds(thr_a + 8 * j);
          const double u1 = __shfl_sync(0xffffffffu, e0, (j + 1) & 31);
          const double u2 = __shfl_sync(0xffffffffu, e0, (j + 2) & 31);
          double lq[17];
#pragma unroll
          for (int m = 3; m < 17; m++)
            lq[m] = cf_lds(lq_a + 8 * m); // l(j-1+m, j-1): complete since the __syncwarp that closed the previous iteration
          const bool ok = d > thrj; // thr >= 0
          const double rs = cf_rsqrt(d); // speculative: a rejected pivot discards it
          const double invp = ok ? rs : 0.0;
          const double l = e0 * invp, l1 = u1 * invp, l2 = u2 * invp;
          dcur = fma(-l1, l1, ediag);
          mydiag = fma(-l, l, mydiag);
          ediag = __shfl_sync(0xffffffffu, mydiag, (j + 2) & 31);
          bad |= !ok;
          const double x2 = fma(-lprev, lq[3], q[2]);
#pragma unroll
          for (int k = 2; k < 15; k++)
            q[k] = fma(-lprev, lq[k + 2], q[k + 1]);
          q[15] = 0.0;
          const double e0n = fma(-l, l1, e1);
          e1 = fma(-l, l2, x2);
          const unsigned par = (j & 1) * 256;
          cf_sts_if(lst_a + par, l, p_lo16);
          cf_sts_if(row_a + j * row_s, l, p_row | (p_diag & (lane_r >= j)));
          cf_sts_if(piv_a + 8 * j, invp, p_w0 & (lane_r == j));
          e0 = e0n;
          lprev = l;
          lq_a = lb_a + par + 8 * j;
          __syncwarp(); // l(., j) line complete for the next iteration's deferred update
        }
      }
      if (strict && bad && tid == 0)
        atomicExch(info, 1);
      PT(1 + 3 * (c0 >> 4))
    }
    __syncthreads();
    PT(2 + 3 * (c0 >> 4))
    if (CF_B - c0 - 16 > 0 && nbp == 16) // (a partial block is the last pivot block: everything to its right is never read)
      cf_trail16(a, c0);
    __syncthreads();
    PT(3 + 3 * (c0 >> 4))
  }
}

// U <- U L^-T for a 64-row tile U, given the factor tile Lkk and the 16x16 inverses of ITS diagonal blocks (X, from the pivot
// chain): a right-side triangular solve is independent per row, so warp w owns rows 8w..8w+7 through all four 16-column steps
//   S = U[:, b] - U[:, <b] Lkk[b, <b]^T      (U[:, <b] already holds the result)
//   U[:, b] = S X_bb^T
// with a private 8 x 16 scratch (S) and no CTA barrier at all.  No 64x64 inverse is ever formed.  X is the compact inverse image
// (CF_XAT); S is a 64 x 16 scratch in tile layout (16 columns of CF_LD).
__device__ __noinline__ void cf_bsolve64(double *U, const double *Lkk, const double *X, double *S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int r = 8 * warp + g;
#pragma unroll 1
  for (int b = 0; b < 4; b++) {
    const int cb = 16 * b;
    double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0};
#pragma unroll 4
    for (int k4 = 0; k4 < cb; k4 += 4) {
      const double av = U[CF_AT(r, k4 + t)];
      dmma_m8n8k4(a0[0], a0[1], av, Lkk[CF_AT(cb + g, k4 + t)]);
      dmma_m8n8k4(a1[0], a1[1], av, Lkk[CF_AT(cb + 8 + g, k4 + t)]);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      S[CF_AT(r, 2 * t + h)] = U[CF_AT(r, cb + 2 * t + h)] - a0[h];
      S[CF_AT(r, 8 + 2 * t + h)] = U[CF_AT(r, cb + 8 + 2 * t + h)] - a1[h];
    }
    __syncwarp();
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
      const double sv = S[CF_AT(r, k4 + t)];
      if (k4 < 8) // X_bb lower triangular: columns 0..7 only need k < 8
        dmma_m8n8k4(c0[0], c0[1], sv, X[CF_XAT(b, g, k4 + t)]);
      dmma_m8n8k4(c1[0], c1[1], sv, X[CF_XAT(b, 8 + g, k4 + t)]);
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; h++) {
      U[CF_AT(r, cb + 2 * t + h)] = c0[h];
      U[CF_AT(r, cb + 8 + 2 * t + h)] = c1[h];
    }
    __syncwarp();
  }
}

static_assert(sizeof(CholFusedArgs) <= 32 * sizeof(double), "the shared-memory copy of the arguments has 32 doubles");
__global__ void __launch_bounds__(256, 1) chol_fused_kernel(CholFusedArgs p_in) {
  extern __shared__ __align__(128) double sm_dyn[];
  double *const sm = pin_shared(sm_dyn); // (gemm.cuh) keeps the base in a register: no S2R SR_CgaCtaId in front of the loops below
  // every shared array is carved from the one dynamic block: addresses of static __shared__ variables are re-derived from
  // SR_CgaCtaId (S2R, ~100+ cycles) wherever the compiler rematerialises them - inside the pivot loop that tripled its latency.
  // Layout (doubles): [0,16) two mbarriers (+ padding to 128 B) | [16,48) the kernel arguments | tile buffers | role-specific
  //
  // The ARGUMENTS are read from this shared-memory copy, not from the constant bank: the compiler re-loads a kernel parameter with
  // LDC / LDCU wherever it needs one (31 loads spread over the spine's step), each a potential constant-cache miss behind the
  // instruction cache the step's code already overflows; as LDS they cannot miss.  Measured: 1.94 -> 1.90 ms of chol_fused per step.
  double *tb = sm + 48; // tile buffers: 128-byte aligned bulk-copy targets
  const unsigned mb0 = cf_saddr(sm);
  __shared__ int s_epoch;
  const int tid = threadIdx.x;
  if (tid == 0) {
    *reinterpret_cast<CholFusedArgs *>(sm + 16) = p_in;
    s_epoch = *(volatile int *)p_in.ctrl + 1;
    cf_mbar_init(mb0, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const CholFusedArgs &p = *reinterpret_cast<const CholFusedArgs *>(sm + 16);
  const int e = s_epoch;
  const int Tp = p.Tp, T = p.T;
  int *fdiag = p.flags, *fpan = p.flags + Tp; // D(k): L(k,k) and Linv(k) published; P(i,k): L(i,k) published
  int *fud = fpan + T * Tp, *fus = fud + T;   // U(j,j) / U(i,i-1): tile updated through all panels but the last, for the spine
  double *slotL = p.xch, *slotUs = p.xch + (size_t)T * T * CF_SLOT, *slotUd = slotUs + (size_t)T * CF_SLOT;
#define CF_SLOTL(i, k) (slotL + ((size_t)(k) * T + (i)) * CF_SLOT)
  const int IO = 224; // lane 0 of warp 7: the thread that issues bulk copies and releases flags (warp 7 never runs a pivot chain)

  if (!p.prefactored && blockIdx.x == 0) {
    // ---- spine: every diagonal block, back to back (warm instruction cache, no global-memory hop on the critical path) ----
    double *a = tb, *b3 = tb + CF_SLOT, *b4 = tb + 2 * CF_SLOT;
    double *xc = tb + 3 * CF_SLOT, *sscr = xc + CF_XSZ, *thr = sscr + 16 * CF_LD, *pivinv = thr + CF_B, *bcast = pivinv + CF_B; // bcast: 8 x 96
    unsigned par_pf = 0;
    CF_TS(0)
    cf_load_tile(a, p.A, p.ld, min(CF_B, p.n), min(CF_B, p.n), true);
    __syncthreads();
    if (tid < CF_B)
      thr[tid] = fmax(p.tol * a[CF_AT(tid, tid)], 0.0); // >= 0: the pivot loop tests d > thr only
    for (int idx = tid; idx < CF_XSZ; idx += 256)
      xc[idx] = 0.0; // the inverse image: only the lower triangles of its blocks are ever written
    for (int k = 0; k < Tp; k++) {
      const int bs = min(CF_B, p.npiv - CF_B * k);
      const bool has_panel = k + 1 < T, next_diag = k + 1 < Tp;
      if (k < 3)
        CF_TS(1 + 2 * k)
      long long *dbgp = (p.dbg && k == 1) ? p.dbg + (size_t)gridDim.x * 16 : nullptr;
      PT(0)
      CfPrefetch pf;
      pf.flag0 = has_panel ? fus + k + 1 : nullptr;
      pf.flag1 = next_diag ? fud + k + 1 : nullptr;
      pf.s0 = b3;
      pf.s1 = b4;
      pf.g0 = slotUs + (size_t)(k + 1) * CF_SLOT;
      pf.g1 = slotUd + (size_t)(k + 1) * CF_SLOT;
      pf.mbar = mb0;
      cf_potrf64(a, xc, bs, thr, p.tol == 0.0, pivinv, p.info, bcast, pf, e, dbgp);
      if (k < 3)
        CF_TS(2 + 2 * k)
      PT(13)
      // strictly-upper entries inside the diagonal 8x8 blocks were touched by the in-tile trailing updates: the factor is later read
      // as a dense operand, so they leave as zeros (nothing in this kernel reads them)
      for (int idx = tid; idx < CF_B * 8; idx += 256) {
        const int c = idx >> 3, r = (c & ~7) + (idx & 7);
        if (r < c)
          a[CF_AT(r, c)] = 0.0;
      }
      cf_fence_async_smem();
      __syncthreads();
      if (tid == IO) { // L(k,k) and its block inverses -> exchange slots (2 bulk copies), the factor columns -> the matrix
        cf_bulk_s2g(CF_SLOTL(k, k), cf_saddr(a), CF_SLOT_BYTES);
        cf_bulk_s2g(p.LinvD + (size_t)k * CF_B * CF_B, cf_saddr(xc), CF_XBYTES);
        cf_bulk_commit();
        // Release D(k) as soon as the two copies have landed.  (It used to be released after the panel solve below, "for free": that
        // held back every consumer of L(k,k) - the tiles of column k and through them the two tiles this CTA needs for step k+1 - by the
        // 5 K cycles of the solve, and once the pivot chains got faster that path became the critical one.)
        cf_publish_end(fdiag + k, e);
      }
      PT(20)
      if (has_panel) {
        if (next_diag && tid < CF_B)
          thr[tid] = fmax(p.tol * __ldcg(p.diag0 + CF_B * (k + 1) + tid), 0.0);
        cf_mbar_wait(mb0, par_pf); // the two tiles of the next step (bulk loads started during the last pivot chain)
        par_pf ^= 1;
        PT(24)
        cf_bsolve64(b3, a, xc, sscr); // L(k+1,k) = U(k+1,k) L(k,k)^-T
        cf_fence_async_smem();
        __syncthreads();
        PT(25)
        if (tid == IO) {
          cf_bulk_s2g(CF_SLOTL(k + 1, k), cf_saddr(b3), CF_SLOT_BYTES);
          cf_bulk_commit();
        }
        PT(26)
        if (next_diag) {
          cf_mma_64<2>(b4, b3, b3, -1.0, true, false); // lower triangle of the next diagonal tile
          __syncthreads();
          double *tmp = a;
          a = b4;
          b4 = tmp;
        }
        PT(27)
        if (tid == IO)
          cf_publish_end(fpan + (k + 1) * Tp + k, e);
        PT(28)
      }
    }
    if (T == 1) { // a single tile: no other CTA exists to copy the factor into the matrix (see the tile CTAs below)
      __syncthreads();
      cf_store_tile(a, p.A, p.ld, min(CF_B, p.n), min(CF_B, p.npiv), false);
    }
  } else if (!p.prefactored && (int)blockIdx.x < p.ntile) {
    double *a = tb, *b1 = tb + CF_SLOT, *b2 = tb + 2 * CF_SLOT;
    double *xc = tb + 3 * CF_SLOT, *sscr = xc + CF_XSZ;
    int j = 0, rem = blockIdx.x;
    while (rem >= T - j) {
      rem -= T - j;
      j++;
    }
    const int i = j + rem;
    const int rv = min(CF_B, p.n - CF_B * i), cv = min(CF_B, p.n - CF_B * j);
    const int bs = min(CF_B, p.npiv - CF_B * j);
    double *gA = p.A + (size_t)(CF_B * j) * p.ld + CF_B * i;
    unsigned par = 0;
    CF_TS(0)
    cf_load_tile(a, gA, p.ld, rv, cv, i == j);
    __syncthreads();
    CF_TS(1)
    if (i == j && tid < CF_B)
      __stcg(p.diag0 + CF_B * j + tid, a[CF_AT(tid, tid)]); // original diagonal: reference of the zero-pivot rule
    const int kmax = (i == j) ? j - 1 : j; // the spine applies the last update of a diagonal tile itself
    for (int k = 0; k < kmax; k++) {
      if (tid == 0) { // wait for the panel tile(s) of column block k, then one bulk load each
        cf_acquire(fpan + i * Tp + k, e);
        if (i != j)
          cf_acquire(fpan + j * Tp + k, e);
        cf_fence_async_all();
        cf_mbar_expect_tx(mb0, (i != j) ? 2 * CF_SLOT_BYTES : CF_SLOT_BYTES);
        cf_bulk_g2s(cf_saddr(b1), CF_SLOTL(i, k), CF_SLOT_BYTES, mb0);
        if (i != j)
          cf_bulk_g2s(cf_saddr(b2), CF_SLOTL(j, k), CF_SLOT_BYTES, mb0);
      }
      cf_mbar_wait(mb0, par);
      par ^= 1;
      if (i != j)
        cf_mma_64<0>(a, b1, b2, -1.0, true, false);
      else
        cf_mma_64<2>(a, b1, b1, -1.0, true, false);
      __syncthreads(); // b1 / b2 are overwritten by the next iteration's loads
    }
    CF_TS(3)
    if (i == j || i == j + 1) { // hand the tile (updated through all panels but the last) to the spine
      cf_fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        cf_publish_begin((i == j ? slotUd + (size_t)j * CF_SLOT : slotUs + (size_t)i * CF_SLOT), cf_saddr(a), CF_SLOT_BYTES);
        cf_publish_end(i == j ? fud + j : fus + i, e);
      }
      CF_TS(4)
      // The spine finishes this tile and publishes the factor tile in its exchange slot; THIS CTA (idle from here on) copies it into the
      // matrix A, which nobody in this launch reads.  (The spine used to do that itself from two idle warps during its next first pivot
      // chain: those warps share issue slots with the chain warps, and the step waited for them at the next CTA barrier.)  Tile (1,0)
      // also copies L(0,0), whose own block index is the spine's.
      for (int pass = 0; pass < ((blockIdx.x == 1) ? 2 : 1); pass++) {
        const int ti = pass ? 0 : i, tj = pass ? 0 : j;
        if (pass)
          __syncthreads(); // the first tile has left shared memory
        if (tid == 0) {
          cf_acquire(ti == tj ? fdiag + tj : fpan + ti * Tp + tj, e);
          cf_fence_async_all();
          cf_mbar_expect_tx(mb0, CF_SLOT_BYTES);
          cf_bulk_g2s(cf_saddr(a), CF_SLOTL(ti, tj), CF_SLOT_BYTES, mb0);
        }
        cf_mbar_wait(mb0, par);
        par ^= 1;
        cf_store_tile(a, p.A + (size_t)(CF_B * tj) * p.ld + CF_B * ti, p.ld, min(CF_B, p.n - CF_B * ti), min(CF_B, p.npiv - CF_B * tj), false);
      }
    } else {
      if (tid == 0) {
        cf_acquire(fdiag + j, e);
        cf_fence_async_all();
        cf_mbar_expect_tx(mb0, CF_SLOT_BYTES + CF_XBYTES);
        cf_bulk_g2s(cf_saddr(b1), CF_SLOTL(j, j), CF_SLOT_BYTES, mb0);                              // L(j,j)
        cf_bulk_g2s(cf_saddr(xc), p.LinvD + (size_t)j * CF_B * CF_B, CF_XBYTES, mb0);                // its 16x16 inverses
      }
      cf_mbar_wait(mb0, par);
      par ^= 1;
      CF_TS(5)
      cf_bsolve64(a, b1, xc, sscr);
      cf_fence_async_smem();
      __syncthreads();
      CF_TS(6)
      if (tid == 0) {
        cf_publish_begin(CF_SLOTL(i, j), cf_saddr(a), CF_SLOT_BYTES);
        cf_publish_end(fpan + i * Tp + j, e);
      }
      cf_store_tile(a, gA, p.ld, rv, bs, false); // the factor tile into the matrix (nobody in this launch reads it there)
      CF_TS(7)
    }
    if (tid == 0)
      cf_bulk_wait_all();
  } else {
    // ---- row block of the right-hand side: Y = M L^-T, right-looking ----
    const int rb = blockIdx.x - (p.prefactored ? 0 : p.ntile);
    const int ms = p.mstride;
    double *Lt = tb;                     // one tile buffer (bulk-load target, 128-byte aligned)
    double *Xt = tb + CF_SLOT;           // compact inverse image of L(k,k)'s diagonal blocks
    double *mrow = Xt + CF_XSZ;          // CF_RB x ms
    double *yk = mrow + CF_RB * ms;      // CF_RB x CF_LD (row-major)
    double *sb = yk + CF_RB * CF_LD;     // 2 warps x 8 x 20 scratch (320 doubles; chi2 reduction: 256)
    const int row0 = rb * CF_RB;
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    unsigned par = 0;
    double wsq = 0.0;
    for (int idx = tid; idx < CF_RB * Tp * CF_B; idx += 256) {
      const int r = idx & (CF_RB - 1), k = idx >> 4;
      const int row = row0 + r;
      double v = 0.0;
      if (k < p.npiv) {
        if (row < p.mrows)
          v = __ldcg(p.M + (size_t)k * p.ldm + row);
        else if (row == p.mrows && p.z)
          v = __ldcg(p.z + (size_t)k * p.zstride);
      }
      mrow[r * ms + k] = v;
    }
    __syncthreads();
    for (int k = 0; k < Tp; k++) {
      const int bs = min(CF_B, p.npiv - CF_B * k);
      if (tid == 0) {
        if (!p.prefactored)
          cf_acquire(fdiag + k, e);
        cf_fence_async_all();
        cf_mbar_expect_tx(mb0, CF_SLOT_BYTES + CF_XBYTES);
        cf_bulk_g2s(cf_saddr(Lt), CF_SLOTL(k, k), CF_SLOT_BYTES, mb0);
        cf_bulk_g2s(cf_saddr(Xt), p.LinvD + (size_t)k * CF_B * CF_B, CF_XBYTES, mb0);
      }
      cf_mbar_wait(mb0, par);
      par ^= 1;
      if (warp < 2) { // yk (16 x 64) = mrow[:, 64k ..] L(k,k)^-T: rows are independent, warp w owns rows 8w..8w+7 (cf_bsolve64 in row-major)
        const int r = 8 * warp + g;
        double *sw = sb + warp * 160;
#pragma unroll 1
        for (int b = 0; b < 4; b++) {
          const int cb = 16 * b;
          double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0};
          for (int k4 = 0; k4 < cb; k4 += 4) {
            const double av = yk[r * CF_LD + k4 + t];
            dmma_m8n8k4(a0[0], a0[1], av, Lt[CF_AT(cb + g, k4 + t)]);
            dmma_m8n8k4(a1[0], a1[1], av, Lt[CF_AT(cb + 8 + g, k4 + t)]);
          }
#pragma unroll
          for (int h = 0; h < 2; h++) {
            sw[g * 20 + 2 * t + h] = mrow[r * ms + CF_B * k + cb + 2 * t + h] - a0[h];
            sw[g * 20 + 8 + 2 * t + h] = mrow[r * ms + CF_B * k + cb + 8 + 2 * t + h] - a1[h];
          }
          __syncwarp();
          double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll
          for (int k4 = 0; k4 < 16; k4 += 4) {
            const double sv = sw[g * 20 + k4 + t];
            if (k4 < 8)
              dmma_m8n8k4(c0[0], c0[1], sv, Xt[CF_XAT(b, g, k4 + t)]);
            dmma_m8n8k4(c1[0], c1[1], sv, Xt[CF_XAT(b, 8 + g, k4 + t)]);
          }
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; h++) {
            yk[r * CF_LD + cb + 2 * t + h] = c0[h];
            yk[r * CF_LD + cb + 8 + 2 * t + h] = c1[h];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      for (int idx = tid; idx < CF_RB * CF_B; idx += 256) {
        const int r = idx & (CF_RB - 1), cc = idx >> 4;
        const int row = row0 + r;
        if (cc < bs) {
          if (row < p.mrows)
            __stcg(p.Y + (size_t)(CF_B * k + cc) * p.ldy + row, yk[r * CF_LD + cc]);
          else if (row == p.mrows && p.w) {
            const double wv = yk[r * CF_LD + cc];
            __stcg(p.w + CF_B * k + cc, wv);
            wsq += wv * wv; // (one z row per launch: exactly one thread per column lands here)
          }
        }
      }
      for (int j = k + 1; j < Tp; j++) {
        if (tid == 0) {
          if (!p.prefactored)
            cf_acquire(fpan + j * Tp + k, e);
          cf_fence_async_all();
          cf_mbar_expect_tx(mb0, CF_SLOT_BYTES);
          cf_bulk_g2s(cf_saddr(Lt), CF_SLOTL(j, k), CF_SLOT_BYTES, mb0);
        }
        cf_mbar_wait(mb0, par);
        par ^= 1;
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        const double *pa = yk + g * CF_LD + t;
        const double *pb = Lt + CF_AT(8 * warp + g, t);
#pragma unroll 4
        for (int k4 = 0; k4 < CF_B; k4 += 4) {
          const double b = pb[k4 * CF_LD];
          dmma_m8n8k4(c00, c01, pa[k4], b);
          dmma_m8n8k4(c10, c11, pa[8 * CF_LD + k4], b);
        }
        double *pm = mrow + g * ms + CF_B * j + 8 * warp + 2 * t;
        pm[0] -= c00;
        pm[1] -= c01;
        pm[8 * ms] -= c10;
        pm[8 * ms + 1] -= c11;
        __syncthreads(); // Lt is overwritten by the next bulk load
      }
    }
    if (p.z && p.mrows >= row0 && p.mrows < row0 + CF_RB && p.chi2) {
      // chi2 = |L^-1 z|^2 in a fixed order (per-thread partial sums by column, then a tree), and the gate of the update
      double *red = sb; // sb is 320 doubles; 256 needed
      __syncthreads();
      red[tid] = wsq;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
          red[tid] += red[tid + o];
        __syncthreads();
      }
      if (tid == 0) {
        p.chi2[0] = red[0];
        if (p.gate_flag)
          *p.gate_flag = (p.gate_thresh < 0.0 || !(red[0] > p.gate_thresh)) ? 1 : 0;
      }
    }
  }
  // ---- epoch bookkeeping: the last CTA to finish opens the next epoch ----
  CF_TS(15)
  __syncthreads();
  if (tid == 0) {
    const int done = atomicAdd(p.ctrl + 1, 1);
    if (done == (int)gridDim.x - 1) {
      p.ctrl[1] = 0;
      __threadfence();
      atomicAdd(p.ctrl, 1);
    }
  }
}


// Factor the leading npiv columns of the n x n lower-stored matrix A in place (rows npiv..n-1 are solved along) and, when M is
// given, solve Y = M L^-T (mrows x npiv) and w = L^-1 z in the same launch.
int chol_fused(Ctx *c, double *A, int ld, int n, int npiv, double tol, const double *M, int ldm, int mrows, const double *z, int zstride,
               double *Y, int ldy, double *w, double gate_thresh, double *chi2, int *gate_flag, long long *dbg) {
  if (npiv <= 0)
    return OVP_OK;
  if (npiv > n || n > ld || (ld & 1) || ((uintptr_t)A & 15))
    return fail(c, OVP_ERR_BAD_ARGS, "chol_fused: bad sizes / alignment n=%d npiv=%d ld=%d", n, npiv, ld);
  CholFusedArgs p;
  p.A = A;
  p.ld = ld;
  p.n = n;
  p.npiv = npiv;
  p.tol = tol;
  p.T = (n + CF_B - 1) / CF_B;
  p.Tp = (npiv + CF_B - 1) / CF_B;
  if (p.T > c->cf_maxT)
    return fail(c, OVP_ERR_CAPACITY, "chol_fused: system %d exceeds the flag workspace (%d tiles)", n, c->cf_maxT);
  p.ntile = 0;
  for (int j = 0; j < p.Tp; j++)
    p.ntile += p.T - j;
  p.LinvD = c->cf_linv;
  p.xch = c->cf_xch;
  p.diag0 = c->cf_diag0;
  p.flags = c->cf_flags;
  p.ctrl = c->cf_ctrl;
  p.info = c->dflags + 1;
  p.M = M;
  p.ldm = ldm;
  p.mrows = M ? mrows : 0;
  p.z = z;
  p.zstride = zstride;
  p.gate_thresh = gate_thresh;
  p.chi2 = chi2;
  p.gate_flag = gate_flag;
  p.Y = Y;
  p.ldy = ldy;
  p.w = w;
  p.dbg = dbg;
  p.prefactored = 0;
  const int vrows = M ? (mrows + (z ? 1 : 0)) : 0;
  p.nrb = (vrows + CF_RB - 1) / CF_RB;
  p.mstride = p.Tp * CF_B + 4;
  size_t smem_tile = ((size_t)48 + 3 * CF_SLOT + CF_XSZ + 16 * CF_LD + 2 * CF_B + 8 * 96) * sizeof(double);
  size_t smem_rows = ((size_t)48 + CF_SLOT + CF_XSZ + (size_t)CF_RB * p.mstride + (size_t)CF_RB * CF_LD + 320) * sizeof(double);
  size_t smem = std::max(smem_tile, p.nrb ? smem_rows : 0);
  if (smem > 220 * 1024)
    return fail(c, OVP_ERR_CAPACITY, "chol_fused: %d columns need %zu B of shared memory", npiv, smem);
  if (!c->cf_attr_set) {
    OVP_CUDA(cudaFuncSetAttribute(chol_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    c->cf_attr_set = true;
  }
  const int grid = p.ntile + p.nrb;
  {
    // the grid must be co-resident (the spine and the tile CTAs wait on each other): one CTA per SM at this shared-memory size
    if (!c->cf_max_coresident) {
      int per_sm = 0, sms = 0;
      OVP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
      OVP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_fused_kernel, 256, 220 * 1024));
      c->cf_max_coresident = std::max(1, per_sm) * sms;
    }
    const int max_coresident = c->cf_max_coresident;
    if (p.ntile > max_coresident)
      return fail(c, OVP_ERR_CAPACITY, "chol_fused: a %d-wide system needs %d co-resident tile CTAs, the device holds %d", n, p.ntile, max_coresident);
    if (grid > max_coresident) {
      // two-launch fallback: factor first (the tile CTAs must be co-resident with the spine), then the right-hand-side row blocks in an
      // ordinary launch that reads the factor tiles from the exchange slots without waiting on flags
      double flops1 = (double)npiv * npiv * npiv / 3.0 + (double)(n - npiv) * npiv * npiv;
      CholFusedArgs p1 = p;
      p1.M = nullptr;
      p1.mrows = 0;
      p1.z = nullptr;
      p1.nrb = 0;
      p1.prefactored = 0;
      prof_begin(c, PROF_POTRF, flops1);
      void *k1[] = {(void *)&p1};
      OVP_CUDA(cudaLaunchCooperativeKernel((const void *)chol_fused_kernel, dim3(p.ntile), dim3(256), k1, smem, c->stream));
      c->launches++;
      prof_end(c);
      CholFusedArgs p2 = p;
      p2.prefactored = 1;
      prof_begin(c, PROF_POTRF, (double)vrows * npiv * npiv);
      chol_fused_kernel<<<p.nrb, 256, smem, c->stream>>>(p2);
      c->launches++;
      prof_end(c);
      return OVP_OK;
    }
  }
  double flops = (double)npiv * npiv * npiv / 3.0 + (double)(n - npiv) * npiv * npiv + (double)vrows * npiv * npiv;
  prof_begin(c, PROF_POTRF, flops);
  // cooperative launch: the spine waits on tile CTAs that wait on the spine, so the whole grid must be co-resident
  void *kargs[] = {(void *)&p};
  OVP_CUDA(cudaLaunchCooperativeKernel((const void *)chol_fused_kernel, dim3(grid), dim3(256), kargs, smem, c->stream));
  c->launches++;
  prof_end(c);
  return OVP_OK;
}

} // namespace ovp
