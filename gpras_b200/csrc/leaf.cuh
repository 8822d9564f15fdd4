// Single-CTA factorisation of one 128 x 128 diagonal block: L = chol(A_jj) and W = L^-1.
//
// The block lives in shared memory as S[128][132] doubles (row pitch 132 == 4 mod 16 makes every
// DMMA fragment LDS.64 conflict free).  L occupies the lower triangle (incl. diagonal); W is built
// transposed in the strictly-upper triangle (W[i][k] at S[k][i], i > k) with its diagonal 1/L_ii in
// dvec[], so L, W and the temporaries of the inverse all fit in one 135 KB tile.
//
//   1. potrf, right-looking over 16 panels of 8 columns with one panel of look-ahead: warps 0..3 factor panel
//      J+1 (one row per lane; the 8x8 diagonal tile is factored redundantly in every lane's registers, so only
//      one rsqrt per column sits on the critical path and no shuffle or barrier) WHILE warps 4..7 apply
//      panel J's rank-8 update to the rest of the trailing triangle with DMMA.8x8x4 on 8 x 8 tiles; only
//      the update of block column J+1 itself (all warps) sits between two panel factorisations.
//   2. inverse by recursive doubling: 8 x 8 diagonal blocks by forward substitution (one thread per
//      column), then for b = 8, 16, 32, 64 every pair of adjacent blocks fills its off-diagonal block
//      W21 = -W22 (L21 W11) with two DMMA products (T = L21 W11 is parked in the pair's mirrored
//      upper block, which W21^T then overwrites).  Each warp advances its b/8 output tiles together so the
//      dependent DMMA chains of different tiles interleave.
//   3. L (upper zeroed) goes back over A_jj; W (upper zeroed) goes to the W buffer; sum(log L_ii) to
//      logdet_part[jb]; a non-positive pivot records info = global column + 1 (LAPACK convention).
#pragma once
#include "common.cuh"

namespace gpras {

constexpr int LEAF_N = 128, LEAF_LD = 132, LEAF_THREADS = 256;
constexpr int LEAF_SMEM_BYTES = (LEAF_N * LEAF_LD + LEAF_N) * (int)sizeof(double);

// 8x8 tiles of the 16 x 16 lower block triangle in column-major order: tiles of block column c start at
// LEAF_COL_OFF[c]; entry t is (row LEAF_TILE_I[t], column LEAF_TILE_C[t]).
struct LeafTileTable {
  unsigned char I[136], C[136];
  unsigned char off[17];
};
__host__ __device__ constexpr LeafTileTable make_leaf_table() {
  LeafTileTable t{};
  int n = 0;
  for (int c = 0; c < 16; c++) {
    t.off[c] = (unsigned char)n;
    for (int i = c; i < 16; i++) {
      t.I[n] = (unsigned char)i;
      t.C[n] = (unsigned char)c;
      n++;
    }
  }
  t.off[16] = (unsigned char)n;
  return t;
}
static __constant__ LeafTileTable c_leaf_tiles = make_leaf_table();

#ifdef LEAF_TIMING
static __device__ long long* g_leaf_timing;
#define LT_DECL long long lt_acc[16] = {0}; const unsigned lt_sa = (unsigned)__cvta_generic_to_shared(smem); long long lt_prev = clock64(), lt_start = lt_prev
// the shared load + dependent predicate makes the clock read wait for a preceding barrier's RELEASE
#define LT_MARK(i) do { unsigned lt_d; long long lt_now; \
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(lt_d) : "r"(lt_sa)); \
    asm volatile("{ .reg .pred p; setp.eq.u32 p, %1, 0x7fc12345; @p trap; mov.u64 %0, %%clock64; }" : "=l"(lt_now) : "r"(lt_d)); \
    lt_acc[i] += lt_now - lt_prev; lt_prev = lt_now; } while (0)
#define LT_DUMP do { __syncthreads(); LT_MARK(12); if (tid == 0 || tid == 128) { \
    for (int i = 0; i < 13; i++) if ((tid == 0) != (i == 4 || i == 5)) g_leaf_timing[i] = lt_acc[i]; \
    if (tid == 0) g_leaf_timing[13] = clock64() - lt_start; } } while (0)
#else
#define LT_DECL
#define LT_MARK(i)
#define LT_DUMP
#endif

__device__ __forceinline__ double leaf_getW(const double* S, const double* dvec, int i, int k) {
  const double s = S[k * LEAF_LD + i];
  const double dg = dvec[i];
  return i == k ? dg : (i > k ? s : 0.0);
}

// rank-8 update, with panel columns [j0, j0+8), of up to NT 8x8 tiles listed in c_leaf_tiles[t0 + n*stride]
// (n < NT, t < t_end); the NT dependent LDS -> DMMA -> DMMA -> STS chains are interleaved.
template <int NT>
__device__ __forceinline__ void leaf_update_tiles(double* S, int t0, int stride, int t_end, int j0, int g, int q) {
  double2 cv[NT];
  double a0[NT], a1[NT], b0[NT], b1[NT];
  double* cp[NT];
#pragma unroll
  for (int n = 0; n < NT; n++) {
    const int t = t0 + n * stride;
    const bool on = t < t_end;  // warp-uniform
    const int I = on ? c_leaf_tiles.I[t] : 15, Cc = on ? c_leaf_tiles.C[t] : 15;
    cp[n] = S + (8 * I + g) * LEAF_LD + 8 * Cc + 2 * q;
    cv[n] = *reinterpret_cast<double2*>(cp[n]);
    a0[n] = S[(8 * I + g) * LEAF_LD + j0 + q], a1[n] = S[(8 * I + g) * LEAF_LD + j0 + 4 + q];
    b0[n] = S[(8 * Cc + g) * LEAF_LD + j0 + q], b1[n] = S[(8 * Cc + g) * LEAF_LD + j0 + 4 + q];
  }
#pragma unroll
  for (int n = 0; n < NT; n++) {
    cv[n].x = -cv[n].x, cv[n].y = -cv[n].y;
    dmma(cv[n].x, cv[n].y, a0[n], b0[n]);
  }
#pragma unroll
  for (int n = 0; n < NT; n++) dmma(cv[n].x, cv[n].y, a1[n], b1[n]);
#pragma unroll
  for (int n = 0; n < NT; n++) {
    if (t0 + n * stride < t_end) *reinterpret_cast<double2*>(cp[n]) = make_double2(-cv[n].x, -cv[n].y);
  }
}

__device__ __forceinline__ double leaf_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // Newton from the ~2^-22 seed, three dependent operations per step: y += y * (1/2 - (x/2) y^2).
  // Two steps reach 2^-43 and then the rounding floor of the arithmetic itself.
  const double hx = 0.5 * x;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double t = hx * y;
    const double u = fma(-t, y, 0.5);
    y = fma(y, u, y);
  }
  return y;
}

// warps 0..3: factor the 8-column panel starting at column j0 (rows j0..127).  Every lane holds the 8x8 diagonal
// tile and factors it redundantly in registers (no shuffles, no barrier on the critical path) while carrying its
// own row j0 + 32*warp + lane of the panel along.
__device__ __forceinline__ void leaf_factor_panel(double* S, double* dvec, int j0, int warp, int lane, int* info,
                                                  int col_base) {
  if (j0 + 32 * warp >= LEAF_N) return;  // warp-uniform: no rows left for this warp
  const int r = j0 + 32 * warp + lane;
  const bool live = r < LEAF_N;
  double D[8][8], v[8];
#pragma unroll
  for (int rr = 0; rr < 8; rr++)
#pragma unroll
    for (int c = 0; c <= rr; c++) D[rr][c] = S[(j0 + rr) * LEAF_LD + j0 + c];
#pragma unroll
  for (int c = 0; c < 8; c++) v[c] = live ? S[r * LEAF_LD + j0 + c] : 0.0;
  // Rows j0 .. j0+7 are the diagonal tile every lane has just read AND the panel rows warp 0 writes back below: the
  // participating warps meet on a named barrier first, or a warp that was held up (the SM is shared with other CTAs
  // when several evaluations are in flight) would read an already factored tile.
  {
    const int nw = (LEAF_N - j0 + 31) / 32;  // warps that own rows of this panel (all of them reach this point)
    asm volatile("bar.sync 1, %0;" ::"r"(32 * nw) : "memory");
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    double dpiv = D[c][c];
    if (!(dpiv > 0.0)) {  // identical on every lane
      if (warp == 0 && lane == 0) atomicCAS(info, 0, col_base + j0 + c + 1);
      dpiv = 1.0;
    }
    const double rs = leaf_rsqrt(dpiv);
    if (warp == 0 && lane == 0) dvec[j0 + c] = rs;
#pragma unroll
    for (int rr = c + 1; rr < 8; rr++) D[rr][c] *= rs;
#pragma unroll
    for (int c2 = c + 1; c2 < 8; c2++)
#pragma unroll
      for (int rr = c2; rr < 8; rr++) D[rr][c2] -= D[rr][c] * D[c2][c];
    v[c] *= rs;
#pragma unroll
    for (int c2 = c + 1; c2 < 8; c2++) v[c2] -= v[c] * D[c2][c];
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < 8; c++) S[r * LEAF_LD + j0 + c] = (r >= j0 + c) ? v[c] : 0.0;
  }
}

// One recursive-doubling level of the inverse at block size B.  A pair of adjacent B-blocks has TB x TB output
// tiles (TB = B/8) and is served by TB warps; warp u of the pair takes tiles (n, (u + n) mod TB), n < TB, so every
// warp sees each row index and each column index once and the triangular k ranges balance.
template <int B>
__device__ __forceinline__ void leaf_inverse_level(double* S, const double* dvec, int warp, int g, int q,
                                                   double* __restrict__ Wg = nullptr, long ldw = 0) {
  constexpr int TB = B / 8;
  const int p = warp / TB, u = warp % TB, r0 = 2 * B * p;
  int i0[TB], jj0[TB];
#pragma unroll
  for (int n = 0; n < TB; n++) {
    i0[n] = 8 * n;
    jj0[n] = 8 * ((u + n) % TB);
  }
  double c0[TB], c1[TB];
  // GEMM1: T[i][j] = sum_{k >= j} L21[i][k] W11[k][j]; stored transposed in the pair's mirrored upper block
#pragma unroll
  for (int n = 0; n < TB; n++) c0[n] = c1[n] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < B; k0 += 4) {
#pragma unroll
    for (int n = 0; n < TB; n++) {
      if (k0 >= jj0[n]) {
        const double a = S[(r0 + B + i0[n] + g) * LEAF_LD + r0 + k0 + q];
        const double bb = leaf_getW(S, dvec, r0 + k0 + q, r0 + jj0[n] + g);
        dmma(c0[n], c1[n], a, bb);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < TB; n++) {
    S[(r0 + jj0[n] + 2 * q) * LEAF_LD + r0 + B + i0[n] + g] = c0[n];
    S[(r0 + jj0[n] + 2 * q + 1) * LEAF_LD + r0 + B + i0[n] + g] = c1[n];
  }
  __syncthreads();
  // GEMM2: W21[i][j] = -sum_{k <= i} W22[i][k] T[k][j]; overwrites T in place after a barrier
#pragma unroll
  for (int n = 0; n < TB; n++) c0[n] = c1[n] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < B; k0 += 4) {
#pragma unroll
    for (int n = 0; n < TB; n++) {
      if (k0 < i0[n] + 8) {
        const double a = leaf_getW(S, dvec, r0 + B + i0[n] + g, r0 + B + k0 + q);
        const double bb = S[(r0 + jj0[n] + g) * LEAF_LD + r0 + B + k0 + q];
        dmma(c0[n], c1[n], a, bb);
      }
    }
  }
  if (B == 64) {
    // last level: W21 (rows 64.., columns ..63) goes straight to global memory, nothing reads it from S any more
#pragma unroll
    for (int n = 0; n < TB; n++)
      *reinterpret_cast<double2*>(Wg + (long)(B + i0[n] + g) * ldw + jj0[n] + 2 * q) = make_double2(-c0[n], -c1[n]);
    return;
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < TB; n++) {
    S[(r0 + jj0[n] + 2 * q) * LEAF_LD + r0 + B + i0[n] + g] = -c0[n];
    S[(r0 + jj0[n] + 2 * q + 1) * LEAF_LD + r0 + B + i0[n] + g] = -c1[n];
  }
  __syncthreads();
}

// Body shared by the three leaf kernels.  POTRF: factor the block (else the block already holds L and only 1 / L_ii is
// formed); INVERSE: build W = L^-1 and store it.  `In` points at the 128 x 128 block to read (pitch ldin), `A` at where L
// goes (pitch lda; only written when POTRF), `W` at where the inverse goes.
template <bool POTRF, bool INVERSE>
__device__ __forceinline__ void leaf_body(const double* In, long ldin, double* A, long lda, double* __restrict__ W, long ldw,
                                          double* __restrict__ logdet_part, int* __restrict__ info, int jb) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double logred[4];
  double* S = smem;
  double* dvec = smem + LEAF_N * LEAF_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;

  LT_DECL;
  // ---- load: every 16-byte chunk that touches the lower triangle, all in flight at once (cp.async).  The strictly
  //      upper part of S is never read before it is overwritten (panel rows are masked on write-back, tile updates
  //      only propagate within an entry), so it is left as it is ----
  for (int e = tid; e < LEAF_N * (LEAF_N / 2); e += LEAF_THREADS) {
    const int r = e >> 6, c2 = (e & 63) * 2;
    if (c2 <= r) cp_async16(S + r * LEAF_LD + c2, In + (long)r * ldin + c2);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  LT_MARK(0);

  if (POTRF) {
    // ---- 1. blocked Cholesky, right-looking over 16 panels of 8 columns with one panel of look-ahead ----
    if (warp < 4) leaf_factor_panel(S, dvec, 0, warp, lane, info, jb * LEAF_N);
    __syncthreads();
    LT_MARK(1);
    for (int J = 0; J < 15; J++) {
      const int j0 = 8 * J;
      // phase A: block column J+1 (<= 15 tiles, <= 2 per warp) gets panel J's update
      leaf_update_tiles<2>(S, c_leaf_tiles.off[J + 1] + warp, 8, c_leaf_tiles.off[J + 2], j0, g, q);
      __syncthreads();
      LT_MARK(2);
      if (warp < 4) {
        // phase B, warps 0..3: factor panel J+1
        leaf_factor_panel(S, dvec, j0 + 8, warp, lane, info, jb * LEAF_N);
        LT_MARK(3);
      } else {
        // phase B, warps 4..7: the rest of the trailing triangle (block columns >= J+2) gets panel J's update
        const int te = c_leaf_tiles.off[16];
        for (int t = c_leaf_tiles.off[J + 2] + (warp - 4); t < te; t += 16) leaf_update_tiles<4>(S, t, 4, te, j0, g, q);
        LT_MARK(4);
      }
      __syncthreads();
      LT_MARK(5);
    }
    // sum(log L_ii) = -sum(log dvec): 128 logs in parallel, fixed-shape reduction
    if (tid < 128) {
      double lg = -log(dvec[tid]);
      lg = warp_sum(lg);
      if (lane == 0) logred[warp] = lg;
    }
    __syncthreads();
    if (tid == 0) logdet_part[jb] = (logred[0] + logred[1]) + (logred[2] + logred[3]);
    LT_MARK(6);
    // ---- write L back: the chunks that touch the lower triangle (the strictly upper part of L's diagonal block is
    //      unspecified: no kernel reads it) ----
#pragma unroll 8
    for (int e = tid; e < LEAF_N * (LEAF_N / 2); e += LEAF_THREADS) {
      const int r = e >> 6, c2 = (e & 63) * 2;
      if (c2 <= r) {
        double2 v = *reinterpret_cast<const double2*>(S + r * LEAF_LD + c2);
        if (c2 + 1 > r) v.y = 0.0;
        *reinterpret_cast<double2*>(A + (long)r * lda + c2) = v;
      }
    }
  } else {
    if (tid < 128) dvec[tid] = 1.0 / S[tid * LEAF_LD + tid];
    __syncthreads();
  }
  if (!INVERSE) {
    LT_DUMP;
    return;
  }

  // ---- 2. inverse: 8x8 diagonal blocks (thread = one column of one block) ----
  if (tid < 128) {
    const int blk = tid >> 3, c = tid & 7, o = 8 * blk;
    double w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k < i && k >= c) s -= S[(o + i) * LEAF_LD + o + k] * w[k];
      w[i] = (i >= c) ? s * dvec[o + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (i > c) S[(o + c) * LEAF_LD + o + i] = w[i];  // W[o+i][o+c] stored transposed
  }
  __syncthreads();
  LT_MARK(7);
  leaf_inverse_level<8>(S, dvec, warp, g, q);
  LT_MARK(8);
  leaf_inverse_level<16>(S, dvec, warp, g, q);
  LT_MARK(9);
  leaf_inverse_level<32>(S, dvec, warp, g, q);
  LT_MARK(10);
  leaf_inverse_level<64>(S, dvec, warp, g, q, W, ldw);
  LT_MARK(11);

  // ---- 3. W: 8 x 4 element blocks of the two diagonal 64 x 64 triangles, transposed conflict-free reads; blocks above the
  //         diagonal are never written (the W buffer is zero-initialised once and the engine relies on those zeros), the
  //         lower-left 64 x 64 block is already in global memory ----
#pragma unroll 4
  for (int blk = warp; blk < 16 * 32; blk += 8) {  // 16 row groups of 8 x 32 column groups of 4
    const int rb = 8 * (blk >> 5), cb = 4 * (blk & 31);
    if (cb <= rb + 7 && !(rb >= 64 && cb < 64)) {
      const int r = rb + g, c = cb + q;
      W[(long)r * ldw + c] = leaf_getW(S, dvec, r, c);
    }
  }
  LT_DUMP;
}

// Factor + invert one diagonal block (single-block matrices: M <= 128 of the sparse model, and the microbenchmark).
static __global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_inv_kernel(const double* In, long ldin, double* A, long lda, double* __restrict__ W,
                      long ldw, double* __restrict__ logdet_part, int* __restrict__ info, int jb, long bs = 0) {
  // batch of independent single-block matrices (sparse models trained together): blockIdx.y = model, bs doubles apart
  if (bs) {
    const bool same = In == A;
    A += blockIdx.y * bs, W += blockIdx.y * bs, logdet_part += blockIdx.y * bs, info += blockIdx.y * bs * 2;
    In = same ? A : In + blockIdx.y * bs;
  }
  // In points at the 128 x 128 block to factor (pitch ldin): the diagonal block of A itself, or a scratch block;
  // L goes to the diagonal block jb of A, W = L^-1 to that of W.
  if (In == A) In += (long)jb * LEAF_N * lda + (long)jb * LEAF_N;
  A += (long)jb * LEAF_N * lda + (long)jb * LEAF_N;
  W += (long)jb * LEAF_N * ldw + (long)jb * LEAF_N;
  leaf_body<true, true>(In, ldin, A, lda, W, ldw, logdet_part, info, jb);
}

// The Cholesky chain's leaf: factor diagonal block jb of A in place, nothing else (the inverse of the block is not
// on the factorisation's critical path any more: the panel below it is a triangular SOLVE, trsm_panel_kernel).
static __global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_kernel(double* A, long lda, double* __restrict__ logdet_part, int* __restrict__ info, int jb) {
  A += (long)jb * LEAF_N * lda + (long)jb * LEAF_N;
  leaf_body<true, false>(A, lda, A, lda, nullptr, 0, logdet_part, info, jb);
}

// W_jj = L_jj^-1 for ALL diagonal blocks in one launch (blockIdx.x = block), after the factorisation.
static __global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_inv_kernel(const double* A, long lda, double* __restrict__ W, long ldw) {
  const int jb = blockIdx.x;
  A += (long)jb * LEAF_N * lda + (long)jb * LEAF_N;
  W += (long)jb * LEAF_N * ldw + (long)jb * LEAF_N;
  leaf_body<false, true>(A, lda, nullptr, 0, W, ldw, nullptr, nullptr, jb);
}

// ---- triangular-solve panel --------------------------------------------------------------------------------------
// X = B L^-T in place, for the rows of A below diagonal block jb (B = A[rows, block column jb], L = the factored
// diagonal block).  One warp owns 8 rows and keeps them in registers as sixteen 8 x 8 DMMA accumulator tiles; the lower
// 8 x 8 blocks of L, the diagonal ones inverted, sit in shared memory.  Right-looking over the
// 8-column blocks: X_b = B_b inv(L_bb)^T (two DMMAs), then B_c -= X_b L_cb^T for every c > b (independent
// accumulators, two DMMAs each).  The accumulator -> A-operand re-layout is four quad shuffles; there is no barrier
// and no shared-memory traffic for X inside the loop, so a row tile finishes in ~16 x (solve + first update) DMMA
// latencies while the other updates fill the pipe.
constexpr int TRSM_WARPS = 4, TRSM_ROWS = 8 * TRSM_WARPS, TRSM_THREADS = 32 * TRSM_WARPS;
// L_jj sits in shared memory as its 136 lower 8 x 8 blocks, block (c, b) at (c (c + 1) / 2 + b) * TRSM_BLK with row pitch 12
// (== 12 mod 16: conflict-free fragment reads); the diagonal slots end up holding the INVERSE of their block.  104 KB, so a
// panel CTA shares an SM with a trailing-update CTA instead of evicting it (with a 135 KB square tile it could not).
constexpr int TRSM_DLD = 12, TRSM_BLK = 8 * TRSM_DLD + 2;  // (+2: the diagonal slots of blocks w, w+4, w+8, w+12 fall in different banks)
constexpr int TRSM_SMEM_BYTES = 136 * TRSM_BLK * (int)sizeof(double);
__device__ __forceinline__ int trsm_blk(int c, int b) { return (c * (c + 1) / 2 + b) * TRSM_BLK; }

// C-fragment (lane (g, q) holds columns 2q, 2q+1 of row g) -> the two A-fragments of the k = 0..3 / 4..7 halves
// (lane (g, q) holds column q / column 4 + q of row g)
__device__ __forceinline__ void trsm_c_to_a(double cx, double cy, int lane, int q, double& a_lo, double& a_hi) {
  const int base = lane & ~3, src_lo = base | (q >> 1), src_hi = base | (2 + (q >> 1));
  const double lx = __shfl_sync(0xffffffffu, cx, src_lo), ly = __shfl_sync(0xffffffffu, cy, src_lo);
  const double hx = __shfl_sync(0xffffffffu, cx, src_hi), hy = __shfl_sync(0xffffffffu, cy, src_hi);
  a_lo = (q & 1) ? ly : lx;
  a_hi = (q & 1) ? hy : hx;
}

static __global__ void __launch_bounds__(TRSM_THREADS, 2)
trsm_panel_kernel(double* A, long lda, int jb) {
  extern __shared__ __align__(16) double smem[];
  double* S = smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const double* Ljj = A + (long)jb * LEAF_N * lda + (long)jb * LEAF_N;
  for (int e = tid; e < LEAF_N * (LEAF_N / 2); e += TRSM_THREADS) {
    const int r = e >> 6, c2 = (e & 63) * 2;
    // (the chunk that straddles the diagonal brings one element of the upper triangle along: it lands in the diagonal
    //  block's slot and is never read)
    if (c2 <= r) cp_async16(S + trsm_blk(r >> 3, c2 >> 3) + (r & 7) * TRSM_DLD + (c2 & 7), Ljj + (long)r * lda + c2);
  }
  cp_async_commit();
  // this warp's 8 rows: sixteen accumulator tiles, loaded while L streams in
  double* X = A + ((long)(jb + 1) * LEAF_N + (long)blockIdx.x * TRSM_ROWS + 8 * warp + g) * lda + (long)jb * LEAF_N + 2 * q;
  double2 acc[16];
#pragma unroll
  for (int c = 0; c < 16; c++) acc[c] = *reinterpret_cast<const double2*>(X + 8 * c);
  cp_async_wait<0>();
  __syncthreads();
  {  // inverses of the 8 x 8 diagonal blocks, in place: thread = one column of one block (forward substitution)
    const int blk = 4 * ((tid >> 3) & 3) + warp, c = tid & 7;  // a warp's four blocks sit in four different banks
    double* Dg = S + trsm_blk(blk, blk);
    double w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k < i && k >= c) s -= Dg[i * TRSM_DLD + k] * w[k];
      w[i] = (i >= c) ? s / Dg[i * TRSM_DLD + i] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; i++) Dg[i * TRSM_DLD + c] = w[i];
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 16; b++) {
    double a_lo, a_hi;
    trsm_c_to_a(acc[b].x, acc[b].y, lane, q, a_lo, a_hi);
    // X_b[r][n] = sum_k B_b[r][k] inv(L_bb)[n][k]:  DMMA operand B[k][n] = inv(L_bb)[n][k], lane (g, q) reads n = g, k = q
    const double* Dg = S + trsm_blk(b, b) + g * TRSM_DLD + q;
    double x0 = 0.0, x1 = 0.0;
    dmma(x0, x1, a_lo, Dg[0]);
    dmma(x0, x1, a_hi, Dg[4]);
    *reinterpret_cast<double2*>(X + 8 * b) = make_double2(x0, x1);
    if (b < 15) {
      trsm_c_to_a(-x0, -x1, lane, q, a_lo, a_hi);
#pragma unroll
      for (int c = b + 1; c < 16; c++) {
        // B_c -= X_b L_cb^T:  operand B[k][n] = L[8c + n][8b + k]
        const double* Lg = S + trsm_blk(c, b) + g * TRSM_DLD + q;
        dmma(acc[c].x, acc[c].y, a_lo, Lg[0]);
        dmma(acc[c].x, acc[c].y, a_hi, Lg[4]);
      }
    }
  }
}

}  // namespace gpras
