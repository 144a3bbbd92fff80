// hank_dense.cu — the one dense contraction of the path: J̅⁻¹ for the Newton driver's preconditioner solve
// (NewtonRaphson.jl:97: `gmres!(R, J̅, rhs)`, replaced by R = J̅⁻¹·rhs in the LU modes of hank_newton.cu).
//
// Hand-written FP64 inverse by BLOCKED GAUSS-JORDAN with partial (row) pivoting, in place, column-major.
// Round 1 used cuSOLVER getrf + getrs on the identity (6.0 ms at n = 1196, ~2 % of the FP64 peak: three library
// phases, each a chain of small latency-bound kernels).  Gauss-Jordan produces the inverse directly in ONE pass of
// n/B block steps, each step two kernels queued back to back on the context's stream (no host round trip: pivots,
// permutations and the singularity flag stay on the device):
//
//   k_gj_panel  (one CTA, the n x B column panel in shared memory):
//       partial-pivoted LU of the rows at and below the block (two barriers per column: the pivot reduction and the
//       pivot-row broadcast; every thread only ever writes the rows it owns), then per row the two small triangular
//       solves that turn the panel into the block column of the Gauss-Jordan transform
//           Gcol = [ -P_top W ;  W ;  -P_bot W ],   W = P_mid^-1 = U^-1 L_mid^-1,   -P_bot W = -L_bot L_mid^-1,
//       which replaces the panel in place, plus the net effect of the block's B row interchanges as a list of at most
//       2B (position <- source row) moves for the update kernel.
//   k_gj_update (148+ CTAs, 8 columns each): applies the interchanges to its columns through shared memory (two L2
//       round trips instead of B dependent swaps), then the rank-B update  C <- C + Gcol·C_mid  (block rows:
//       C_mid <- W·C_mid) with the panel columns streamed from L2: 2·n·B FMAs per column.
//
// After the last block the row interchanges are undone as ONE column permutation (k_gj_rowperm + k_gj_unscramble).
// Work: 2 n^3 flops like getrf + getri; L2 traffic (n/B)·2·8·n^2 bytes (1.7 GB at n = 1196, B = 16); the critical
// path is the n pivot columns of the panels (two barriers each).  FP64 DMMA (mma.sync.m8n8k4.f64) would fit the
// rank-B update, but at B = 16 that update is L2-bound (16 flops per 8-byte element), not pipe-bound, and the B200
// FP64 tensor rate equals its FP64 FMA rate, so plain DFMA is used.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "hank_ctx.h"
#include "hank_tangent_tma.cuh"   // mbarrier + cp.async.bulk helpers
#include "../../include/hankb200.h"

namespace hank {

#define CK(call)                                                   \
  do {                                                             \
    int rc__ = hank::cuda_check(c, (call), #call);                 \
    if (rc__) return rc__;                                         \
  } while (0)

#ifndef HANK_GJ_DEBUG
#define HANK_GJ_DEBUG 0
#endif
__device__ long long gj_dbg[8];   // HANK_GJ_DEBUG=1 builds: clock64 stamps of the last panel step's phases

constexpr int GJ_NT = 480;       // panel kernel threads that own rows (15 warps; the 16th does the bookkeeping)
constexpr int GJ_PT = GJ_NT + 32;
constexpr int GJ_UT = 256;       // update kernel threads
constexpr int GJ_MAXMOVES = 32;  // 2 * largest block width


// shared -> global bulk copy (TMA), committed and drained by the issuing thread
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst)),
               "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// 1/x to ~1 ulp without the IEEE division sequence: hardware approximation + two Newton steps (pivots are the
// largest entries of their columns: no denormals to care about; 0 gives inf like the division)
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
// Warp argmax of |a| with first-index tie-break in two hardware reductions: the key is the magnitude's bit pattern with
// its low 13 mantissa bits replaced by (8191 - row), so candidates equal to 2^-39 relative count as ties.
__device__ __forceinline__ unsigned long long pivot_key(double a, int r) {
  return ((unsigned long long)__double_as_longlong(fabs(a)) & ~0x1FFFull) | (unsigned long long)(8191 - r);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
  const unsigned hi = (unsigned)(k >> 32);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned lo = hi == mh ? (unsigned)k : 0u;
  const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
  return ((unsigned long long)mh << 32) | ml;
}

template <int B> constexpr size_t gj_panel_smem(int n) {
  return ((size_t)B * n + (size_t)B * B + B + 2 * (GJ_NT / 32) + 2) * sizeof(double) + (size_t)(B + 4) * sizeof(int) + 128;
}

// One block step on the column panel A[:, k0 : k0+bw).  ipiv[k] = pivot row of column k (LAPACK convention, 0-based).
// moves: [2*GJ_MAXMOVES] ints for this step: pos[i], src[i] (i < GJ_MAXMOVES; pos < 0: unused): after the block's
// interchanges position pos[i] holds what row src[i] held before them; entries i < bw are the block rows k0 + i.
// The panel is one contiguous piece of column-major storage: it comes in and goes out as ONE TMA bulk copy.
// The column loop is unrolled over j so that "columns left of / right of the pivot" are compile-time ranges: a row
// costs one LDS + DFMA + STS per remaining column, and only the two owners of the interchanged rows touch whole rows.
// Warp GJ_NT/32 owns no rows: it turns the pivots into the move list while the others run the row solves.
template <int B>
__global__ void __launch_bounds__(GJ_PT, 1)
k_gj_panel(double* __restrict__ A, int n, int k0, int bw, int* __restrict__ ipiv, int* __restrict__ moves,
           int* __restrict__ info) {
  static_assert(2 * B <= GJ_MAXMOVES, "move list too short");
  constexpr int NW = GJ_NT / 32;
  static_assert(NW + 1 == 16, "the final pivot reduction assumes 16 warps in all");
  extern __shared__ __align__(128) unsigned char smem_gj[];
  double* pan = reinterpret_cast<double*>(smem_gj);          // [bw][n] column-major panel
  double* LU = pan + (size_t)B * n;                            // [B][B] row-major copy of the factored block rows
  double* dinv = LU + B * B;                                   // 1 / U[c][c]
  unsigned long long* red = reinterpret_cast<unsigned long long*>(dinv + B);   // [NW] per-warp pivot keys
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * NW);
  int* pivs = reinterpret_cast<int*>(bar + 1);                // [B] pivot rows of this block
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool rows = tid < GJ_NT;                               // (the bookkeeping warp takes part in the barriers only)
  double* Ap = A + (size_t)k0 * n;                             // the panel is contiguous in column-major storage
  const int tot = n * bw;
  const bool bulk = ((size_t)tot * 8) % 16 == 0 && (reinterpret_cast<uintptr_t>(Ap) & 15) == 0;
  if (HANK_GJ_DEBUG && tid == 0) gj_dbg[0] = clock64();
  if (bulk) {
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, (uint32_t)tot * 8u);
      bulk_g2s(pan, Ap, (uint32_t)tot * 8u, bar);
      mbar_wait(bar, 0);
    }
  } else {
#pragma unroll 8
    for (int i = tid; i < tot; i += GJ_PT) pan[i] = Ap[i];
  }
  __syncthreads();

#pragma unroll
  for (int j = 0; j < B; ++j) {
    if (j < bw) {
      const int k = k0 + j;
      double* colj = pan + (size_t)j * n;
      // ---- pivot: largest |a| among rows >= k, first one on ties (idamax)
      unsigned long long key = 0;
      if (rows)
        for (int r = tid; r < n; r += GJ_NT)
          if (r >= k) { const unsigned long long q = pivot_key(colj[r], r); key = q > key ? q : key; }
      key = warp_max_u64(key);
      if (lane == 0) red[warp] = key;   // (the bookkeeping warp contributes 0)
      __syncthreads();
      key = warp_max_u64(red[lane & 15]);
      int p = 8191 - (int)(key & 0x1FFFull);
      if (p >= n || p < k) p = k;   // a column of NaNs / zeros keeps the diagonal
      // ---- the pivot row right of the pivot for everyone; the two whole rows for the owners of rows k and p
      double prow[B];
#pragma unroll
      for (int c = j; c < B; ++c) prow[c] = c < bw ? pan[(size_t)c * n + p] : 0.0;
      const bool own_k = (k % GJ_NT) == tid && p != k, own_p = (p % GJ_NT) == tid && p != k;
      double lowp[B], full[B];   // owner of row k: row p left of the pivot; owner of row p: the whole row k
      if (own_k) {
#pragma unroll
        for (int c = 0; c < j; ++c) lowp[c] = pan[(size_t)c * n + p];
      }
      if (own_p) {
#pragma unroll
        for (int c = 0; c < B; ++c) full[c] = c < bw ? pan[(size_t)c * n + k] : 0.0;
      }
      __syncthreads();
      if (tid == GJ_NT) {
        ipiv[k] = p; pivs[j] = p;
        if (prow[j] == 0.0) atomicCAS(info, 0, k + 1);   // exactly singular (getrf's info > 0)
      }
      const double inv = fast_rcp(prow[j]);
      if (rows)
        for (int r = tid; r < n; r += GJ_NT) {
          if (r <= k || r == p) continue;   // rows above, the block's finished U rows, and the two interchanged rows
          const double lmul = colj[r] * inv;
          colj[r] = lmul;
#pragma unroll
          for (int c = j + 1; c < B; ++c)
            if (c < bw) pan[(size_t)c * n + r] = fma(-lmul, prow[c], pan[(size_t)c * n + r]);
        }
      if (own_k) {
#pragma unroll
        for (int c = 0; c < B; ++c) if (c < bw) pan[(size_t)c * n + k] = c < j ? lowp[c] : prow[c];
      }
      if (own_p) {   // row p receives the old row k, eliminated like every other row below the pivot
        const double lmul = full[j] * inv;
#pragma unroll
        for (int c = 0; c < B; ++c) {
          if (c >= bw) continue;
          pan[(size_t)c * n + p] = c < j ? full[c] : (c == j ? lmul : fma(-lmul, prow[c], full[c]));
        }
      }
    }
  }
  __syncthreads();
  if (HANK_GJ_DEBUG && tid == 0) gj_dbg[2] = clock64();
  // ---- the factored block rows (unit-lower L_mid below the diagonal, U on and above it) and 1 / diag(U)
  for (int i = tid; i < B * B; i += GJ_PT) {
    const int row = i / B, col = i - row * B;
    const double v = (row < bw && col < bw) ? pan[(size_t)col * n + k0 + row] : (row == col ? 1.0 : 0.0);
    LU[i] = v;
    if (row == col) dinv[row] = 1.0 / v;
  }
  __syncthreads();
  if (HANK_GJ_DEBUG && tid == 0) gj_dbg[3] = clock64();
  if (!rows) {
    // ---- net effect of the block's interchanges: positions k0..k0+bw-1 and the pivot rows outside the block
    int pos = -1;
    if (lane < bw) pos = k0 + lane;
    else if (lane < 2 * bw) {
      const int pj = pivs[lane - bw];
      bool dup = pj < k0 + bw;   // pivots are >= k0; inside the block they are block rows already listed
      for (int q = 0; q < lane - bw; ++q) dup |= pivs[q] == pj;
      pos = dup ? -1 : pj;
    }
    int cur = pos;
    for (int j = 0; j < bw; ++j) {
      const int a = k0 + j, b = pivs[j];
      if (a == b) continue;
      const unsigned ma = __ballot_sync(0xffffffffu, pos == a), mb = __ballot_sync(0xffffffffu, pos == b);
      const int la = __ffs(ma) - 1, lb = __ffs(mb) - 1;
      const int ca = __shfl_sync(0xffffffffu, cur, la), cb = __shfl_sync(0xffffffffu, cur, lb);
      if (lane == la) cur = cb;
      if (lane == lb) cur = ca;
    }
    moves[lane] = pos;
    moves[GJ_MAXMOVES + lane] = cur;
  } else {
    // ---- every row: y = p U^-1 (rows above the block: p = the row; block rows: p = e_row; rows below hold y = L_bot
    // already), x = y L_mid^-1, Gcol row = x for the block rows, -x elsewhere.  Both solves in axpy form (a finished
    // component is subtracted from all later ones at once: independent FMAs instead of one long dot-product chain).
    constexpr int RCH = B >= 16 ? 2 : 3;
    for (int r0 = tid; r0 < n; r0 += RCH * GJ_NT) {
      double y[RCH][B];
      bool fwd[RCH], mid[RCH], on[RCH];
#pragma unroll
      for (int i = 0; i < RCH; ++i) {
        const int r = r0 + i * GJ_NT;
        on[i] = r < n;
        mid[i] = on[i] && r >= k0 && r < k0 + bw;
        fwd[i] = on[i] && r < k0 + bw;   // rows above the block and block rows need the U solve first
#pragma unroll
        for (int c = 0; c < B; ++c) {
          const double v = (on[i] && c < bw && !mid[i]) ? pan[(size_t)c * n + r] : 0.0;
          y[i][c] = (mid[i] && c == r - k0) ? 1.0 : v;
        }
      }
      if (r0 < k0 + bw) {   // (rows below the block skip the U solve altogether)
#pragma unroll
        for (int cc = 0; cc < B; ++cc) {
          const double di = dinv[cc];
          double yc[RCH];
#pragma unroll
          for (int i = 0; i < RCH; ++i) { yc[i] = fwd[i] ? y[i][cc] * di : 0.0; if (fwd[i]) y[i][cc] = yc[i]; }
#pragma unroll
          for (int m = cc + 1; m < B; ++m) {
            const double u = LU[cc * B + m];
#pragma unroll
            for (int i = 0; i < RCH; ++i) y[i][m] = fma(-yc[i], u, y[i][m]);
          }
        }
      }
#pragma unroll
      for (int cc = B - 1; cc > 0; --cc) {
#pragma unroll
        for (int m = 0; m < cc; ++m) {
          const double l = LU[cc * B + m];
#pragma unroll
          for (int i = 0; i < RCH; ++i) y[i][m] = fma(-y[i][cc], l, y[i][m]);
        }
      }
#pragma unroll
      for (int i = 0; i < RCH; ++i) {
        if (!on[i]) continue;
        const int r = r0 + i * GJ_NT;
#pragma unroll
        for (int c = 0; c < B; ++c) if (c < bw) pan[(size_t)c * n + r] = mid[i] ? y[i][c] : -y[i][c];
      }
    }
  }
  __syncthreads();
  if (HANK_GJ_DEBUG && tid == 0) gj_dbg[4] = clock64();
  if (bulk) {
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy writes above, before the bulk read
      bulk_s2g(Ap, pan, (uint32_t)tot * 8u);
    }
  } else {
#pragma unroll 8
    for (int i = tid; i < tot; i += GJ_PT) Ap[i] = pan[i];
  }
  if (HANK_GJ_DEBUG) { __syncthreads(); if (tid == 0) gj_dbg[5] = clock64(); }
}

// Interchanges + rank-bw update of TC columns per tile (all rows); a CTA loops over tiles blockIdx.x, += gridDim.x.
// smem: stage[GJ_MAXMOVES][TC] | slot[n].  A thread keeps the accumulators of RB of its rows x TC columns in
// registers, so a panel coefficient (one coalesced L2 load) is used TC times and a block-row value (one shared-memory
// broadcast) RB times; the panel coefficients of the rows are requested half a block ahead of the arithmetic.
template <int B, int TC>
__global__ void __launch_bounds__(GJ_UT)
k_gj_update(double* __restrict__ A, int n, int k0, int bw, const int* __restrict__ moves) {
  constexpr int RB = 5, HB = B >= 16 ? (TC >= 8 ? B / 4 : B / 2) : (B >= 8 && TC >= 8 ? B / 2 : B);
  extern __shared__ __align__(16) unsigned char smem_gj[];
  double* stage = reinterpret_cast<double*>(smem_gj);
  signed char* slot = reinterpret_cast<signed char*>(stage + GJ_MAXMOVES * TC);   // [n]: index into the move list, -1 otherwise
  const int tid = threadIdx.x;
  // tiles that lie inside the panel are skipped when the panel is tile-aligned (then `nlive` tiles remain)
  const bool aligned = k0 % TC == 0 && bw % TC == 0;
  const int ntiles = (n + TC - 1) / TC, dead0 = aligned ? k0 / TC : ntiles, ndead = aligned ? bw / TC : 0;
  const int nlive = ntiles - ndead;
  for (int r = tid; r < n; r += GJ_UT) slot[r] = -1;
  __syncthreads();
  if (tid < GJ_MAXMOVES) { const int pos = moves[tid]; if (pos >= 0) slot[pos] = (signed char)tid; }
  const double* G = A + (size_t)k0 * n;
  for (int lt = blockIdx.x; lt < nlive; lt += gridDim.x) {
    const int col0 = (lt < dead0 ? lt : lt + ndead) * TC;
    __syncthreads();   // slot[] complete / the previous tile's stage[] no longer read
    for (int q = tid; q < GJ_MAXMOVES * TC; q += GJ_UT) {
      const int i = q / TC, tc = q - i * TC;
      const int pos = moves[i], src = moves[GJ_MAXMOVES + i], col = col0 + tc;
      stage[q] = (pos >= 0 && col < n) ? A[(size_t)col * n + src] : 0.0;
    }
    __syncthreads();
    // block rows after the interchanges = entries 0..bw-1 of the move list; columns of the panel itself are skipped
    bool live[TC];
#pragma unroll
    for (int tc = 0; tc < TC; ++tc) { const int col = col0 + tc; live[tc] = col < n && !(col >= k0 && col < k0 + bw); }
    for (int r0 = tid; r0 < n; r0 += RB * GJ_UT) {
      double acc[RB][TC];
      bool on[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) on[i] = r0 + i * GJ_UT < n;
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = r0 + i * GJ_UT;
        const int s = on[i] ? slot[r] : -1;
        const bool mid = r >= k0 && r < k0 + bw;
#pragma unroll
        for (int tc = 0; tc < TC; ++tc)
          acc[i][tc] = (!on[i] || !live[tc] || mid) ? 0.0 : (s >= 0 ? stage[s * TC + tc] : A[(size_t)(col0 + tc) * n + r]);
      }
#pragma unroll
      for (int h = 0; h < B / HB; ++h) {
        double g[HB][RB];   // HB x RB independent L2 loads in flight, then the arithmetic
#pragma unroll
        for (int c = 0; c < HB; ++c)
#pragma unroll
          for (int i = 0; i < RB; ++i) g[c][i] = (on[i] && h * HB + c < bw) ? G[(size_t)(h * HB + c) * n + r0 + i * GJ_UT] : 0.0;
#pragma unroll
        for (int c = 0; c < HB; ++c) {
          double cm[TC];
#pragma unroll
          for (int tc = 0; tc < TC; ++tc) cm[tc] = stage[(h * HB + c) * TC + tc];
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int tc = 0; tc < TC; ++tc) acc[i][tc] = fma(g[c][i], cm[tc], acc[i][tc]);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i)
        if (on[i])
#pragma unroll
          for (int tc = 0; tc < TC; ++tc)
            if (live[tc]) A[(size_t)(col0 + tc) * n + r0 + i * GJ_UT] = acc[i][tc];
    }
  }
}

// rp[i] = original row now at position i after all interchanges (serial by nature: n dependent swaps in shared memory)
__global__ void k_gj_rowperm(const int* __restrict__ ipiv, int n, int* __restrict__ rp) {
  extern __shared__ int srp[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) srp[i] = i;
  int* sp = srp + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sp[i] = ipiv[i];
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < n; ++k) {
      const int p = sp[k];
      if (p != k) { const int t = srp[k]; srp[k] = srp[p]; srp[p] = t; }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) rp[i] = srp[i];
}
// (P A)^-1 = A^-1 P^-1  ->  A^-1[:, rp[i]] = X[:, i]
__global__ void k_gj_unscramble(const double* __restrict__ X, int n, const int* __restrict__ rp, double* __restrict__ out) {
  const int i = blockIdx.y;
  const double* src = X + (size_t)i * n;
  double* dst = out + (size_t)rp[i] * n;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) dst[r] = src[r];
}

template <int B, int TC>
static void gj_update_launch(hank_ctx* c, double* X, int n, int k0, int bw, const int* mv) {
  const size_t usm = (size_t)(GJ_MAXMOVES * TC) * sizeof(double) + (size_t)((n + 15) / 16 * 16);
  const int ntiles = (n + TC - 1) / TC;
  k_gj_update<B, TC><<<std::min(ntiles, c->sm_count), GJ_UT, usm, c->stream>>>(X, n, k0, bw, mv);
}
template <int B>
static int gj_run(hank_ctx* c, double* X, int n, double* out, int* iw) {
  int* ipiv = iw; int* rp = iw + n; int* info = iw + 2 * n; int* moves = iw + 2 * n + 4;
  const size_t psm = gj_panel_smem<B>(n);
  CK(cudaFuncSetAttribute(k_gj_panel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
  CK(cudaMemsetAsync(info, 0, sizeof(int), c->stream));
  const int nsteps = (n + B - 1) / B;
  // columns per update tile: 4 while that covers the matrix in one wave of CTAs (one per SM), else 8
  const int tc = (n + 3) / 4 <= c->sm_count ? 4 : 8;
  for (int s = 0; s < nsteps; ++s) {
    const int k0 = s * B, bw = std::min(B, n - k0);
    int* mv = moves + (size_t)(s & 1) * 2 * GJ_MAXMOVES;
    k_gj_panel<B><<<1, GJ_PT, psm, c->stream>>>(X, n, k0, bw, ipiv, mv, info);
    if (tc == 4) gj_update_launch<B, 4>(c, X, n, k0, bw, mv);
    else gj_update_launch<B, 8>(c, X, n, k0, bw, mv);
  }
  k_gj_rowperm<<<1, 256, 2 * (size_t)n * sizeof(int), c->stream>>>(ipiv, n, rp);
  k_gj_unscramble<<<dim3(4, n), 256, 0, c->stream>>>(X, n, rp, out);
  c->launches += 2 * nsteps + 2;
  return cuda_check(c, cudaGetLastError(), "gauss-jordan inverse");
}

// ints of workspace dense_inverse_dev needs
size_t dense_inverse_iwork(int n) { return 2 * (size_t)n + 4 + 4 * GJ_MAXMOVES; }

// out = X^-1 (both n x n column-major on the device; X is destroyed).  The block width is the largest of 16 / 8 / 4
// whose n x B panel fits in shared memory.  *d_info (device, = iw + 2n) is 0, or k+1 if the k-th pivot was exactly zero.
int dense_inverse_dev(hank_ctx* c, double* X, int n, double* out, int* iw) {
  if (n < 1) return set_error(c, HANK_ERR_ARG, "dense inverse: n < 1");
  if (gj_panel_smem<16>(n) <= (size_t)c->smem_max) return gj_run<16>(c, X, n, out, iw);
  if (gj_panel_smem<8>(n) <= (size_t)c->smem_max) return gj_run<8>(c, X, n, out, iw);
  if (gj_panel_smem<4>(n) <= (size_t)c->smem_max) return gj_run<4>(c, X, n, out, iw);
  return set_error(c, HANK_ERR_ARG, "dense inverse: n = " + std::to_string(n) + " exceeds the shared-memory panel (n <= ~7000)");
}

}  // namespace hank

using namespace hank;

// Host-pointer entry point (tests, and callers that want J̅⁻¹ itself): Ainv = A^-1, both n x n column-major.
extern "C" int hank_dense_inverse(hank_ctx* c, int n, const double* A, double* Ainv) {
  if (!c || !A || !Ainv || n < 1) return HANK_ERR_ARG;
  CK(cudaSetDevice(c->device));
  const size_t nn = (size_t)n * n;
  double* d = nullptr; int* iw = nullptr;
  auto run = [&]() -> int {
    CK(cudaMalloc((void**)&d, 2 * nn * sizeof(double)));
    CK(cudaMalloc((void**)&iw, dense_inverse_iwork(n) * sizeof(int)));
    CK(cudaMemcpyAsync(d, A, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (getenv("HANK_DENSE_TIME")) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, c->stream); }
    int rc = dense_inverse_dev(c, d, n, d + nn, iw);
    if (rc) return rc;
    if (e0) {
      cudaEventRecord(e1, c->stream); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      fprintf(stderr, "[hank_dense_inverse] n = %d: %.3f ms on the device (copies excluded)\n", n, ms);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    int h_info = 0;
    CK(cudaMemcpyAsync(&h_info, iw + 2 * n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(Ainv, d + nn, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (HANK_GJ_DEBUG) {
      long long d[8];
      cudaMemcpyFromSymbol(d, gj_dbg, sizeof d);
      fprintf(stderr, "[gj panel, last step] load %lld | first column to 2nd barrier %lld | all columns %lld | LU copy + moves %lld | row solves %lld | store %lld cycles\n",
              d[1] - d[0], 0LL, d[2] - d[1], d[3] - d[2], d[4] - d[3], d[5] - d[4]);
    }
    if (h_info > 0)
      return set_error(c, HANK_ERR_CUDA, "dense inverse: the matrix is singular (pivot " + std::to_string(h_info) + " is exactly zero)");
    return HANK_OK;
  };
  const int rc = run();
  if (d) cudaFree(d);
  if (iw) cudaFree(iw);
  return rc;
}
