// hank_primal_dsmem.cuh — cluster primal sweeps whose per-period exchange goes through distributed
// shared memory instead of global memory.
//
// In hank_primal_cluster.cuh every period ends in a barrier.cluster whose release fence has to wait
// for ALL global stores of the period — the exchange column and the 60 B of tape per point — which
// showed up as the top stall of both kernels (`membar` 4.3 / 2.5 cycles per issue,
// profiles/r01_notes.md).  Here each thread hands its value straight to the shared memory of all
// NE CTAs with `st.async.shared::cluster ... mbarrier::complete_tx::bytes`; a CTA waits on its own
// mbarrier until the NE columns of the period (NE*n_a*8 bytes) have landed.  No fence, no cluster
// barrier inside the loop, and the tape stores stay fire-and-forget.
//
// Two buffers / two mbarriers alternate by period.  Safety without further synchronisation:
//  * a peer can only send period i+1 after it has received period i from every CTA, so bytes for
//    buffer b never arrive while the previous phase of mbarrier b is still open;
//  * a buffer is overwritten by period i+2 data only after all of this CTA's period i+1 sends,
//    which every thread issues after its own reads of period i.
// The per-point arithmetic is the same as in the other primal kernels, in the same order.
#pragma once
#include <cooperative_groups.h>
#include "hank_kernels.cuh"
#include "hank_tangent_tma.cuh"   // mbarrier helpers

namespace hank {

__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v),
               "r"(remote_bar)
               : "memory");
}
// Data written by peer CTAs into THIS CTA's shared memory (st.async + complete_tx).  The wait uses the default
// CTA-scope acquire, as for TMA loads: shared memory has one copy, and the cluster-scope acquire makes ptxas emit
// CCTL.IVALL (L1 invalidate-all) after every wait (profiles/r02_notes.md).  -DHANK_ACQUIRE_CLUSTER restores it.
// Bounded like mbar_wait.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
#ifdef HANK_ACQUIRE_CLUSTER
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
#else
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
#endif
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

template <int NE, int LDA>
constexpr size_t bp_ds_smem() { return ((size_t)2 * NE * LDA + 3 * LDA) * 8 + 16; }
template <int NE, int LDA>
constexpr size_t fp_ds_smem() { return ((size_t)2 * NE * LDA + 3 * LDA) * 8 + ((size_t)LDA + LDA + 4) * 4 + 16; }

// ======================================================================================
// Backward primal sweep.  smem: xb[2][NE][LDA] ∂V/∂a of all columns | ks[2][LDA] | g[LDA] | bar[2]
// ======================================================================================
// Launched with NT + 32 threads: the extra warp owns no rows.  It joins every CTA barrier and, when `flags` is given,
// publishes after the barrier of iteration it that this CTA's column of the backward tape is complete for the first
// `it` iterations (periods P-1 .. P-it): __threadfence + store, off the compute warps' critical path.  A backward
// tangent sweep running NEXT TO this kernel follows those counters (k_backward_tangent_ring_ne, wait_tape_flag).
template <int NE, int R, int NT, bool G2>
__device__ __forceinline__ void bp_ds_body(const Consts<NE>& M, const Tape& tp, const double* __restrict__ grid,
                                           const double* __restrict__ valueT, const double* __restrict__ rpath,
                                           const double* __restrict__ wpath, int* __restrict__ status, int* __restrict__ flags) {
  constexpr int LDA = NT * R;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ __align__(16) double smem_ds[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int e = (int)cluster.block_rank();
  const int n_a = M.n_a, P = M.P;
  double* xb = smem_ds;
  double* ksb = xb + 2 * GP;
  double* g = ksb + 2 * LDA;
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + LDA);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int a = tid; a < n_a; a += NT) g[a] = grid[a];
  double pi_row[NE];  // Π[e, ·]
#pragma unroll
  for (int e2 = 0; e2 < NE; ++e2) pi_row[e2] = M.Pi[0][0];
#pragma unroll
  for (int e1 = 0; e1 < NE; ++e1)
    if (e1 == e) {
#pragma unroll
      for (int e2 = 0; e2 < NE; ++e2) pi_row[e2] = M.Pi[e1][e2];
    }
  double ze = M.z[0];
#pragma unroll
  for (int e1 = 0; e1 < NE; ++e1) if (e1 == e) ze = M.z[e1];
  // this thread's slot (column e, row tid) of buffer 0 in every CTA, and every CTA's barrier 0
  uint32_t rdst[NE], rbar[NE];
  {
    const uint32_t mine = smem_u32(xb + (size_t)e * LDA + tid), b0 = smem_u32(bar);
#pragma unroll
    for (int c = 0; c < NE; ++c) { rdst[c] = map_to_cta(mine, c); rbar[c] = map_to_cta(b0, c); }
  }
  cluster.sync();   // every CTA's barriers exist before anyone sends
  if (tid >= NT) {   // ---- the publishing warp
    for (int it = 0; it <= P; ++it) {
      __syncthreads();   // every row of this CTA is through iteration it-1
      if (flags && tid == NT && it > 0) {
        __threadfence();
        *reinterpret_cast<volatile int*>(flags + e) = it;
      }
    }
    cluster.sync();
    return;
  }
  const uint32_t col_bytes = (uint32_t)NE * (uint32_t)n_a * 8u;
  double rn = rpath[P - 1], wn = wpath[P - 1];
  double Vlast[R];
  for (int it = 0; it < P; ++it) {
    const int t = P - 1 - it, wb = it & 1;
    const double r = rn, w = wn;
    if (t > 0) { rn = rpath[t - 1]; wn = wpath[t - 1]; }
    const double opr = 1.0 + r, rho = 1.0 / opr;
    if (tid == 0 && e == 0) tp.rho[t] = rho;
    if (tid == 0) mbar_expect_tx(&bar[wb], col_bytes);   // this period's incoming columns
    double* bwf = bw_fields<LDA>(tp, NE, t, e) + tid;
    int* bwi = bw_idx<LDA>(tp, NE, t, e) + tid;
    double* polt = tp.pol + (size_t)t * GP + (size_t)e * LDA + tid;
    double* ks = ksb + wb * LDA;
    const double* vsrc = xb + (size_t)(wb ^ 1) * GP;
    if (it > 0) mbar_wait_cluster(&bar[wb ^ 1], ((it - 1) >> 1) & 1);
    // ---- phase 1: Euler inversion for this column
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        double vrow[NE];
        if (it == 0) {
#pragma unroll
          for (int e2 = 0; e2 < NE; ++e2) vrow[e2] = valueT[(size_t)e2 * LDA + a];
        } else {
#pragma unroll
          for (int e2 = 0; e2 < NE; ++e2) vrow[e2] = vsrc[(size_t)e2 * LDA + a];
        }
        double ev = 0.0;
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) ev += vrow[e2] * pi_row[e2];
        const EulerPoint u = egm_euler_point<G2>(M, ev, w * ze, g[a], rho, opr, status, a, e, t);
        ks[a] = u.knot;
        bwf[BW_A1 * LDA + j * NT] = u.a1;
        bwf[BW_KR * LDA + j * NT] = u.kr;
      }
    }
    __syncthreads();
    // ---- phase 2: interpolation on the exogenous grid, clamp, marginal value
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const InterpPoint q = egm_interp_point<G2, LDA>(M, ks, g, n_a, a, w * ze, opr, status, e, t);
        // hand ∂V/∂a(a, e) to every CTA first: it is what the next period waits for
        const uint32_t off = (uint32_t)(wb * GP + j * NT) * 8u;
#pragma unroll
        for (int c = 0; c < NE; ++c) st_async_f64(rdst[c] + off, q.vnew, rbar[c] + 8u * wb);
        polt[j * NT] = q.p;
        bwi[j * NT] = q.i;
        bwf[BW_CA * LDA + j * NT] = q.cA;
        bwf[BW_CB * LDA + j * NT] = q.cB;
        bwf[BW_VF * LDA + j * NT] = q.vf;
        bwf[BW_E1 * LDA + j * NT] = q.E1;
        Vlast[j] = q.vnew;
      }
    }
  }
  __syncthreads();   // (the publishing warp's last count: all P periods of this column are written)
  // the last period's columns are still in flight towards this CTA: drain before anyone exits
  mbar_wait_cluster(&bar[(P - 1) & 1], ((P - 1) >> 1) & 1);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int a = tid + j * NT;
    if (a < n_a) tp.value_first[(size_t)e * LDA + a] = Vlast[j];
  }
  cluster.sync();
  if (tid == 0) {   // (k_primal_ds_both reuses this shared memory for the forward sweep)
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
  }
}
template <int NE, int R, int NT, bool G2>
__global__ void __launch_bounds__(NT + 32, 1)
k_backward_primal_ds(const Consts<NE> M, const Tape tp, const double* __restrict__ grid,
                     const double* __restrict__ valueT, const double* __restrict__ rpath,
                     const double* __restrict__ wpath, int* __restrict__ status, int* __restrict__ flags) {
  bp_ds_body<NE, R, NT, G2>(M, tp, grid, valueT, rpath, wpath, status, flags);
}

// ======================================================================================
// Forward primal sweep.  smem: xb[2][NE][LDA] post-lottery masses of all columns | X[LDA] | Y[LDA] |
// g[LDA] | ms[LDA] (int) | st[LDA+4] (int) | bar[2].   kdpart: [P][NE*NT/32].
// ======================================================================================
// (pol_in is not __restrict__: in k_primal_ds_both it is the policy the backward sweep of the same kernel wrote)
template <int NE, int R, int NT>
__device__ __forceinline__ void fp_ds_body(const Consts<NE>& M, const Tape& tp, const double* __restrict__ grid,
                                           const double* __restrict__ D0, const double* pol_in,
                                           double* __restrict__ kdpart, int* __restrict__ status) {
  constexpr int LDA = NT * R, NS = LDA + 4, NW = NT / 32;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ __align__(16) double smem_ds[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int e = (int)cluster.block_rank();
  const int n_a = M.n_a, P = M.P;
  double* xb = smem_ds;
  double* X = xb + 2 * GP;
  double* Y = X + LDA;
  double* g = Y + LDA;
  int* ms = reinterpret_cast<int*>(g + LDA);
  int* st = ms + LDA;
  uint64_t* bar = reinterpret_cast<uint64_t*>(st + NS);   // NS*4 is a multiple of 8
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int a = tid; a < n_a; a += NT) g[a] = grid[a];
  for (int i = tid; i < NS; i += NT) st[i] = n_a;
  double pi_col[NE];  // Π[·, e]
#pragma unroll
  for (int e1 = 0; e1 < NE; ++e1) pi_col[e1] = M.Pi[0][0];
#pragma unroll
  for (int e2 = 0; e2 < NE; ++e2)
    if (e2 == e) {
#pragma unroll
      for (int e1 = 0; e1 < NE; ++e1) pi_col[e1] = M.Pi[e1][e2];
    }
  double D[R], pc[R], pn[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int a = tid + j * NT;
    D[j] = a < n_a ? D0[(size_t)e * LDA + a] : 0.0;
    pc[j] = a < n_a ? pol_in[(size_t)e * LDA + a] : 0.0;
    pn[j] = 0.0;
  }
  uint32_t rdst[NE], rbar[NE];
  {
    const uint32_t mine = smem_u32(xb + (size_t)e * LDA + tid), b0 = smem_u32(bar);
#pragma unroll
    for (int c = 0; c < NE; ++c) { rdst[c] = map_to_cta(mine, c); rbar[c] = map_to_cta(b0, c); }
  }
  cluster.sync();
  const uint32_t col_bytes = (uint32_t)NE * (uint32_t)n_a * 8u;
  for (int t = 0; t < P; ++t) {
    const int wb = t & 1;
    if (tid == 0) mbar_expect_tx(&bar[wb], col_bytes);
    const double* polt = pol_in + (size_t)t * GP + (size_t)e * LDA + tid;
    if (t + 1 < P) {
#pragma unroll
      for (int j = 0; j < R; ++j) pn[j] = tid + j * NT < n_a ? polt[GP + j * NT] : 0.0;
    }
    double* fwf = fw_fields<LDA>(tp, NE, t, e) + tid;
    int* mbt = tp.mbr + (size_t)t * GP + (size_t)e * LDA + tid;
    // ---- phase A: brackets and lottery masses of this column
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const double p = pc[j];
        const LotteryPoint q = lottery_point<LDA>(g, n_a, p, D[j]);
        X[a] = q.om * D[j];
        Y[a] = (1.0 - q.om) * D[j];
        ms[a] = q.m;
        fwf[FW_OM * LDA + j * NT] = q.om;
        fwf[FW_DCO * LDA + j * NT] = q.dco;
        fwf[FW_P * LDA + j * NT] = p;
        mbt[j * NT] = q.m;
      }
    }
    __syncthreads();
    // ---- phase B: source-range starts per destination row
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        lottery_starts_point(ms, st, n_a, a, status, e, t);
      }
    }
    __syncthreads();
    // ---- phase C: gather in ascending source order; hand this column's masses to every CTA
    int* so = fw_start<LDA>(tp, NE, t, e);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const int s0 = st[a + 1], s1 = st[a + 2], s2 = st[a + 3];
        const double acc = lottery_gather_point(X, Y, s0, s1, s2);
        const uint32_t off = (uint32_t)(wb * GP + j * NT) * 8u;
#pragma unroll
        for (int c = 0; c < NE; ++c) st_async_f64(rdst[c] + off, acc, rbar[c] + 8u * wb);
        so[a + 1] = s0;
        if (a == n_a - 1) { so[a + 2] = s1; so[a + 3] = s2; }
      }
    }
    __syncthreads();   // X, Y, ms, st are rewritten by the next period's phases
    mbar_wait_cluster(&bar[wb], (t >> 1) & 1);
    // ---- Markov mix for this column and its share of <p_t, D_t>
    const double* tsrc = xb + (size_t)wb * GP;
    double kacc = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        double d = 0.0;
#pragma unroll
        for (int e1 = 0; e1 < NE; ++e1) d += pi_col[e1] * tsrc[(size_t)e1 * LDA + a];
        D[j] = d;
        fwf[FW_D * LDA + j * NT] = d;
        kacc += pc[j] * d;
      }
    }
    kacc = warp_sum(kacc);
    if (lane == 0) kdpart[((size_t)t * NE + e) * NW + warp] = kacc;
#pragma unroll
    for (int j = 0; j < R; ++j) pc[j] = pn[j];
  }
  cluster.sync();   // no CTA leaves while a peer may still be writing into it
}
template <int NE, int R, int NT>
__global__ void __launch_bounds__(NT, 1)
k_forward_primal_ds(const Consts<NE> M, const Tape tp, const double* __restrict__ grid,
                    const double* __restrict__ D0, const double* __restrict__ pol_in,
                    double* __restrict__ kdpart, int* __restrict__ status) {
  fp_ds_body<NE, R, NT>(M, tp, grid, D0, pol_in, kdpart, status);
}

// ======================================================================================
// Both primal sweeps of a linearisation in ONE launch (pipelined linearisation, hank_ctx.h): the cluster keeps its n_e
// SMs from the first backward period to the last forward period.  With two launches the SMs of the backward sweep go to
// the pending CTAs of the tangent sweep running next to it the moment it ends, and the forward primal sweep then waits
// for a wave boundary (measured: no gain from the overlap at all, profiles/r02_notes.md).  CTA e reads back only the
// policy column it wrote itself.
// ======================================================================================
template <int NE, int R, int NT, bool G2>
__global__ void __launch_bounds__(NT + 32, 1)
k_primal_ds_both(const Consts<NE> M, const Tape tp, const double* __restrict__ grid, const double* __restrict__ valueT,
                 const double* __restrict__ rpath, const double* __restrict__ wpath, int* __restrict__ status,
                 int* __restrict__ flags, const double* __restrict__ D0, double* __restrict__ kdpart) {
  bp_ds_body<NE, R, NT, G2>(M, tp, grid, valueT, rpath, wpath, status, flags);
  if (threadIdx.x >= NT) return;   // the publishing warp is done (exited warps do not count at later barriers)
  __syncthreads();                 // (the invalidated barriers' bytes become the forward sweep's bracket array)
  fp_ds_body<NE, R, NT>(M, tp, grid, D0, tp.pol, kdpart, status);
}

}  // namespace hank
