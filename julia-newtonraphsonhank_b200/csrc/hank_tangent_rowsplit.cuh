// hank_tangent_rowsplit.cuh — tangent-lane sweeps for FEW lanes: one lane group spread over a
// thread-block cluster, the asset rows split across the cluster's CTAs.
//
// Why: in hank_tangent_tma.cuh a lane group is one CTA, so a pass with fewer lane groups than SMs
// leaves SMs idle and every busy SM has to pull the whole primal tape of a period (n_e column chunks)
// through its own L2 port (~43 B/clk/SM): 2.4 / 3.4 us per period and lane at 500x7, whatever the
// arithmetic.  That is the regime of Newton's inner JVP (K = 1, NewtonRaphson.jl:95), of the 64-lane
// passes on the large grids, and of every rank of a Jacobian build sharded over 8 GPUs.
//
// Here a cluster of NC CTAs carries L lanes; CTA `rank` owns asset rows [rank*NT, (rank+1)*NT) of every
// income state (one row per thread, all n_e states of the row in that thread's registers, so the Markov
// mix stays thread-local exactly as in the one-CTA kernels).  Each CTA stages only ITS rows of the tape
// (n_e*52*NT bytes per period instead of n_e*52*LDA) with cp.async.bulk copies (one per field and
// column, SASS UBLKCP) into a ring several periods/columns ahead.
//
// The only coupling across rows is the interpolation / lottery neighbourhood: a row reads k̇ at its
// knot indices i, i+1 (backward) or the masses x, y of its source rows (forward), which may live in
// another CTA of the cluster.  Those values are written to the owner's shared memory and PULLED
// through distributed shared memory (mapa + ld.shared::cluster; plain ld.shared when the owner is the
// reader itself).  Hand-shake per exchange group (GC columns): the owner's threads write, one
// barrier among the compute warps (which drains their shared-memory stores), NC threads
// `mbarrier.arrive.shared::cluster` on every CTA's `ready` barrier (count NC), readers `try_wait` on
// their own.  No cluster barrier inside the sweep.  The waits use the default (CTA-scope) acquire, as
// CUTLASS's ClusterBarrier does: shared memory has a single copy, so nothing has to be invalidated; the
// cluster-scope acquire makes ptxas emit CCTL.IVALL (L1 invalidate-all), which was 27 % of all stall
// samples of the first version of these kernels (profiles/r02_notes.md).
//
// A dedicated producer warp (warp NT/32) issues the bulk copies; slots are handed back through `empty`
// mbarriers (one arrival per compute warp).  A bulk copy costs the issuing warp ~60 cycles whatever its
// size (UBLKCP runs on the uniform datapath, one lane at a time), so the rows of a CTA have to be ONE
// contiguous piece per exchange group: the kernels read a row-block-major copy of the tape,
//   backward: [t][rank][e][6 fields x NT doubles | NT ints],   forward: [t][rank][e][4 fields x NT doubles | NT+4 ints],
// made by k_tape_rowblocks after the primal sweeps (~100 MB moved at 500x7, T=300: tens of microseconds
// per linearisation), and exchange ṗ between the two sweeps as [t][cluster][rank][e][lane][NT].  With
// the column-major tape of the one-CTA kernels a period was 49 copies of 512 B per CTA and the producer
// set the pace at ~3300 cycles per period (profiles/r02_notes.md).
//
// Software pipeline: the write stage (A) of group q runs LA groups ahead of its read stage (B), so the
// DSMEM round trip of one group hides behind the arithmetic of others.  With stage order
// [A(q), sync, arrive(q), B(q-LA)] a writer that reaches A(q) knows every peer has finished B(q-2LA-2),
// hence NB = 2LA+2 exchange buffers are enough and no "buffer free" signal is needed.  GC = NE, LA = 0
// exchanges a whole period at once (fewest hand-shakes: the latency regime); GC = 1, LA = 2 streams
// column by column (smallest buffers: the large grids).
//
// Summation order: the same ascending-source order as gather_row in hank_tangent.cuh for the first two
// sources of each range; the parity bar (1e-10 relative / 1e-12 absolute) holds either way.
#pragma once
#include "hank_tangent.cuh"
#include "hank_tangent_tma.cuh"
#include "hank_primal_dsmem.cuh"   // map_to_cta, mbar_wait_cluster

namespace hank {

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The remote arrival RELEASES at cluster scope: it publishes shared memory that other threads of the CTA wrote before the
// named barrier in front of it, to readers in other CTAs.  A relaxed arrival here (round 2's first version) was a real
// race: alone on the GPU it never showed in thousands of sweeps, but with the Jacobian build's overflow clusters
// running next to the 148-CTA main wave one unit-seed column in ~15 builds came out 3e-7 .. 1e-4 off (a pulled value
// read before the owner's store was visible to the cluster).  Measured remedies (profiles/r02_notes.md): a CTA-scope
// fence before a relaxed arrival and a cluster-scope acquire on the reader's side both still fail; only the
// cluster-scope release is clean (0 of 200 builds), and it costs ~0.45 us per hand-shake whoever issues it.
// The arrivals are issued by a SIGNALLING WARP that owns no rows: the compute warps `bar.arrive` on a named barrier
// after their shared-memory writes and go on to the work that does not depend on the peers; the signalling warp
// `bar.sync`s on it and then releases.  The slot hand-back to the producer warp stays relaxed: the loads of that slot
// have already been consumed.  (The fence-free way to publish would be to PUSH the values with st.async +
// complete_tx into every reader's shared memory, as the primal sweeps do; that is the next step for these kernels.)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// (HANK_RS_RELAXED=1, A/B measurements only: the racy fast form)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// the signalling warp's loop: one hand-shake per exchange group, buffers in ring order
template <int NC, int NB>
__device__ __forceinline__ void signal_groups(uint64_t* ready, int ngroups, int nthreads_bar, int lane, int release) {
  const uint32_t rdy_remote = lane < NC ? map_to_cta(smem_u32(ready), (uint32_t)lane) : 0u;
  int b = 0;
  for (int q = 0; q < ngroups; ++q) {
    asm volatile("bar.sync %0, %1;" ::"r"(1), "r"(nthreads_bar) : "memory");
    if (lane < NC) { if (release) mbar_arrive_remote(rdy_remote + 8u * (uint32_t)b); else mbar_arrive_remote_relaxed(rdy_remote + 8u * (uint32_t)b); }
    if (++b == NB) b = 0;
  }
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double ld_cluster_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// element `off` (doubles) of an exchange buffer whose copy in this CTA starts at `lbase` (generic
// pointer) / `sbase` (shared address), read from CTA `owner` of the cluster.  BR: take the plain
// ld.shared path when the owner is the reader (the throughput shapes: DSMEM moves ~20 B/clk per SM);
// without it every read goes through mapa + ld.shared::cluster (the latency shapes: no branches).
template <bool BR>
__device__ __forceinline__ double pull(const double* lbase, uint32_t sbase, int off, uint32_t owner, uint32_t rank) {
  if (BR) { if (owner == rank) return lbase[off]; }
  return ld_cluster_f64(map_to_cta(sbase + (uint32_t)off * 8u, owner));
}

__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

template <int NT> __host__ __device__ constexpr size_t rs_bw_col_bytes() { return (size_t)BW_NF * NT * 8 + (size_t)NT * 4; }
template <int NT> __host__ __device__ constexpr size_t rs_fw_tape_col_bytes() { return (size_t)FW_NF * NT * 8 + (size_t)(NT + 4) * 4; }
template <int NT, int L> __host__ __device__ constexpr size_t rs_fw_col_bytes() { return rs_fw_tape_col_bytes<NT>() + (size_t)L * NT * 8; }
template <int NT, int L, int GC, int LA>
constexpr size_t rs_bw_smem(int S) { return (size_t)S * GC * rs_bw_col_bytes<NT>() + (size_t)(2 * LA + 2) * GC * L * NT * 8 + (size_t)(2 * S + 2 * LA + 2) * 8 + 128; }
template <int NT, int L, int GC, int LA>
constexpr size_t rs_fw_smem(int S) {
  return (size_t)S * GC * rs_fw_col_bytes<NT, L>() + (size_t)(2 * LA + 2) * 2 * GC * L * NT * 8 + (size_t)2 * L * (NT / 32) * 8 +
         (size_t)(2 * S + 2 * LA + 2) * 8 + 128;
}

// ring / buffer cursor: index and phase parity, advanced without divisions
struct Cursor {
  int i = 0, par = 0;
  __device__ __forceinline__ void next(int n) { if (++i == n) { i = 0; par ^= 1; } }
};

// ======================================================================================
// Backward tangent sweep, rows split over a cluster.  Grid = NC * ceil(K / L) CTAs of NT + 32 threads
// (NT/32 compute warps + the producer warp), cluster (NC,1,1).
// smem: ring[S][GC][6 fields x NT doubles | NT ints] | kb[NB][GC][L][NT] | full[S] | empty[S] | ready[NB]
// tape_rs: row-block-major backward tape.  dpol (written): [P][cluster][rank][NE][L][NT].
// thi (nullable): seed horizons per kThiGroup lanes (hank_ks_jacobian_columns).
// ======================================================================================
template <int NE, int NC, int NT, int L, int GC, int LA>
__global__ void __launch_bounds__(NT + 64, 1)
k_backward_tangent_rs(const Consts<NE> M, const Tape tp, const unsigned char* __restrict__ tape_rs, int K, int S,
                      const int* __restrict__ thi, const double* __restrict__ dr, const double* __restrict__ dw,
                      double* __restrict__ dpol, int release) {
  static_assert(GC >= 1 && GC <= NE, "columns per exchange group");   // the last group may be shorter
  static_assert(NT % 32 == 0 && (NT & (NT - 1)) == 0, "NT must be a power of two >= 32");
  constexpr int LDA = NC * NT, NG = (NE + GC - 1) / GC, NB = 2 * LA + 2, LOGNT = ilog2c(NT), NW = NT / 32;
  constexpr int COLB = (int)rs_bw_col_bytes<NT>(), COLD = COLB / 8;   // 52*NT bytes, a multiple of 16
  constexpr int SLOTD = GC * COLD, KBD = GC * L * NT;
  constexpr bool BR = L > 1;
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster * L;
  const int Pfull = M.P, P = thi ? min(M.P, thi[lane0 / kThiGroup]) : M.P;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* kb = ring + (size_t)S * SLOTD;
  uint64_t* full = reinterpret_cast<uint64_t*>(kb + (size_t)NB * KBD);
  uint64_t* empty = full + S;
  uint64_t* ready = empty + S;
  const uint32_t kb_s = smem_u32(kb);
  const int ngroups = P * NG;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    for (int b = 0; b < NB; ++b) mbar_init(&ready[b], NC);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers exist before anyone arrives on them

  if (warp == NW) {
    // ---- producer warp: group gq = pi*NG + g covers columns [g*GC, (g+1)*GC) of period t = P-1-pi; this CTA's rows
    // of those columns are one contiguous piece of the row-block-major tape: one bulk copy
    if (lane == 0) {
      Cursor cs;
      int t = P - 1, g = 0;
      for (int gq = 0; gq < ngroups; ++gq) {
        if (gq >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        const uint32_t gcols = (uint32_t)min(GC, NE - g * GC);
        mbar_expect_tx(&full[cs.i], gcols * (uint32_t)COLB);
        bulk_g2s(ring + (size_t)cs.i * SLOTD, tape_rs + (((size_t)t * NC + rank) * NE + g * GC) * COLB, gcols * (uint32_t)COLB, &full[cs.i]);
        cs.next(S);
        if (++g == NG) { g = 0; --t; }
      }
    }
  } else if (warp == NW + 1) {
    signal_groups<NC, NB>(ready, ngroups, NT + 32, lane, release);   // ---- signalling warp (see mbar_arrive_remote)
  } else {
    double Vd[L][NE];
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
      for (int e = 0; e < NE; ++e) Vd[l][e] = 0.0;

    Cursor sA, sB, bA, bB;     // tape slot / exchange buffer of the next write (A) and read (B) stage
    double rho = __ldg(tp.rho + (P > 0 ? P - 1 : 0));
    double drn[L], dwn[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const bool on = lane0 + l < K && P > 0;
      drn[l] = on ? __ldg(dr + (size_t)(lane0 + l) * Pfull + P - 1) : 0.0;
      dwn[l] = on ? __ldg(dw + (size_t)(lane0 + l) * Pfull + P - 1) : 0.0;
    }
    for (int t = P - 1; t >= 0; --t) {
      const double rho_t = rho;
      double drl[L], dwl[L];
#pragma unroll
      for (int l = 0; l < L; ++l) { drl[l] = drn[l]; dwl[l] = dwn[l]; }
      if (t > 0) {   // next period's scalars, requested a period ahead
        rho = __ldg(tp.rho + t - 1);
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const bool on = lane0 + l < K;
          drn[l] = on ? __ldg(dr + (size_t)(lane0 + l) * Pfull + t - 1) : 0.0;
          dwn[l] = on ? __ldg(dw + (size_t)(lane0 + l) * Pfull + t - 1) : 0.0;
        }
      }
      // ---- ĖV in place of V̇⁺ (registers only)
#pragma unroll
      for (int l = 0; l < L; ++l) {
        double ev[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          double s = 0.0;
#pragma unroll
          for (int e2 = 0; e2 < NE; ++e2) s = fma(M.Pi[e][e2], Vd[l][e2], s);
          ev[e] = s;
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) Vd[l][e] = ev[e];
      }
      double* dp_t = dpol + (((size_t)t * ncl + cluster) * NC + rank) * (NE * L * NT) + tid;   // [t][cluster][rank][e][l][NT]
#pragma unroll
      for (int st = 0; st < NG + LA; ++st) {
        // ---- B1: this row's coefficients and knot indices for the columns of group st-LA, then k̇ at the two knots
        // of every lane, requested before the write stage below so that the DSMEM round trips hide behind it
        // (with LA = 0 the group is its own look-ahead: B1 has to follow the write stage)
        double cA[GC], cB[GC], E1[GC], vf[GC], k0[GC][L], k1[GC][L];
        auto B1 = [&]() {
          const double* sl = ring + (size_t)sB.i * SLOTD + tid;
          const double* kr_l = kb + (size_t)bB.i * KBD;
          const uint32_t kr_s = kb_s + (uint32_t)(bB.i * KBD) * 8u;
          int i0[GC];
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            if ((st - LA) * GC + ce >= NE) continue;   // short last group
            cA[ce] = sl[ce * COLD + BW_CA * NT]; cB[ce] = sl[ce * COLD + BW_CB * NT];
            E1[ce] = sl[ce * COLD + BW_E1 * NT]; vf[ce] = sl[ce * COLD + BW_VF * NT];
            i0[ce] = reinterpret_cast<const int*>(sl - tid + ce * COLD + BW_NF * NT)[tid];
          }
          mbar_wait(&ready[bB.i], bB.par);
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            if ((st - LA) * GC + ce >= NE) continue;
            const int i1 = i0[ce] + 1;
            const uint32_t o0 = (uint32_t)i0[ce] >> LOGNT, o1 = (uint32_t)i1 >> LOGNT;
            const int f0 = (i0[ce] & (NT - 1)) + ce * L * NT, f1 = (i1 & (NT - 1)) + ce * L * NT;
#pragma unroll
            for (int l = 0; l < L; ++l) {
              k0[ce][l] = pull<BR>(kr_l, kr_s, f0 + l * NT, o0, rank);
              k1[ce][l] = pull<BR>(kr_l, kr_s, f1 + l * NT, o1, rank);
            }
          }
        };
        if (LA > 0 && st >= LA) B1();
        if (st < NG) {   // ---- A: k̇ of this CTA's rows for the columns of group st
          const double* sl = ring + (size_t)sA.i * SLOTD + tid;
          double* kw = kb + (size_t)bA.i * KBD + tid;
          mbar_wait(&full[sA.i], sA.par);
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            const int e = st * GC + ce;
            if (e >= NE) continue;
            const double a1 = sl[ce * COLD + BW_A1 * NT], kr = sl[ce * COLD + BW_KR * NT];
            const double cw = -(rho_t * M.z[e]);
#pragma unroll
            for (int l = 0; l < L; ++l) kw[(ce * L + l) * NT] = fma(a1, Vd[l][e], fma(kr, drl[l], cw * dwl[l]));
          }
          // this thread's k̇ of the group are written: the signalling warp tells the peers.  Without look-ahead the next
          // thing a thread does is wait for the peers' arrivals of this very group, so it need not wait here; with
          // look-ahead it could reach the next group's barrier before this one completed, so it waits like everyone
          if constexpr (LA == 0) named_bar_arrive(1, NT + 32); else named_bar_sync(1, NT + 32);
          sA.next(S); bA.next(NB);
        }
        if (LA == 0) B1();
        if (st >= LA) {   // ---- B2: ṗ and V̇
          const int g = st - LA;
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            const int e = g * GC + ce;
            if (e >= NE) continue;
            const double ze = M.z[e];
            double* dpc = dp_t + e * L * NT;
#pragma unroll
            for (int l = 0; l < L; ++l) {
              const double pd = fma(cA[ce], k0[ce][l], cB[ce] * k1[ce][l]);
              __stcs(dpc + l * NT, pd);
              Vd[l][e] = fma(vf[ce], fma(ze, dwl[l], -pd), E1[ce] * drl[l]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_local(&empty[sB.i]);   // the tape slot goes back to the producer
          sB.next(S); bB.next(NB);
        }
      }
    }
  }
  __syncwarp();
  cluster_sync_all();   // peers may still be reading this CTA's last exchange buffer
}

// ======================================================================================
// Forward tangent sweep, rows split over a cluster.
// smem: ring[S][ GC x (4 fields x NT doubles | NT+4 ints) | ṗ [GC][L][NT] ] |
//       xy[NB][2][GC][L][NT] | red[2][L][NT/32] | full[S] | empty[S] | ready[NB]
// dkdpart: [K][P][NC] per-CTA partial sums of K̇D.
// ======================================================================================
// First two sources of each of the two ranges of a destination row: unconditional loads (clamped index, the caller
// masks the value) so that the loads of all columns of a group are in flight at once.
template <int L, int NT, bool BR>
__device__ __forceinline__ void gather_rs_first(const double* xl, const double* yl, uint32_t xs, uint32_t ys, int s0, int s1,
                                                int s2, uint32_t rank, double (&xv)[2][L], double (&yv)[2][L]) {
  constexpr int LOGNT = ilog2c(NT);
  const int self = (int)(rank << LOGNT);   // an absent source reads (and discards) one of this CTA's own rows
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int bx = s1 - s0 > d ? s0 + d : self, by = s2 - s1 > d ? s1 + d : self;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      xv[d][l] = pull<BR>(xl, xs, l * NT + (bx & (NT - 1)), (uint32_t)bx >> LOGNT, rank);
      yv[d][l] = pull<BR>(yl, ys, l * NT + (by & (NT - 1)), (uint32_t)by >> LOGNT, rank);
    }
  }
}
// Sources beyond the second of each range: up to eight more per thread with all loads in flight before the adds,
// long ranges (the mass piling up at the borrowing constraint) by the whole warp, one row at a time.  Same
// ascending order as gather_row in hank_tangent.cuh.  Must be called by all 32 lanes of a warp.
template <int L, int NT, bool BR>
__device__ __forceinline__ void gather_rs_rest(const double* xl, const double* yl, uint32_t xs, uint32_t ys, int s0, int s1,
                                               int s2, uint32_t rank, int lane, double (&acc)[L]) {
  constexpr int LOGNT = ilog2c(NT), U = 2, kSerial = 8;
  auto px = [&](int l, int b) { return pull<BR>(xl, xs, l * NT + (b & (NT - 1)), (uint32_t)b >> LOGNT, rank); };
  auto py = [&](int l, int b) { return pull<BR>(yl, ys, l * NT + (b & (NT - 1)), (uint32_t)b >> LOGNT, rank); };
  const int n1 = s1 - s0, n2 = s2 - s1;
  const int mx = max(n1, n2) - U;
  if (!__any_sync(0xffffffffu, mx > 0)) return;
  const int self = (int)(rank << LOGNT);
  // batches of two more sources per range (most rows that get here have three or four), loads before adds; the
  // x tail is finished before the y tail starts (gather_row's order), so the y batches wait for the x vote
  if (mx > 0 && mx <= kSerial) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      for (int d = U; d < n1; d += 2) {
        const double a = px(l, s0 + d), b = px(l, d + 1 < n1 ? s0 + d + 1 : self);
        acc[l] += a;
        if (d + 1 < n1) acc[l] += b;
      }
      for (int d = U; d < n2; d += 2) {
        const double a = py(l, s1 + d), b = py(l, d + 1 < n2 ? s1 + d + 1 : self);
        acc[l] += a;
        if (d + 1 < n2) acc[l] += b;
      }
    }
  }
  unsigned bal = __ballot_sync(0xffffffffu, mx > kSerial);
  while (bal) {
    const int src = __ffs(bal) - 1;
    bal &= bal - 1;
    const int b0 = __shfl_sync(0xffffffffu, s0, src), b1 = __shfl_sync(0xffffffffu, s1, src),
              b2 = __shfl_sync(0xffffffffu, s2, src);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      double v = 0.0;
      for (int b = b0 + U + lane; b < b1; b += 32) v += px(l, b);
      for (int b = b1 + U + lane; b < b2; b += 32) v += py(l, b);
      v = warp_sum(v);
      if (lane == src) acc[l] += v;
    }
  }
}

template <int NE, int NC, int NT, int L, int GC, int LA>
__global__ void __launch_bounds__(NT + 64, 1)
k_forward_tangent_rs(const Consts<NE> M, const unsigned char* __restrict__ tape_rs, int K, int Kp, int S,
                     const int* __restrict__ thi, const double* __restrict__ zeros, const double* __restrict__ dpol,
                     int pd_rs, double* __restrict__ dkdpart, int release) {
  static_assert(GC >= 1 && GC <= NE, "columns per exchange group");   // the last group may be shorter
  constexpr int LDA = NC * NT, NG = (NE + GC - 1) / GC, NB = 2 * LA + 2, NW = NT / 32;
  constexpr int COLB = (int)rs_fw_tape_col_bytes<NT>(), COLD = COLB / 8;   // 36*NT + 16 bytes, a multiple of 16
  constexpr int ST_OFF = FW_NF * NT;                                       // doubles from the column start
  constexpr int PD_OFF = GC * COLD;                                        // ṗ [GC][L][NT] follows the GC tape columns
  constexpr int SLOTD = GC * COLD + GC * L * NT, XYD = GC * L * NT;
  constexpr bool BR = L > 1;
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster * L;
  const int P = M.P;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* xy = ring + (size_t)S * SLOTD;
  double* red = xy + (size_t)NB * 2 * XYD;
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * L * NW);
  uint64_t* empty = full + S;
  uint64_t* ready = empty + S;
  const uint32_t xy_s = smem_u32(xy);
  const int ngroups = P * NG;
  const int pe = thi ? min(P, thi[lane0 / kThiGroup]) : P;   // ṗ is zero (and unwritten) from period pe on

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    for (int b = 0; b < NB; ++b) mbar_init(&ready[b], NC);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NW) {
    // ---- producer warp: the group's tape columns (one piece of the row-block-major tape) and ṗ of the L lanes:
    // one more piece when the backward sweep of this pass was the row-split one (pd_rs), else GC*L pieces of the
    // caller's [t][e][Kp][LDA] array (hank_forward_policies)
    Cursor cs;
    int t = 0, g = 0;
    for (int gq = 0; gq < ngroups; ++gq) {
      double* dst = ring + (size_t)cs.i * SLOTD;
      const int gcols = min(GC, NE - g * GC);
      const uint32_t PDB = (uint32_t)(gcols * L * NT * 8);
      if (lane == 0) {
        if (gq >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        mbar_expect_tx(&full[cs.i], (uint32_t)(gcols * COLB) + PDB);
        bulk_g2s(dst, tape_rs + (((size_t)t * NC + rank) * NE + g * GC) * COLB, (uint32_t)(gcols * COLB), &full[cs.i]);
        if (t >= pe) {   // beyond the seed horizon: zeros (the page holds kZeroBytes)
          for (uint32_t o = 0; o < PDB; o += kZeroBytes)
            bulk_g2s(dst + PD_OFF + o / 8, zeros, min(PDB - o, (uint32_t)kZeroBytes), &full[cs.i]);
        } else if (pd_rs) {
          bulk_g2s(dst + PD_OFF, dpol + ((((size_t)t * ncl + cluster) * NC + rank) * NE + g * GC) * (L * NT), PDB, &full[cs.i]);
        }
      }
      __syncwarp();
      if (t < pe && !pd_rs)
        for (int i = lane; i < gcols * L; i += 32) {
          const int ce = i / L, l = i - ce * L;
          bulk_g2s(dst + PD_OFF + (size_t)i * NT, dpol + ((((size_t)t * NE + g * GC + ce) * Kp + lane0 + l) * LDA + rank * NT), NT * 8,
                   &full[cs.i]);
        }
      cs.next(S);
      if (++g == NG) { g = 0; ++t; }
    }
  } else if (warp == NW + 1) {
    signal_groups<NC, NB>(ready, ngroups, NT + 32, lane, release);   // ---- signalling warp (see mbar_arrive_remote)
  } else {
    double Dd[L][NE];
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
      for (int e = 0; e < NE; ++e) Dd[l][e] = 0.0;
    Cursor sA, sB, bA, bB;
    for (int t = 0; t < P; ++t) {
      double kacc[L];
#pragma unroll
      for (int l = 0; l < L; ++l) kacc[l] = 0.0;
      double pv[NE];
#pragma unroll
      for (int st = 0; st < NG + LA; ++st) {
        // ---- B1: the first sources of this CTA's destination rows for the columns of group st-LA are requested
        // before the write stage below, so that the DSMEM round trips hide behind it
        int s0[GC], s1[GC], s2[GC];
        double xv[GC][2][L], yv[GC][2][L];
        const double* xr_l = xy + (size_t)bB.i * 2 * XYD;
        const uint32_t xr_s = xy_s + (uint32_t)(bB.i * 2 * XYD) * 8u;
        auto B1 = [&]() {   // (with LA = 0 the group is its own look-ahead: B1 has to follow the write stage)
          const double* slb = ring + (size_t)sB.i * SLOTD;
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            if ((st - LA) * GC + ce >= NE) continue;   // short last group
            const int* sst = reinterpret_cast<const int*>(slb + ce * COLD + ST_OFF) + tid + 1;
            s0[ce] = sst[0]; s1[ce] = sst[1]; s2[ce] = sst[2];
          }
          mbar_wait(&ready[bB.i], bB.par);
#pragma unroll
          for (int ce = 0; ce < GC; ++ce)
            if ((st - LA) * GC + ce < NE)
              gather_rs_first<L, NT, BR>(xr_l + ce * L * NT, xr_l + XYD + ce * L * NT, xr_s + (uint32_t)(ce * L * NT) * 8u,
                                       xr_s + (uint32_t)(XYD + ce * L * NT) * 8u, s0[ce], s1[ce], s2[ce], rank, xv[ce], yv[ce]);
        };
        if (LA > 0 && st >= LA) B1();
        if (st < NG) {   // ---- A: lottery masses ẋ, ẏ of this CTA's rows
          const double* sl = ring + (size_t)sA.i * SLOTD + tid;
          double* xw = xy + (size_t)bA.i * 2 * XYD + tid;
          mbar_wait(&full[sA.i], sA.par);
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            const int e = st * GC + ce;
            if (e >= NE) continue;
            double pd[L];
#pragma unroll
            for (int l = 0; l < L; ++l) pd[l] = sl[PD_OFF + (ce * L + l) * NT];
            const double om = sl[ce * COLD + FW_OM * NT], dco = sl[ce * COLD + FW_DCO * NT], Dn = sl[ce * COLD + FW_D * NT];
            pv[e] = sl[ce * COLD + FW_P * NT];
#pragma unroll
            for (int l = 0; l < L; ++l) {
              const double xd = fma(om, Dd[l][e], dco * pd[l]);
              xw[(ce * L + l) * NT] = xd;
              xw[XYD + (ce * L + l) * NT] = Dd[l][e] - xd;
              kacc[l] = fma(pd[l], Dn, kacc[l]);
            }
          }
          // this thread's masses of the group are written: the signalling warp tells the peers (see the backward sweep)
          if constexpr (LA == 0) named_bar_arrive(1, NT + 32); else named_bar_sync(1, NT + 32);
          if (st == 0 && t > 0) named_bar_sync(3, NT);   // every warp's partial of the previous period is in `red`
          if (st == 0 && t > 0 && tid < L && lane0 + tid < K) {   // the previous period's K̇D share of this CTA
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red[(((t - 1) & 1) * L + tid) * NW + w];
            dkdpart[((size_t)(lane0 + tid) * P + (t - 1)) * NC + rank] = s;
          }
          sA.next(S); bA.next(NB);
        }
        if (LA == 0) B1();
        if (st >= LA) {   // ---- B2: sums, and the rows with more than one source per range
          const int g = st - LA;
#pragma unroll
          for (int ce = 0; ce < GC; ++ce) {
            const int e = g * GC + ce;
            if (e >= NE) continue;
            double acc[L];
#pragma unroll
            for (int l = 0; l < L; ++l) {   // 0 + x0 + y0 + x1 + y1: the order of gather_row's predicated chain
              const int n1 = s1[ce] - s0[ce], n2 = s2[ce] - s1[ce];
              double a = n1 > 0 ? xv[ce][0][l] : 0.0;
              if (n2 > 0) a += yv[ce][0][l];
              if (n1 > 1) a += xv[ce][1][l];
              if (n2 > 1) a += yv[ce][1][l];
              acc[l] = a;
            }
            gather_rs_rest<L, NT, BR>(xr_l + ce * L * NT, xr_l + XYD + ce * L * NT, xr_s + (uint32_t)(ce * L * NT) * 8u,
                                      xr_s + (uint32_t)(XYD + ce * L * NT) * 8u, s0[ce], s1[ce], s2[ce], rank, lane, acc);
#pragma unroll
            for (int l = 0; l < L; ++l) Dd[l][e] = acc[l];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_local(&empty[sB.i]);
          sB.next(S); bB.next(NB);
        }
      }
      // ---- Markov mix (in place) and second aggregation term <p_t, Ḋ_t>
#pragma unroll
      for (int l = 0; l < L; ++l) {
        double d[NE];
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) {
          double s = 0.0;
#pragma unroll
          for (int e = 0; e < NE; ++e) s = fma(M.Pi[e][e2], Dd[l][e], s);
          d[e2] = s;
        }
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) { Dd[l][e2] = d[e2]; kacc[l] = fma(pv[e2], d[e2], kacc[l]); }
      }
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const double s = warp_sum(kacc[l]);
        if (lane == 0) red[((t & 1) * L + l) * NW + warp] = s;
      }
    }
    named_bar_sync(3, NT);
    if (tid < L && lane0 + tid < K && P > 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += red[(((P - 1) & 1) * L + tid) * NW + w];
      dkdpart[((size_t)(lane0 + tid) * P + (P - 1)) * NC + rank] = s;
    }
  }
  __syncwarp();
  cluster_sync_all();
}

// ======================================================================================
// Forward tangent sweep of ONE lane per cluster with a thread per (income state, row) instead of per row.
//
// In the latency shapes (one lane, a whole period per exchange) the kernel above gives a CTA NT/32 compute warps —
// one per scheduler at 64 rows — and each of them walks all n_e columns of its rows: ~585 dependent-latency
// instructions per warp and period (profiles/r02_notes.md: 4.2 M warp instructions for 299 periods on 8 CTAs, issue
// slots 17 % busy), i.e. 1.84 us per period whatever the hardware could overlap.  Here the n_e columns of a row go to
// n_e different warps (thread = (e, row), NE*NT threads + the producer and the signalling warp), so a warp's chain per period is one
// column (~70 instructions) and four to seven warps share a scheduler.  The price is the Markov mix: the post-lottery
// masses of a row meet through shared memory (one more named barrier per period) and every thread forms the mixed
// mass of its own income state.  Everything else — ring, producer warp, exchange buffers, hand-shake, DSMEM pulls,
// summation order within a column — is the kernel above with L = 1, GC = NE, LA = 0.
// smem: ring[S][ NE x (4 fields x NT doubles | NT+4 ints) | ṗ [NE][NT] ] | xy[2][2][NE][NT] | mix[NE][NT] |
//       red[2][NE*NT/32] | full[S] | empty[S] | ready[2]
// ======================================================================================
template <int NE, int NT>
constexpr size_t rs_fw_ce_smem(int S) {
  return (size_t)S * NE * rs_fw_col_bytes<NT, 1>() + (size_t)2 * 2 * NE * NT * 8 + (size_t)NE * NT * 8 + (size_t)2 * (NE * NT / 32) * 8 +
         (size_t)(2 * S + 2) * 8 + 128;
}
template <int NE, int NC, int NT>
__global__ void __launch_bounds__(NE * NT + 64, 1)
k_forward_tangent_rs_ce(const Consts<NE> M, const unsigned char* __restrict__ tape_rs, int K, int Kp, int S,
                        const int* __restrict__ thi, const double* __restrict__ zeros, const double* __restrict__ dpol,
                        int pd_rs, double* __restrict__ dkdpart, int release) {
  constexpr int LDA = NC * NT, NB = 2, NTC = NE * NT, NWC = NTC / 32;
  constexpr int COLB = (int)rs_fw_tape_col_bytes<NT>(), COLD = COLB / 8;
  constexpr int ST_OFF = FW_NF * NT, PD_OFF = NE * COLD, SLOTD = NE * COLD + NE * NT, XYD = NE * NT;
  static_assert(NT % 32 == 0, "a warp must not straddle two income states");
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster;
  const int P = M.P;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* xy = ring + (size_t)S * SLOTD;
  double* mix = xy + (size_t)NB * 2 * XYD;
  double* red = mix + XYD;
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * NWC);
  uint64_t* empty = full + S;
  uint64_t* ready = empty + S;
  const uint32_t xy_s = smem_u32(xy);
  const int pe = thi ? min(P, thi[lane0 / kThiGroup]) : P;   // ṗ is zero (and unwritten) from period pe on

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
    for (int b = 0; b < NB; ++b) mbar_init(&ready[b], NC);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NWC) {
    // ---- producer warp: the period's tape columns of this CTA's rows (one piece) and ṗ of the lane
    Cursor cs;
    for (int t = 0; t < P; ++t) {
      double* dst = ring + (size_t)cs.i * SLOTD;
      constexpr uint32_t PDB = (uint32_t)(NE * NT * 8);
      if (lane == 0) {
        if (t >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        mbar_expect_tx(&full[cs.i], (uint32_t)(NE * COLB) + PDB);
        bulk_g2s(dst, tape_rs + (((size_t)t * NC + rank) * NE) * COLB, (uint32_t)(NE * COLB), &full[cs.i]);
        if (t >= pe) {
          for (uint32_t o = 0; o < PDB; o += kZeroBytes)
            bulk_g2s(dst + PD_OFF + o / 8, zeros, min(PDB - o, (uint32_t)kZeroBytes), &full[cs.i]);
        } else if (pd_rs) {
          bulk_g2s(dst + PD_OFF, dpol + ((((size_t)t * ncl + cluster) * NC + rank) * NE) * NT, PDB, &full[cs.i]);
        }
      }
      __syncwarp();
      if (t < pe && !pd_rs)
        for (int i = lane; i < NE; i += 32)
          bulk_g2s(dst + PD_OFF + (size_t)i * NT, dpol + ((((size_t)t * NE + i) * Kp + lane0) * LDA + rank * NT), NT * 8, &full[cs.i]);
      cs.next(S);
    }
  } else if (warp == NWC + 1) {
    signal_groups<NC, NB>(ready, P, NTC + 32, lane, release);   // ---- signalling warp (see mbar_arrive_remote)
  } else {
    const int e = tid / NT, row = tid - e * NT;   // (warp-uniform e)
    double pic[NE];                               // Π[·, e]: this thread forms the mixed mass of income state e
#pragma unroll
    for (int e1 = 0; e1 < NE; ++e1) pic[e1] = M.Pi[e1][e];
    double Dd = 0.0;
    Cursor sl_c, b_c;
    for (int t = 0; t < P; ++t) {
      // ---- A: lottery masses ẋ, ẏ of this (row, income state)
      const double* sl = ring + (size_t)sl_c.i * SLOTD;
      double* xw = xy + (size_t)b_c.i * 2 * XYD;
      mbar_wait(&full[sl_c.i], sl_c.par);
      const double pd = sl[PD_OFF + e * NT + row];
      const double om = sl[e * COLD + FW_OM * NT + row], dco = sl[e * COLD + FW_DCO * NT + row];
      const double Dn = sl[e * COLD + FW_D * NT + row], pv = sl[e * COLD + FW_P * NT + row];
      const int* sst = reinterpret_cast<const int*>(sl + e * COLD + ST_OFF) + row + 1;
      const int s0 = sst[0], s1 = sst[1], s2 = sst[2];
      const double xd = fma(om, Dd, dco * pd);
      xw[e * NT + row] = xd;
      xw[XYD + e * NT + row] = Dd - xd;
      double kacc = pd * Dn;
      named_bar_arrive(1, NTC + 32);   // this thread's masses are written: the signalling warp tells the peers
      // ---- B: the sources of this destination row, pulled from their owners
      const double* xr_l = xw;
      const uint32_t xr_s = xy_s + (uint32_t)(b_c.i * 2 * XYD) * 8u;
      mbar_wait(&ready[b_c.i], b_c.par);
      double xv[2][1], yv[2][1];
      gather_rs_first<1, NT, false>(xr_l + e * NT, xr_l + XYD + e * NT, xr_s + (uint32_t)(e * NT) * 8u,
                                    xr_s + (uint32_t)(XYD + e * NT) * 8u, s0, s1, s2, rank, xv, yv);
      double acc[1];
      {   // 0 + x0 + y0 + x1 + y1: the order of gather_row's predicated chain
        const int n1 = s1 - s0, n2 = s2 - s1;
        double a = n1 > 0 ? xv[0][0] : 0.0;
        if (n2 > 0) a += yv[0][0];
        if (n1 > 1) a += xv[1][0];
        if (n2 > 1) a += yv[1][0];
        acc[0] = a;
      }
      gather_rs_rest<1, NT, false>(xr_l + e * NT, xr_l + XYD + e * NT, xr_s + (uint32_t)(e * NT) * 8u,
                                   xr_s + (uint32_t)(XYD + e * NT) * 8u, s0, s1, s2, rank, lane, acc);
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&empty[sl_c.i]);   // the tape slot goes back to the producer
      // ---- Markov mix across the income states of the row (through shared memory) and <p_t, Ḋ_t>
      mix[e * NT + row] = acc[0];
      named_bar_sync(2, NTC);
      if (t > 0 && tid == 0 && lane0 < K) {   // the previous period's K̇D share of this CTA (every warp wrote its
        double s = 0.0;                       // partial before it reached this barrier)
#pragma unroll
        for (int w = 0; w < NWC; ++w) s += red[((t - 1) & 1) * NWC + w];
        dkdpart[((size_t)lane0 * P + (t - 1)) * NC + rank] = s;
      }
      double d = 0.0;
#pragma unroll
      for (int e1 = 0; e1 < NE; ++e1) d = fma(pic[e1], mix[e1 * NT + row], d);
      Dd = d;
      kacc = fma(pv, d, kacc);
      const double s = warp_sum(kacc);
      if (lane == 0) red[(t & 1) * NWC + warp] = s;
      sl_c.next(S); b_c.next(NB);
    }
    named_bar_sync(3, NTC);
    if (tid == 0 && lane0 < K && P > 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NWC; ++w) s += red[((P - 1) & 1) * NWC + w];
      dkdpart[((size_t)lane0 * P + (P - 1)) * NC + rank] = s;
    }
  }
  __syncwarp();
  cluster_sync_all();
}

// ======================================================================================
// One-lane sweeps that PUSH instead of pull (k_backward_tangent_rs_push, k_forward_tangent_rs_push).
//
// The pull kernels above publish generic-proxy shared-memory writes to the cluster, which needs a cluster-scope
// release (~0.45 us per hand-shake) on top of the remote-load round trips.  Here an owner's block of exchanged values
// travels through the ASYNC proxy instead, the way TMA stores leave shared memory: the compute threads write their
// values to a local staging block and execute fence.proxy.async, a named barrier hands the block to a pushing warp,
// and that warp copies it with ONE cp.async.bulk.shared::cluster per destination CTA into the destination's
// full-width buffer, completing bytes on the destination's mbarrier (complete_tx).  Readers wait on their own
// mbarrier — the sanctioned completion mechanism for async-proxy writes, no fence — and then gather from LOCAL
// shared memory.  Only the CTAs that need a block get it: the range of source blocks a destination block reads in a
// period (lottery source ranges / interpolation knots, both known from the primal tape) is tabulated per
// (period, destination) by k_rs_ranges; everybody else receives a 16-byte token, so that every CTA still hears from
// every CTA once per period — that all-to-all is what makes the two-deep buffers safe (a CTA can only be one period
// ahead of any other).  Thread = (income state, row) as in k_forward_tangent_rs_ce; one lane per cluster, a whole
// period per exchange.
// ======================================================================================
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// range of source blocks each destination block reads, per period: out[t*NC + d] = lo | hi << 8 (lo > hi: none)
// forward: sources of the lottery ranges [s0, s2) of the block's rows; backward: the knots i, i+1 of its rows
static __global__ void k_rs_ranges(const unsigned char* __restrict__ tape_rs, int P, int NE, int NC, int NT, int n_a,
                                   int forward, int* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= P * NC) return;
  const int t = w / NC, d = w - t * NC;
  int lognt = 0;
  while ((1 << lognt) < NT) ++lognt;
  const size_t colb = forward ? (size_t)FW_NF * NT * 8 + (size_t)(NT + 4) * 4 : (size_t)BW_NF * NT * 8 + (size_t)NT * 4;
  const size_t into = forward ? (size_t)FW_NF * NT * 8 : (size_t)BW_NF * NT * 8;
  int lo = NC, hi = -1;
  for (int e = 0; e < NE; ++e) {
    const int* ints = reinterpret_cast<const int*>(tape_rs + (((size_t)t * NC + d) * NE + e) * colb + into);
    for (int r = lane; r < NT; r += 32) {
      if (d * NT + r >= n_a) continue;
      if (forward) {
        const int s0 = ints[r + 1], s2 = ints[r + 3];
        if (s2 > s0) { lo = min(lo, s0 >> lognt); hi = max(hi, (s2 - 1) >> lognt); }
      } else {
        const int i0 = ints[r];
        lo = min(lo, i0 >> lognt); hi = max(hi, (i0 + 1) >> lognt);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if (lane == 0) { lo = max(lo, 0); hi = min(hi, NC - 1); out[w] = lo > hi ? 1 : (lo | (hi << 8)); }
}
__device__ __forceinline__ bool rng_has(int rng, int s) { return (rng & 255) <= s && s <= ((rng >> 8) & 255); }
__device__ __forceinline__ int rng_count(int rng) { const int lo = rng & 255, hi = (rng >> 8) & 255; return hi >= lo ? hi - lo + 1 : 0; }

// the pushing warp's loop: per period, arm this CTA's barrier with what it will receive, wait for the local block,
// send it (or a token) to every CTA.  BLKB: bytes of a block.  tok: [2][NC] 16-byte slots.
template <int NC>
__device__ __forceinline__ void push_periods(const int* __restrict__ rng, int t_first, int t_step, int P, uint32_t rank,
                                             const double* stg, int stg_stride_d, double* full_buf, int full_stride_d, int blkd,
                                             uint32_t blkb, double* tok, uint64_t* xbar, int nthreads_bar, int lane) {
  const uint32_t full_s = smem_u32(full_buf), tok_s = smem_u32(tok), bar_s = smem_u32(xbar);
  int t = t_first;
  for (int it = 0; it < P; ++it, t += t_step) {
    const int b = it & 1;
    const int mine = __ldg(rng + (size_t)t * NC + rank);            // what this CTA receives this period
    const int theirs = lane < NC ? __ldg(rng + (size_t)t * NC + lane) : 1;
    if (lane == 0) mbar_expect_tx(&xbar[b], (uint32_t)rng_count(mine) * blkb + (uint32_t)(NC - rng_count(mine)) * 16u);
    asm volatile("bar.sync %0, %1;" ::"r"(1), "r"(nthreads_bar) : "memory");   // the block is staged (and proxy-fenced)
    if (lane < NC) {
      const uint32_t bar_r = map_to_cta(bar_s + 8u * (uint32_t)b, (uint32_t)lane);
      if (rng_has(theirs, (int)rank))
        bulk_s2c(map_to_cta(full_s + (uint32_t)(b * full_stride_d + (int)rank * blkd) * 8u, (uint32_t)lane), stg + (size_t)b * stg_stride_d, blkb, bar_r);
      else
        bulk_s2c(map_to_cta(tok_s + (uint32_t)((b * NC + (int)rank) * 2) * 8u, (uint32_t)lane), tok + (size_t)(2 * NC) * 2, 16u, bar_r);
    }
  }
}

template <int NE, int NT> constexpr size_t rs_push_ring_slot_fw() { return (size_t)NE * rs_fw_col_bytes<NT, 1>(); }
template <int NE, int NC, int NT>
constexpr size_t rs_fw_push_smem(int S) {   // ring | stg[2][2 NE NT] | full[2][NC][2 NE NT] | mix[NE NT] | red | tok[(2 NC + 1) x 16 B] | barriers
  return (size_t)S * rs_push_ring_slot_fw<NE, NT>() + (size_t)2 * 2 * NE * NT * 8 + (size_t)2 * NC * 2 * NE * NT * 8 + (size_t)NE * NT * 8 +
         (size_t)2 * (NE * NT / 32) * 8 + (size_t)(2 * NC + 1) * 16 + (size_t)(2 * S + 2) * 8 + 128;
}
template <int NE, int NC, int NT>
__global__ void __launch_bounds__(NE * NT + 64, 1)
k_forward_tangent_rs_push(const Consts<NE> M, const unsigned char* __restrict__ tape_rs, const int* __restrict__ rng, int K, int Kp,
                          int S, const int* __restrict__ thi, const double* __restrict__ zeros, const double* __restrict__ dpol,
                          int pd_rs, double* __restrict__ dkdpart) {
  constexpr int LDA = NC * NT, NTC = NE * NT, NWC = NTC / 32, LOGNT = ilog2c(NT);
  constexpr int COLB = (int)rs_fw_tape_col_bytes<NT>(), COLD = COLB / 8;
  constexpr int ST_OFF = FW_NF * NT, PD_OFF = NE * COLD, SLOTD = NE * COLD + NE * NT;
  constexpr int BLKD = 2 * NE * NT;                 // a block: x[NE][NT] | y[NE][NT]
  static_assert(NT % 32 == 0 && NC <= 32, "shape");
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster;
  const int P = M.P, n_a = M.n_a;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* stg = ring + (size_t)S * SLOTD;
  double* fullb = stg + 2 * BLKD;
  double* mix = fullb + (size_t)2 * NC * BLKD;
  double* red = mix + NTC;
  double* tok = red + 2 * NWC;                      // [2][NC] receive slots + one send slot, 16 bytes each
  uint64_t* full = reinterpret_cast<uint64_t*>(tok + (2 * NC + 1) * 2);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;
  const int pe = thi ? min(P, thi[lane0 / kThiGroup]) : P;
  (void)n_a;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
    mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < (2 * NC + 1) * 2) tok[tid] = 0.0;
  __syncthreads();
  cluster_sync_all();

  if (warp == NWC) {
    // ---- producer warp (as in k_forward_tangent_rs_ce)
    Cursor cs;
    for (int t = 0; t < P; ++t) {
      double* dst = ring + (size_t)cs.i * SLOTD;
      constexpr uint32_t PDB = (uint32_t)(NE * NT * 8);
      if (lane == 0) {
        if (t >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        mbar_expect_tx(&full[cs.i], (uint32_t)(NE * COLB) + PDB);
        bulk_g2s(dst, tape_rs + (((size_t)t * NC + rank) * NE) * COLB, (uint32_t)(NE * COLB), &full[cs.i]);
        if (t >= pe) {
          for (uint32_t o = 0; o < PDB; o += kZeroBytes)
            bulk_g2s(dst + PD_OFF + o / 8, zeros, min(PDB - o, (uint32_t)kZeroBytes), &full[cs.i]);
        } else if (pd_rs) {
          bulk_g2s(dst + PD_OFF, dpol + ((((size_t)t * ncl + cluster) * NC + rank) * NE) * NT, PDB, &full[cs.i]);
        }
      }
      __syncwarp();
      if (t < pe && !pd_rs)
        for (int i = lane; i < NE; i += 32)
          bulk_g2s(dst + PD_OFF + (size_t)i * NT, dpol + ((((size_t)t * NE + i) * Kp + lane0) * LDA + rank * NT), NT * 8, &full[cs.i]);
      cs.next(S);
    }
  } else if (warp == NWC + 1) {
    push_periods<NC>(rng, 0, 1, P, rank, stg, BLKD, fullb, NC * BLKD, BLKD, (uint32_t)BLKD * 8u, tok, xbar, NTC + 32, lane);
  } else {
    const int e = tid / NT, row = tid - e * NT;
    double pic[NE];
#pragma unroll
    for (int e1 = 0; e1 < NE; ++e1) pic[e1] = M.Pi[e1][e];
    double Dd = 0.0;
    Cursor sl_c;
    for (int t = 0; t < P; ++t) {
      const int b = t & 1;
      const double* sl = ring + (size_t)sl_c.i * SLOTD;
      mbar_wait(&full[sl_c.i], sl_c.par);
      const double pd = sl[PD_OFF + e * NT + row];
      const double om = sl[e * COLD + FW_OM * NT + row], dco = sl[e * COLD + FW_DCO * NT + row];
      const double Dn = sl[e * COLD + FW_D * NT + row], pv = sl[e * COLD + FW_P * NT + row];
      const int* sst = reinterpret_cast<const int*>(sl + e * COLD + ST_OFF) + row + 1;
      const int s0 = sst[0], s1 = sst[1], s2 = sst[2];
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&empty[sl_c.i]);   // everything this warp needs of the slot is in registers
      const double xd = fma(om, Dd, dco * pd);
      stg[b * BLKD + e * NT + row] = xd;
      stg[b * BLKD + NTC + e * NT + row] = Dd - xd;
      double kacc = pd * Dn;
      fence_proxy_async_smem();                 // this thread's staged values, before the pushing warp's bulk copies read them
      named_bar_arrive(1, NTC + 32);
      // ---- the sources of this destination row, from the blocks the peers pushed into this CTA
      mbar_wait_cluster(&xbar[b], (uint32_t)((t >> 1) & 1));
      const double* fb = fullb + (size_t)b * NC * BLKD + e * NT;
      auto X = [&](int src) { return fb[(src >> LOGNT) * BLKD + (src & (NT - 1))]; };
      auto Y = [&](int src) { return fb[(src >> LOGNT) * BLKD + NTC + (src & (NT - 1))]; };
      const int n1 = s1 - s0, n2 = s2 - s1;
      double acc = 0.0;                         // 0 + x0 + y0 + x1 + y1 + the rest of x, then of y: gather_row's order
      if (n1 > 0) acc = X(s0);
      if (n2 > 0) acc += Y(s1);
      if (n1 > 1) acc += X(s0 + 1);
      if (n2 > 1) acc += Y(s1 + 1);
      const int mx = max(n1, n2) - 2;
      if (__any_sync(0xffffffffu, mx > 0)) {
        constexpr int kSerial = 8;
        if (mx > 0 && mx <= kSerial) {
          for (int q = s0 + 2; q < s1; ++q) acc += X(q);
          for (int q = s1 + 2; q < s2; ++q) acc += Y(q);
        }
        unsigned bal = __ballot_sync(0xffffffffu, mx > kSerial);
        while (bal) {                            // long ranges (the mass at the borrowing constraint): the whole warp
          const int src = __ffs(bal) - 1;
          bal &= bal - 1;
          const int b0 = __shfl_sync(0xffffffffu, s0, src), b1 = __shfl_sync(0xffffffffu, s1, src), b2 = __shfl_sync(0xffffffffu, s2, src);
          double v = 0.0;
          for (int q = b0 + 2 + lane; q < b1; q += 32) v += X(q);
          for (int q = b1 + 2 + lane; q < b2; q += 32) v += Y(q);
          v = warp_sum(v);
          if (lane == src) acc += v;
        }
      }
      // ---- Markov mix across the income states of the row and <p_t, Ḋ_t>
      mix[e * NT + row] = acc;
      named_bar_sync(2, NTC);
      if (t > 0 && tid == 0 && lane0 < K) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWC; ++w) s += red[((t - 1) & 1) * NWC + w];
        dkdpart[((size_t)lane0 * P + (t - 1)) * NC + rank] = s;
      }
      double d = 0.0;
#pragma unroll
      for (int e1 = 0; e1 < NE; ++e1) d = fma(pic[e1], mix[e1 * NT + row], d);
      Dd = d;
      kacc = fma(pv, d, kacc);
      const double s = warp_sum(kacc);
      if (lane == 0) red[(t & 1) * NWC + warp] = s;
      sl_c.next(S);
    }
    named_bar_sync(3, NTC);
    if (tid == 0 && lane0 < K && P > 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NWC; ++w) s += red[((P - 1) & 1) * NWC + w];
      dkdpart[((size_t)lane0 * P + (P - 1)) * NC + rank] = s;
    }
  }
  __syncwarp();
  cluster_sync_all();
}

template <int NE, int NC, int NT>
constexpr size_t rs_bw_push_smem(int S) {   // ring | stg[2][NE NT] | full[2][NC][NE NT] | mixv[NE NT] | tok | barriers
  return (size_t)S * NE * rs_bw_col_bytes<NT>() + (size_t)2 * NE * NT * 8 + (size_t)2 * NC * NE * NT * 8 + (size_t)NE * NT * 8 +
         (size_t)(2 * NC + 1) * 16 + (size_t)(2 * S + 2) * 8 + 128;
}
template <int NE, int NC, int NT>
__global__ void __launch_bounds__(NE * NT + 64, 1)
k_backward_tangent_rs_push(const Consts<NE> M, const Tape tp, const unsigned char* __restrict__ tape_rs, const int* __restrict__ rng,
                           int K, int S, const int* __restrict__ thi, const double* __restrict__ dr, const double* __restrict__ dw,
                           double* __restrict__ dpol) {
  constexpr int NTC = NE * NT, NWC = NTC / 32, LOGNT = ilog2c(NT);
  constexpr int COLB = (int)rs_bw_col_bytes<NT>(), COLD = COLB / 8, SLOTD = NE * COLD;
  constexpr int BLKD = NE * NT;                     // a block: k̇[NE][NT]
  static_assert(NT % 32 == 0 && NC <= 32, "shape");
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster;
  const int Pfull = M.P, P = thi ? min(M.P, thi[lane0 / kThiGroup]) : M.P, n_a = M.n_a;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* stg = ring + (size_t)S * SLOTD;
  double* fullb = stg + 2 * BLKD;
  double* mixv = fullb + (size_t)2 * NC * BLKD;
  double* tok = mixv + NTC;
  uint64_t* full = reinterpret_cast<uint64_t*>(tok + (2 * NC + 1) * 2);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
    mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < (2 * NC + 1) * 2) tok[tid] = 0.0;
  __syncthreads();
  cluster_sync_all();

  if (warp == NWC) {
    if (lane == 0) {   // ---- producer: this CTA's rows of the period's tape, one piece
      Cursor cs;
      for (int it = 0; it < P; ++it) {
        const int t = P - 1 - it;
        if (it >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        mbar_expect_tx(&full[cs.i], (uint32_t)(NE * COLB));
        bulk_g2s(ring + (size_t)cs.i * SLOTD, tape_rs + (((size_t)t * NC + rank) * NE) * COLB, (uint32_t)(NE * COLB), &full[cs.i]);
        cs.next(S);
      }
    }
  } else if (warp == NWC + 1) {
    push_periods<NC>(rng, P - 1, -1, P, rank, stg, BLKD, fullb, NC * BLKD, BLKD, (uint32_t)BLKD * 8u, tok, xbar, NTC + 32, lane);
  } else {
    const int e = tid / NT, row = tid - e * NT;
    const bool live = (int)rank * NT + row < n_a;
    double pir[NE];                                   // Π[e, ·]
#pragma unroll
    for (int e2 = 0; e2 < NE; ++e2) pir[e2] = M.Pi[e][e2];
    const double ze = M.z[e];
    const bool on = lane0 < K && P > 0;
    double Vd = 0.0;
    double rho = __ldg(tp.rho + (P > 0 ? P - 1 : 0));
    double drn = on ? __ldg(dr + (size_t)lane0 * Pfull + P - 1) : 0.0, dwn = on ? __ldg(dw + (size_t)lane0 * Pfull + P - 1) : 0.0;
    Cursor sl_c;
    for (int it = 0; it < P; ++it) {
      const int t = P - 1 - it, b = it & 1;
      const double rho_t = rho, drl = drn, dwl = dwn;
      if (t > 0) {
        rho = __ldg(tp.rho + t - 1);
        drn = on ? __ldg(dr + (size_t)lane0 * Pfull + t - 1) : 0.0;
        dwn = on ? __ldg(dw + (size_t)lane0 * Pfull + t - 1) : 0.0;
      }
      // ---- ĖV of this income state: the row's V̇ of all income states meet in shared memory
      mixv[e * NT + row] = Vd;
      named_bar_sync(3, NTC);
      double ev = 0.0;
#pragma unroll
      for (int e2 = 0; e2 < NE; ++e2) ev = fma(pir[e2], mixv[e2 * NT + row], ev);
      // ---- k̇ of this (row, income state), staged for the pushing warp
      const double* sl = ring + (size_t)sl_c.i * SLOTD + e * COLD + row;
      mbar_wait(&full[sl_c.i], sl_c.par);
      const double a1 = sl[BW_A1 * NT], kr = sl[BW_KR * NT];
      const double cA = sl[BW_CA * NT], cB = sl[BW_CB * NT], E1 = sl[BW_E1 * NT], vf = sl[BW_VF * NT];
      const int i0 = reinterpret_cast<const int*>(sl - row + BW_NF * NT)[row];
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&empty[sl_c.i]);
      stg[b * BLKD + e * NT + row] = fma(a1, ev, fma(kr, drl, -(rho_t * ze) * dwl));
      fence_proxy_async_smem();
      named_bar_arrive(1, NTC + 32);
      // ---- ṗ and V̇ from k̇ at the two knots (pushed into this CTA by their owners)
      mbar_wait_cluster(&xbar[b], (uint32_t)((it >> 1) & 1));
      const double* fb = fullb + (size_t)b * NC * BLKD + e * NT;
      const int i1 = i0 + 1;
      const double k0 = live ? fb[(i0 >> LOGNT) * BLKD + (i0 & (NT - 1))] : 0.0;
      const double k1 = live ? fb[(i1 >> LOGNT) * BLKD + (i1 & (NT - 1))] : 0.0;
      const double pd = fma(cA, k0, cB * k1);
      __stcs(dpol + (((size_t)t * ncl + cluster) * NC + rank) * (size_t)(NE * NT) + e * NT + row, pd);
      Vd = fma(vf, fma(ze, dwl, -pd), E1 * drl);
      sl_c.next(S);
    }
  }
  __syncwarp();
  cluster_sync_all();
}

// ======================================================================================
// One-lane sweeps whose threads push their OWN values (k_backward_tangent_rs_st, k_forward_tangent_rs_st).
//
// The bulk-copy push above still pays fence.proxy.async + a named barrier + the copy engine's issue latency per period,
// and it moves whole blocks (3.5 / 7 KB per destination) through the ~20 B/clk DSMEM port when a neighbour needs a few
// boundary rows.  But where a value has to go is known from the primal tape: in the forward sweep the mass x of source
// q feeds exactly one destination row (the one whose x-range holds q) and y exactly one; in the backward sweep k̇ of
// row g is read by the rows whose knots are g-1 or g (their owners: a bit mask).  k_rs_st_mark_* tabulate that per
// (period, income state, source row) once per linearisation, together with the number of values every CTA receives per
// period.  In the sweep each compute thread sends its value(s) with st.async (8 bytes, complete_tx on the destination's
// mbarrier) straight from its registers into the destination's full-width buffer — no staging, no proxy fence, no
// pushing warp — and the readers wait on their own mbarrier for (values + NC tokens) x 8 bytes and gather locally.
// The NC tokens (one from every CTA of the cluster, sent by threads that have passed the period's CTA barrier) keep
// any CTA at most one period ahead of any other, which is what makes the two-deep buffers and the two-phase barrier
// safe.  Arithmetic and summation order are those of k_forward_tangent_rs_ce / k_backward_tangent_rs_push.
// ======================================================================================
// backward: info[(t*NC+s)*NE+e][r] |= 1 << owner of every live row whose knots i, i+1 include row s*NT+r
static __global__ void k_rs_st_mark_bw(const unsigned char* __restrict__ tape_rs, int P, int NE, int NC, int NT, int n_a,
                                       int* __restrict__ info) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P * NC * NE * NT) return;
  const int r = (int)(i % NT);
  const size_t col = i / NT;                         // (t*NC + d)*NE + e
  const int e = (int)(col % NE);
  const size_t td = col / NE;
  const int d = (int)(td % NC);
  const size_t t = td / NC;
  if (d * NT + r >= n_a) return;
  const size_t colb = (size_t)BW_NF * NT * 8 + (size_t)NT * 4;
  const int i0 = reinterpret_cast<const int*>(tape_rs + col * colb + (size_t)BW_NF * NT * 8)[r];
  for (int k = i0; k <= i0 + 1; ++k)
    if (k >= 0 && k < NC * NT) atomicOr(&info[((t * NC + k / NT) * NE + e) * NT + k % NT], 1 << d);
}
// cnt[t*NC+d] += number of set bits d among the sources of period t (NC*NE*NT is a multiple of 32: a warp has one t)
static __global__ void k_rs_st_count_bw(const int* __restrict__ info, int P, int NE, int NC, int NT, int* __restrict__ cnt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per_t = (size_t)NC * NE * NT, total = (size_t)P * per_t;
  const int mask = i < total ? info[i] : 0;
  const size_t t = (i - (threadIdx.x & 31)) / per_t;
  for (int d = 0; d < NC; ++d) {
    const unsigned bal = __ballot_sync(0xffffffffu, (mask >> d) & 1);
    if ((threadIdx.x & 31) == 0 && bal && t < (size_t)P) atomicAdd(&cnt[t * NC + d], __popc(bal));
  }
}
// forward: byte 0 of info[...source...] = owner of the row whose x-range holds the source, byte 1 likewise for y
// (bytes preset to 0xFF: nobody); cnt[t*NC+d] += sources of the rows of CTA d
static __global__ void k_rs_st_mark_fw(const unsigned char* __restrict__ tape_rs, int P, int NE, int NC, int NT, int n_a,
                                       int* __restrict__ info, int* __restrict__ cnt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P * NC * NE * NT) return;
  const int r = (int)(i % NT);
  const size_t col = i / NT;
  const int e = (int)(col % NE);
  const size_t td = col / NE;
  const int d = (int)(td % NC);
  const size_t t = td / NC;
  if (d * NT + r >= n_a) return;
  const size_t colb = (size_t)FW_NF * NT * 8 + (size_t)(NT + 4) * 4;
  const int* ints = reinterpret_cast<const int*>(tape_rs + col * colb + (size_t)FW_NF * NT * 8) + r + 1;
  const int s0 = max(ints[0], 0), s1 = ints[1], s2 = min(ints[2], NC * NT);
  unsigned char* b = reinterpret_cast<unsigned char*>(info);
  for (int q = s0; q < s1; ++q) b[(((t * NC + q / NT) * NE + e) * NT + q % NT) * 4] = (unsigned char)d;
  for (int q = s1; q < s2; ++q) b[(((t * NC + q / NT) * NE + e) * NT + q % NT) * 4 + 1] = (unsigned char)d;
  if (s2 > s0) atomicAdd(&cnt[t * NC + d], s2 - s0);
}

template <int NE, int NC, int NT> constexpr size_t rs_fw_st_slot() { return (size_t)NE * rs_fw_col_bytes<NT, 1>() + (size_t)NE * NT * 4; }
template <int NE, int NC, int NT>
constexpr size_t rs_fw_st_smem(int S) {   // ring | xy[2][2][NE][LDA] | mix[NE NT] | red | tok[2][NC] | barriers
  return (size_t)S * rs_fw_st_slot<NE, NC, NT>() + (size_t)2 * 2 * NE * NC * NT * 8 + (size_t)NE * NT * 8 + (size_t)2 * (NE * NT / 32) * 8 +
         (size_t)2 * NC * 8 + (size_t)(2 * S + 2) * 8 + 128;
}
template <int NE, int NC, int NT>
__global__ void __launch_bounds__(NE * NT + 32, 1)
k_forward_tangent_rs_st(const Consts<NE> M, const unsigned char* __restrict__ tape_rs, const int* __restrict__ info,
                        const int* __restrict__ cnt, const int* __restrict__ status, int K, int Kp, int S, const int* __restrict__ thi,
                        const double* __restrict__ zeros, const double* __restrict__ dpol, int pd_rs, double* __restrict__ dkdpart) {
  constexpr int LDA = NC * NT, NTC = NE * NT, NWC = NTC / 32;
  // a primal sweep that reported an error (a policy that is not monotone in a, ...) leaves a tape whose destinations and
  // counts do not add up: every CTA of every cluster skips the sweep, the caller reads the status after the pass
  if (status[0] != 0) return;
  constexpr int COLB = (int)rs_fw_tape_col_bytes<NT>(), COLD = COLB / 8;
  constexpr int ST_OFF = FW_NF * NT, PD_OFF = NE * COLD, IN_OFF = PD_OFF + NE * NT, SLOTD = IN_OFF + NE * NT / 2;
  constexpr int XYD = NE * LDA;                     // one of x, y of one buffer
  static_assert(NT % 32 == 0 && NC <= 32 && (NE * NT) % 4 == 0, "shape");
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster;
  const int P = M.P;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* xy = ring + (size_t)S * SLOTD;
  double* mix = xy + (size_t)4 * XYD;
  double* red = mix + NTC;
  double* tok = red + 2 * NWC;
  uint64_t* full = reinterpret_cast<uint64_t*>(tok + 2 * NC);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;
  const int pe = thi ? min(P, thi[lane0 / kThiGroup]) : P;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
    mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NWC) {
    // ---- producer warp: tape columns of this CTA's rows, ṗ of the lane, destinations of its sources
    Cursor cs;
    for (int t = 0; t < P; ++t) {
      double* dst = ring + (size_t)cs.i * SLOTD;
      constexpr uint32_t PDB = (uint32_t)(NE * NT * 8), INB = (uint32_t)(NE * NT * 4);
      if (lane == 0) {
        if (t >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        mbar_expect_tx(&full[cs.i], (uint32_t)(NE * COLB) + PDB + INB);
        bulk_g2s(dst, tape_rs + (((size_t)t * NC + rank) * NE) * COLB, (uint32_t)(NE * COLB), &full[cs.i]);
        bulk_g2s(dst + IN_OFF, info + ((size_t)t * NC + rank) * NTC, INB, &full[cs.i]);
        if (t >= pe) {
          for (uint32_t o = 0; o < PDB; o += kZeroBytes)
            bulk_g2s(dst + PD_OFF + o / 8, zeros, min(PDB - o, (uint32_t)kZeroBytes), &full[cs.i]);
        } else if (pd_rs) {
          bulk_g2s(dst + PD_OFF, dpol + ((((size_t)t * ncl + cluster) * NC + rank) * NE) * NT, PDB, &full[cs.i]);
        }
      }
      __syncwarp();
      if (t < pe && !pd_rs)
        for (int i = lane; i < NE; i += 32)
          bulk_g2s(dst + PD_OFF + (size_t)i * NT, dpol + ((((size_t)t * NE + i) * Kp + lane0) * LDA + rank * NT), NT * 8, &full[cs.i]);
      cs.next(S);
    }
  } else {
    const int e = tid / NT, row = tid - e * NT;
    double pic[NE];
#pragma unroll
    for (int e1 = 0; e1 < NE; ++e1) pic[e1] = M.Pi[e1][e];
    const uint32_t xy_s = smem_u32(xy), xbar_s = smem_u32(xbar);
    const uint32_t mine_off = (uint32_t)(e * LDA + (int)rank * NT + row) * 8u;   // this source's place in a full-width array
    const uint32_t tok_dst = tid < NC ? map_to_cta(smem_u32(tok) + rank * 8u, (uint32_t)tid) : 0u;
    const uint32_t tok_bar = tid < NC ? map_to_cta(xbar_s, (uint32_t)tid) : 0u;
    int cnt_next = (tid == 0 && P > 0) ? __ldg(cnt + rank) : 0;
    double Dd = 0.0;
    Cursor sl_c;
    // Software pipeline: the tape values of period t+1 are fetched (and the slot handed back) and the <p, Ḋ> partial of
    // period t-1 is reduced while the pushes of period t travel — the warp would only wait there.
    double pd = 0.0, om = 0.0, dco = 0.0, Dn = 0.0, pv = 0.0, kprev = 0.0;
    int s0 = 0, s1 = 0, s2 = 0, inf = -1;
    auto fetch = [&](double& pd_, double& om_, double& dco_, double& Dn_, double& pv_, int& s0_, int& s1_, int& s2_, int& inf_) {
      const double* sl = ring + (size_t)sl_c.i * SLOTD;
      mbar_wait(&full[sl_c.i], sl_c.par);
      pd_ = sl[PD_OFF + e * NT + row];
      om_ = sl[e * COLD + FW_OM * NT + row]; dco_ = sl[e * COLD + FW_DCO * NT + row];
      Dn_ = sl[e * COLD + FW_D * NT + row]; pv_ = sl[e * COLD + FW_P * NT + row];
      const int* sst = reinterpret_cast<const int*>(sl + e * COLD + ST_OFF) + row + 1;
      s0_ = sst[0]; s1_ = sst[1]; s2_ = sst[2];
      inf_ = reinterpret_cast<const int*>(sl + IN_OFF)[e * NT + row];
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&empty[sl_c.i]);   // everything this warp needs of the slot is in registers
      sl_c.next(S);
    };
    if (P > 0) fetch(pd, om, dco, Dn, pv, s0, s1, s2, inf);
    for (int t = 0; t < P; ++t) {
      const int b = t & 1;
      if (tid == 0) {
        mbar_expect_tx(&xbar[b], (uint32_t)(cnt_next + NC) * 8u);
        if (t + 1 < P) cnt_next = __ldg(cnt + (size_t)(t + 1) * NC + rank);
      }
      // ---- the masses of this source go to the owners of their destination rows
      const double xd = fma(om, Dd, dco * pd);
      const uint32_t dx = (uint32_t)inf & 255u, dy = ((uint32_t)inf >> 8) & 255u;
      const uint32_t xoff = xy_s + (uint32_t)(b * 2 * XYD) * 8u + mine_off;
      if (dx < (uint32_t)NC) st_async_f64(map_to_cta(xoff, dx), xd, map_to_cta(xbar_s + 8u * (uint32_t)b, dx));
      if (dy < (uint32_t)NC) st_async_f64(map_to_cta(xoff + (uint32_t)XYD * 8u, dy), Dd - xd, map_to_cta(xbar_s + 8u * (uint32_t)b, dy));
      if (tid < NC) st_async_f64(tok_dst + (uint32_t)(b * NC) * 8u, 0.0, tok_bar + 8u * (uint32_t)b);
      double kacc = pd * Dn;
      // ---- while they travel: last period's <p, Ḋ> partial of the warp, next period's tape values
      if (t > 0) {
        const double s = warp_sum(kprev);
        if (lane == 0) red[((t - 1) & 1) * NWC + warp] = s;
      }
      const int c0 = s0, c1 = s1, c2 = s2;
      const double pvc = pv;
      if (t + 1 < P) fetch(pd, om, dco, Dn, pv, s0, s1, s2, inf);
      // ---- the sources of this destination row, now in this CTA's buffer
      mbar_wait_cluster(&xbar[b], (uint32_t)((t >> 1) & 1));
      const double* xb = xy + (size_t)(b * 2) * XYD + e * LDA;
      const double* yb = xb + XYD;
      const int n1 = c1 - c0, n2 = c2 - c1;
      double acc = 0.0;                         // 0 + x0 + y0 + x1 + y1 + the rest of x, then of y: gather_row's order
      if (n1 > 0) acc = xb[c0];
      if (n2 > 0) acc += yb[c1];
      if (n1 > 1) acc += xb[c0 + 1];
      if (n2 > 1) acc += yb[c1 + 1];
      const int mx = max(n1, n2) - 2;
      if (__any_sync(0xffffffffu, mx > 0)) {
        constexpr int kSerial = 8;
        if (mx > 0 && mx <= kSerial) {
          for (int q = c0 + 2; q < c1; ++q) acc += xb[q];
          for (int q = c1 + 2; q < c2; ++q) acc += yb[q];
        }
        unsigned bal = __ballot_sync(0xffffffffu, mx > kSerial);
        while (bal) {                            // long ranges (the mass at the borrowing constraint): the whole warp
          const int src = __ffs(bal) - 1;
          bal &= bal - 1;
          const int b0 = __shfl_sync(0xffffffffu, c0, src), b1 = __shfl_sync(0xffffffffu, c1, src), b2 = __shfl_sync(0xffffffffu, c2, src);
          double v = 0.0;
          for (int q = b0 + 2 + lane; q < b1; q += 32) v += xb[q];
          for (int q = b1 + 2 + lane; q < b2; q += 32) v += yb[q];
          v = warp_sum(v);
          if (lane == src) acc += v;
        }
      }
      // ---- Markov mix across the income states of the row and <p_t, Ḋ_t>
      mix[e * NT + row] = acc;
      named_bar_sync(2, NTC);
      if (t > 0 && tid == 0 && lane0 < K) {   // (every warp wrote its partial of period t-1 before this barrier)
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWC; ++w) s += red[((t - 1) & 1) * NWC + w];
        dkdpart[((size_t)lane0 * P + (t - 1)) * NC + rank] = s;
      }
      double d = 0.0;
#pragma unroll
      for (int e1 = 0; e1 < NE; ++e1) d = fma(pic[e1], mix[e1 * NT + row], d);
      Dd = d;
      kprev = fma(pvc, d, kacc);
    }
    if (P > 0) {
      const double s = warp_sum(kprev);
      if (lane == 0) red[((P - 1) & 1) * NWC + warp] = s;
    }
    named_bar_sync(3, NTC);
    if (tid == 0 && lane0 < K && P > 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NWC; ++w) s += red[((P - 1) & 1) * NWC + w];
      dkdpart[((size_t)lane0 * P + (P - 1)) * NC + rank] = s;
    }
  }
  __syncwarp();
  cluster_sync_all();
}

template <int NE, int NC, int NT> constexpr size_t rs_bw_st_slot() { return (size_t)NE * rs_bw_col_bytes<NT>() + (size_t)NE * NT * 4; }
template <int NE, int NC, int NT>
constexpr size_t rs_bw_st_smem(int S) {   // ring | kd[2][NE][LDA] | mixv[NE NT] | tok[2][NC] | barriers
  return (size_t)S * rs_bw_st_slot<NE, NC, NT>() + (size_t)2 * NE * NC * NT * 8 + (size_t)NE * NT * 8 + (size_t)2 * NC * 8 +
         (size_t)(2 * S + 2) * 8 + 128;
}
template <int NE, int NC, int NT>
__global__ void __launch_bounds__(NE * NT + 32, 1)
k_backward_tangent_rs_st(const Consts<NE> M, const Tape tp, const unsigned char* __restrict__ tape_rs, const int* __restrict__ info,
                         const int* __restrict__ cnt, const int* __restrict__ status, int K, int S, const int* __restrict__ thi,
                         const double* __restrict__ dr, const double* __restrict__ dw, double* __restrict__ dpol) {
  constexpr int LDA = NC * NT, NTC = NE * NT, NWC = NTC / 32;
  if (status[0] != 0) return;   // (see k_forward_tangent_rs_st)
  constexpr int COLB = (int)rs_bw_col_bytes<NT>(), COLD = COLB / 8, IN_OFF = NE * COLD, SLOTD = IN_OFF + NE * NT / 2;
  constexpr int KD = NE * LDA;
  static_assert(NT % 32 == 0 && NC <= 32 && (NE * NT) % 4 == 0, "shape");
  extern __shared__ __align__(128) unsigned char smem_rs[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / NC, ncl = gridDim.x / NC;
  const int lane0 = cluster;
  const int Pfull = M.P, P = thi ? min(M.P, thi[lane0 / kThiGroup]) : M.P, n_a = M.n_a;
  double* ring = reinterpret_cast<double*>(smem_rs);
  double* kdb = ring + (size_t)S * SLOTD;
  double* mixv = kdb + (size_t)2 * KD;
  double* tok = mixv + NTC;
  uint64_t* full = reinterpret_cast<uint64_t*>(tok + 2 * NC);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); }
    mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NWC) {
    if (lane == 0) {   // ---- producer: this CTA's rows of the period's tape and the readers of its rows
      Cursor cs;
      constexpr uint32_t INB = (uint32_t)(NE * NT * 4);
      for (int it = 0; it < P; ++it) {
        const int t = P - 1 - it;
        if (it >= S) mbar_wait(&empty[cs.i], cs.par ^ 1);
        mbar_expect_tx(&full[cs.i], (uint32_t)(NE * COLB) + INB);
        bulk_g2s(ring + (size_t)cs.i * SLOTD, tape_rs + (((size_t)t * NC + rank) * NE) * COLB, (uint32_t)(NE * COLB), &full[cs.i]);
        bulk_g2s(ring + (size_t)cs.i * SLOTD + IN_OFF, info + ((size_t)t * NC + rank) * NTC, INB, &full[cs.i]);
        cs.next(S);
      }
    }
  } else {
    const int e = tid / NT, row = tid - e * NT;
    const bool live = (int)rank * NT + row < n_a;
    double pir[NE];                                   // Π[e, ·]
#pragma unroll
    for (int e2 = 0; e2 < NE; ++e2) pir[e2] = M.Pi[e][e2];
    const double ze = M.z[e];
    const bool on = lane0 < K && P > 0;
    const uint32_t kd_s = smem_u32(kdb), xbar_s = smem_u32(xbar);
    const uint32_t mine_off = (uint32_t)(e * LDA + (int)rank * NT + row) * 8u;
    const uint32_t tok_dst = tid < NC ? map_to_cta(smem_u32(tok) + rank * 8u, (uint32_t)tid) : 0u;
    const uint32_t tok_bar = tid < NC ? map_to_cta(xbar_s, (uint32_t)tid) : 0u;
    int cnt_next = (tid == 0 && P > 0) ? __ldg(cnt + (size_t)(P - 1) * NC + rank) : 0;
    double Vd = 0.0;
    double rho = __ldg(tp.rho + (P > 0 ? P - 1 : 0));
    double drn = on ? __ldg(dr + (size_t)lane0 * Pfull + P - 1) : 0.0, dwn = on ? __ldg(dw + (size_t)lane0 * Pfull + P - 1) : 0.0;
    Cursor sl_c;
    // (software pipeline as in k_forward_tangent_rs_st: next period's tape values are fetched while the pushes travel)
    double a1 = 0.0, kr = 0.0, cA = 0.0, cB = 0.0, E1 = 0.0, vf = 0.0;
    int i0 = 0;
    unsigned mk = 0u;
    auto fetch = [&]() {
      const double* sl = ring + (size_t)sl_c.i * SLOTD + e * COLD + row;
      mbar_wait(&full[sl_c.i], sl_c.par);
      a1 = sl[BW_A1 * NT]; kr = sl[BW_KR * NT];
      cA = sl[BW_CA * NT]; cB = sl[BW_CB * NT]; E1 = sl[BW_E1 * NT]; vf = sl[BW_VF * NT];
      i0 = reinterpret_cast<const int*>(sl - row + BW_NF * NT)[row];
      mk = (unsigned)reinterpret_cast<const int*>(ring + (size_t)sl_c.i * SLOTD + IN_OFF)[e * NT + row];
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&empty[sl_c.i]);
      sl_c.next(S);
    };
    if (P > 0) fetch();
    for (int it = 0; it < P; ++it) {
      const int t = P - 1 - it, b = it & 1;
      const double rho_t = rho, drl = drn, dwl = dwn;
      if (t > 0) {
        rho = __ldg(tp.rho + t - 1);
        drn = on ? __ldg(dr + (size_t)lane0 * Pfull + t - 1) : 0.0;
        dwn = on ? __ldg(dw + (size_t)lane0 * Pfull + t - 1) : 0.0;
      }
      // ---- ĖV of this income state: the row's V̇ of all income states meet in shared memory
      mixv[e * NT + row] = Vd;
      named_bar_sync(3, NTC);          // (every thread of the CTA is through the previous period's reads from here on)
      if (tid == 0) {
        mbar_expect_tx(&xbar[b], (uint32_t)(cnt_next + NC) * 8u);
        if (t > 0) cnt_next = __ldg(cnt + (size_t)(t - 1) * NC + rank);
      }
      double ev = 0.0;
#pragma unroll
      for (int e2 = 0; e2 < NE; ++e2) ev = fma(pir[e2], mixv[e2 * NT + row], ev);
      // ---- k̇ of this (row, income state), sent to the owners of the rows that interpolate on it
      const double kd = fma(a1, ev, fma(kr, drl, -(rho_t * ze) * dwl));
      const uint32_t koff = kd_s + (uint32_t)(b * KD) * 8u + mine_off;
      unsigned m = mk;
      while (m) {
        const uint32_t d = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        st_async_f64(map_to_cta(koff, d), kd, map_to_cta(xbar_s + 8u * (uint32_t)b, d));
      }
      if (tid < NC) st_async_f64(tok_dst + (uint32_t)(b * NC) * 8u, 0.0, tok_bar + 8u * (uint32_t)b);
      const double cAc = cA, cBc = cB, E1c = E1, vfc = vf;
      const int i0c = i0;
      if (it + 1 < P) fetch();
      // ---- ṗ and V̇ from k̇ at the two knots
      mbar_wait_cluster(&xbar[b], (uint32_t)((it >> 1) & 1));
      const double* kb = kdb + (size_t)b * KD + e * LDA;
      const double k0 = live ? kb[i0c] : 0.0;
      const double k1 = live ? kb[i0c + 1] : 0.0;
      const double pd = fma(cAc, k0, cBc * k1);
      __stcs(dpol + (((size_t)t * ncl + cluster) * NC + rank) * (size_t)(NE * NT) + e * NT + row, pd);
      Vd = fma(vfc, fma(ze, dwl, -pd), E1c * drl);
    }
  }
  __syncwarp();
  cluster_sync_all();
}

// ======================================================================================
// Row-block-major copies of the primal tape for the kernels above (see the header comment).
// One thread per (column, row): the row's coefficients and index / range start.
// ======================================================================================
static __global__ void k_tape_rowblocks_bw(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int ncols,
                                    int NE, int LDA, int NT) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncols * LDA) return;
  const int col = (int)(i / LDA), row = (int)(i - (size_t)col * LDA), t = col / NE, e = col - t * NE;
  const int NC = LDA / NT, rank = row / NT, r = row - rank * NT;
  const unsigned char* s = src + (size_t)col * ((size_t)BW_NF * LDA * 8 + (size_t)LDA * 4);
  unsigned char* d = dst + (((size_t)t * NC + rank) * NE + e) * ((size_t)BW_NF * NT * 8 + (size_t)NT * 4);
#pragma unroll
  for (int f = 0; f < BW_NF; ++f)
    reinterpret_cast<double*>(d)[f * NT + r] = reinterpret_cast<const double*>(s)[(size_t)f * LDA + row];
  reinterpret_cast<int*>(d + (size_t)BW_NF * NT * 8)[r] = reinterpret_cast<const int*>(s + (size_t)BW_NF * LDA * 8)[row];
}
static __global__ void k_tape_rowblocks_fw(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int ncols,
                                    int NE, int LDA, int NT) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncols * LDA) return;
  const int col = (int)(i / LDA), row = (int)(i - (size_t)col * LDA), t = col / NE, e = col - t * NE;
  const int NC = LDA / NT, rank = row / NT, r = row - rank * NT;
  const unsigned char* s = src + (size_t)col * ((size_t)FW_NF * LDA * 8 + (size_t)(LDA + 4) * 4);
  unsigned char* d = dst + (((size_t)t * NC + rank) * NE + e) * ((size_t)FW_NF * NT * 8 + (size_t)(NT + 4) * 4);
#pragma unroll
  for (int f = 0; f < FW_NF; ++f)
    reinterpret_cast<double*>(d)[f * NT + r] = reinterpret_cast<const double*>(s)[(size_t)f * LDA + row];
  const int* st = reinterpret_cast<const int*>(s + (size_t)FW_NF * LDA * 8);
  int* dt = reinterpret_cast<int*>(d + (size_t)FW_NF * NT * 8);
  dt[r] = st[row];
  if (r < 4) dt[NT + r] = st[row + NT];   // the block's range starts overlap the next block's by four entries
}

}  // namespace hank
