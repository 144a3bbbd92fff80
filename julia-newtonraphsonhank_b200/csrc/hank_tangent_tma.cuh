// hank_tangent_tma.cuh — tangent-lane sweeps with the primal tape (and, in the forward sweep, the
// policy-tangent stream itself) staged through a shared-memory ring by the TMA bulk-copy engine.
//
// Why: the sweeps are chains of dependent periods.  With plain loads every column of every period
// paid a DRAM/L2 round trip (measured 1000-2000 cycles per column), which capped the first
// kernels at 27 % / 12 % of the HBM roofline (profiles/r01_notes.md).  Here ONE `cp.async.bulk`
// copy per column (two in the forward sweep; SASS UBLKCP — the tape is laid out as one contiguous
// chunk per column for exactly this purpose) is issued several columns ahead, by lane 0 of a
// different warp each column so the issue cost is spread.  Completion is tracked by one mbarrier
// per ring slot (`mbarrier.arrive.expect_tx` / `complete_tx`); the per-column __syncthreads() the
// algorithm needs anyway doubles as the "slot is free" signal, so no second barrier set and no
// extra producer warp (which would cost 24 registers per thread at 512+32 threads) is needed.
// Compute threads never issue a global load inside the period loop.
//
// Bookkeeping is kept off the per-point path: policy tangents use a lane stride Kp padded to a
// multiple of L (no per-lane store predicates), rows beyond n_a run on the zero padding of the
// tape (no per-row branches), and all chunk / slot addresses advance incrementally.
#pragma once
#include "hank_tangent.cuh"
#ifndef HANK_GATHER_U
#define HANK_GATHER_U 2
#endif

namespace hank {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost copy must trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ======================================================================================
// Backward tangent sweep, TMA-staged tape.
// smem: ring[S][bw chunk] | kb[2][L][LDA] | drs[L][P] | dws[L][P] | rhos[P] | full[S]
// Slot c%S is read in both phases of column c, so it is free after the barrier of column c+1:
// chunk c+S-1 is issued there (prefetch distance S-1 columns).
// dpol: [P][NE][Kp][LDA], Kp a multiple of L (lanes >= K carry zero seeds).
// ======================================================================================
template <int NE, int R, int NT, int L, bool SKIP>
__global__ void __launch_bounds__(NT, 1)
k_backward_tangent_tma(const Consts<NE> M, const Tape tp, int K, int Kp, int S, const int* __restrict__ thi,
                       const double* __restrict__ dr, const double* __restrict__ dw, const double* __restrict__ dvalT,
                       double* __restrict__ dpol, double* __restrict__ dvalue_first) {
  constexpr int LDA = NT * R, NW = NT / 32;
  constexpr size_t GP = (size_t)NE * LDA;
  constexpr int CH = (int)bw_chunk_bytes<LDA>();
  constexpr int SLOT_D = CH / 8;
  extern __shared__ __align__(128) unsigned char smem_tma[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lane0 = blockIdx.x * L;
  // SKIP: seeds of this CTA's lanes are zero from period thi[group] on, so V̇ and ṗ are exactly zero there: the
  // sweep of this CTA is the P-period sweep of a shorter horizon (only the seed rows keep the full stride).
  // A template parameter, not a run-time test: the generic kernel keeps the code it was tuned with.
  const int n_a = M.n_a, Pfull = M.P, P = SKIP ? min(M.P, thi[lane0 / kThiGroup]) : M.P;
  double* ring = reinterpret_cast<double*>(smem_tma);
  double* kbuf = ring + (size_t)S * SLOT_D;
  double* drs = kbuf + 2 * L * LDA;
  double* dws = drs + (size_t)L * P;
  double* rhos = dws + (size_t)L * P;
  uint64_t* full = reinterpret_cast<uint64_t*>(rhos + ((P + 1) & ~1));
  const int nchunks = P * NE;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = tid; i < L * P; i += NT) {
    const int l = i / P, t = i - l * P;
    const bool on = lane0 + l < K;
    drs[i] = on ? dr[(size_t)(lane0 + l) * Pfull + t] : 0.0;
    dws[i] = on ? dw[(size_t)(lane0 + l) * Pfull + t] : 0.0;
  }
  for (int t = tid; t < P; t += NT) rhos[t] = tp.rho[t];
  __syncthreads();
  // chunk c lives at tape.bw + ((P-1-c/NE)*NE + c%NE)*CH
  if (tid == 0)
    for (int c = 0; c < S - 1 && c < nchunks; ++c) {
      mbar_expect_tx(&full[c], CH);
      bulk_g2s(ring + (size_t)c * SLOT_D, tp.bw + ((size_t)(P - 1 - c / NE) * NE + c % NE) * CH, CH, &full[c]);
    }
  // issue cursor: next chunk to request, its source and its slot (only dereferenced while ci < nchunks)
  int ci = S - 1, ei = (S - 1) % NE, si = S - 1;
  const unsigned char* isrc = tp.bw + ((ptrdiff_t)(P - 1 - (S - 1) / NE) * NE + ei) * CH;

  double Vd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e)
        Vd[l][j][e] = (dvalT && tid + j * NT < n_a && lane0 + l < K)
            ? dvalT[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] : 0.0;

  const size_t strideKL = (size_t)Kp * LDA;                        // doubles between columns of dpol
  double* dp_t = dpol + ((ptrdiff_t)(P - 1) * NE * Kp + lane0) * LDA + tid;
  const double* sl = ring + tid;                                   // current slot, this thread's row
  int slot = 0, par = 0, pb = 0, iw = 0;
  for (int t = P - 1; t >= 0; --t) {
    const double rho = rhos[t];
    double drl[L], dwl[L];
#pragma unroll
    for (int l = 0; l < L; ++l) { drl[l] = drs[l * P + t]; dwl[l] = dws[l * P + t]; }
    // ---- phase 0 (registers only): ĖV in place of V̇⁺
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        double ev[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          double s = 0.0;
#pragma unroll
          for (int e2 = 0; e2 < NE; ++e2) s = fma(M.Pi[e][e2], Vd[l][j][e2], s);
          ev[e] = s;
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) Vd[l][j][e] = ev[e];
      }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int* sli = reinterpret_cast<const int*>(sl - tid + BW_NF * LDA) + tid;
      double* kb = kbuf + (size_t)pb * L * LDA;
      mbar_wait(&full[slot], par);
      const double cw = -(rho * M.z[e]);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double a1 = sl[BW_A1 * LDA + j * NT], kr = sl[BW_KR * LDA + j * NT];
#pragma unroll
        for (int l = 0; l < L; ++l)
          kb[l * LDA + j * NT + tid] = fma(a1, Vd[l][j][e], fma(kr, drl[l], cw * dwl[l]));
      }
      __syncthreads();
      // every thread has finished reading the previous column's slot: refill it S-1 columns ahead
      if (ci < nchunks) {
        if (warp == iw && lane == 0) {
          mbar_expect_tx(&full[si], CH);
          bulk_g2s(ring + (size_t)si * SLOT_D, isrc, CH, &full[si]);
        }
        ++ci;
        if (++ei == NE) { ei = 0; isrc -= (size_t)(2 * NE - 1) * CH; } else isrc += CH;
        if (++si == S) si = 0;
        if (++iw == NW) iw = 0;
      }
      const double ze = M.z[e];
      double* dpc = dp_t + (size_t)e * strideKL;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double cA = sl[BW_CA * LDA + j * NT], cB = sl[BW_CB * LDA + j * NT];
        const double E1 = sl[BW_E1 * LDA + j * NT], vf = sl[BW_VF * LDA + j * NT];
        const double* kk = kb + sli[j * NT];
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const double pd = fma(cA, kk[l * LDA], cB * kk[l * LDA + 1]);
          __stcs(dpc + (size_t)l * LDA + j * NT, pd);
          Vd[l][j][e] = fma(vf, fma(ze, dwl[l], -pd), E1 * drl[l]);
        }
      }
      pb ^= 1;
      sl += SLOT_D;
      if (++slot == S) { slot = 0; par ^= 1; sl -= (size_t)S * SLOT_D; }
    }
    dp_t -= (size_t)NE * strideKL;
  }
  if (dvalue_first) {
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (lane0 + l < K)
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (tid + j * NT < n_a)
#pragma unroll
            for (int e = 0; e < NE; ++e)
              dvalue_first[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] = Vd[l][j][e];
  }
}

// ======================================================================================
// Backward tangent sweep with ONE RING SLOT PER INCOME STATE (S = NE).
//
// The kernel above spends more issue slots on ring bookkeeping than on arithmetic: of the 173 instructions a warp
// issues per column (L = 4), 61 are FP64 and 24 memory, the other ~87 advance the slot / chunk / issuer cursors and
// their wrap-arounds (IMAD, IADD3, VIADD, SEL, ISETP, PLOP3, R2UR; profiles/r02_notes.md).  With as many slots as
// columns the slot of column e IS slot e: every shared-memory address of the column is a compile-time offset, the
// barrier phase is the period's parity, the exchange buffer alternates with (e + period) & 1, and the refill of the
// slot freed at the barrier of column e is issued by warp e from a pointer that moves once per period: 141
// instructions per column.  The seeds of the period (ṙ, ẇ of the L lanes) are broadcast through a 2-deep shared
// buffer that the first threads fill one period ahead, which frees the [L][P] staging arrays the seven slots need the
// room of.
// smem: ring[NE][bw chunk] | kb[2][L][LDA] | drw[2][2L] | full[NE].   Fits for NE * 52 * LDA + 16 L LDA <= ~225 KB
// (500 x 7 with up to 4 lanes); other shapes keep the kernel above.  No seeded V̇ (single-step callers).
// ======================================================================================
// Pipelined linearisation: chunk (t, e) of the backward tape may be fetched once the primal sweep running next to this
// kernel has published P - t completed periods for income state e (k_backward_primal_ds).  `seen` caches the last
// count read: the primal runs ~4x faster per period than a wave of this sweep, so a handful of polls cover a sweep.
__device__ __forceinline__ void wait_tape_flag(const int* flag, int need, int& seen) {
  if (seen >= need) return;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
    if (clock64() - t0 > 4000000000LL) __trap();
  } while (seen < need);
  asm volatile("fence.proxy.async;" ::: "memory");   // the bulk copy (async proxy) reads what the generic proxy wrote
}

template <int NE, int LDA, int L>
constexpr size_t bt_ring_ne_smem() { return (size_t)NE * bw_chunk_bytes<LDA>() + (size_t)2 * L * LDA * 8 + (size_t)4 * L * 8 + (size_t)NE * 8 + 128; }

template <int NE, int R, int NT, int L, bool SKIP>
__global__ void __launch_bounds__(NT, 1)
k_backward_tangent_ring_ne(const Consts<NE> M, const Tape tp, int K, int Kp, const int* __restrict__ thi,
                           const double* __restrict__ dr, const double* __restrict__ dw, double* __restrict__ dpol,
                           const int* __restrict__ bpflag) {
  constexpr int LDA = NT * R, NW = NT / 32;
  constexpr int CH = (int)bw_chunk_bytes<LDA>();
  constexpr int SLOT_D = CH / 8;
  extern __shared__ __align__(128) unsigned char smem_tma[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lane0 = blockIdx.x * L;
  const int Pfull = M.P, P = SKIP ? min(M.P, thi[lane0 / kThiGroup]) : M.P;
  double* ring = reinterpret_cast<double*>(smem_tma);
  double* kbuf = ring + (size_t)NE * SLOT_D;
  double* drw = kbuf + 2 * L * LDA;                      // [2][2L]: ṙ of the L lanes, then ẇ
  uint64_t* full = reinterpret_cast<uint64_t*>(drw + 4 * L);
  if (P <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < NE; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // seeds of period t for lane slot i < 2L (i < L: ṙ, else ẇ); lanes beyond K carry zeros
  auto seed = [&](int i, int t) -> double {
    const int l = i < L ? i : i - L;
    if (lane0 + l >= K) return 0.0;
    return __ldg((i < L ? dr : dw) + (size_t)(lane0 + l) * Pfull + t);
  };
  if (tid < 2 * L) drw[tid] = seed(tid, P - 1);
  __syncthreads();
  const unsigned char* bw_t = tp.bw + (size_t)(P - 1) * NE * CH;   // chunks of the current period
  // (pipelined linearisation) the column this thread refills — warp w, lane 0 issues column w-1, warp 0 column NE-1 —
  // and the primal's progress on it as last seen
  // (with at least NE warps that is one column per issuing thread; otherwise a thread re-reads the counter every time)
  int seen = bpflag ? 0 : 0x7fffffff;
  if (tid == 0)
    for (int e = 0; e < NE; ++e) {
      if (bpflag) { int s0 = 0; wait_tape_flag(bpflag + e, Pfull - (P - 1), s0); }
      mbar_expect_tx(&full[e], CH);
      bulk_g2s(ring + (size_t)e * SLOT_D, bw_t + (size_t)e * CH, CH, &full[e]);
    }

  double Vd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) Vd[l][j][e] = 0.0;

  const size_t strideKL = (size_t)Kp * LDA;
  double* dp_t = dpol + ((ptrdiff_t)(P - 1) * NE * Kp + lane0) * LDA + tid;
  double* kb0 = kbuf;                 // exchange buffer of the even columns of this period
  double* kb1 = kbuf + L * LDA;       //                     odd
  const double* slt = ring + tid;
  uint32_t par = 0;
  double rho = __ldg(tp.rho + P - 1);
  for (int t = P - 1; t >= 0; --t) {
    const double rho_t = rho;
    if (t > 0) rho = __ldg(tp.rho + t - 1);
    const double* dq = drw + ((P - 1 - t) & 1) * 2 * L;
    double drl[L], dwl[L];
#pragma unroll
    for (int l = 0; l < L; ++l) { drl[l] = dq[l]; dwl[l] = dq[L + l]; }
    // next period's seeds: the other half of drw was last read a period ago (>= NE barriers back)
    if (tid < 2 * L && t > 0) drw[(((P - 1 - t) & 1) ^ 1) * 2 * L + tid] = seed(tid, t - 1);
    // ---- phase 0 (registers only): ĖV in place of V̇⁺
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        double ev[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          double s = 0.0;
#pragma unroll
          for (int e2 = 0; e2 < NE; ++e2) s = fma(M.Pi[e][e2], Vd[l][j][e2], s);
          ev[e] = s;
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) Vd[l][j][e] = ev[e];
      }
    double* dpc = dp_t;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const double* sl = slt + e * SLOT_D;
      const int* sli = reinterpret_cast<const int*>(ring + e * SLOT_D + BW_NF * LDA) + tid;
      double* kb = (e & 1) ? kb1 : kb0;
      mbar_wait(&full[e], par);
      const double cw = -(rho_t * M.z[e]);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double a1 = sl[BW_A1 * LDA + j * NT], kr = sl[BW_KR * LDA + j * NT];
#pragma unroll
        for (int l = 0; l < L; ++l)
          kb[l * LDA + j * NT + tid] = fma(a1, Vd[l][j][e], fma(kr, drl[l], cw * dwl[l]));
      }
      __syncthreads();
      // every thread is through both phases of the previous column: its slot takes the chunk that column needs next
      // (column e-1 of period t-1; for e = 0 the last column of THIS period, six columns ahead)
      if (warp == e % NW && lane == 0) {
        if (e > 0) {
          if (t > 0) {
            if (NW < NE && bpflag) seen = 0;
            wait_tape_flag(bpflag + (e - 1), Pfull - (t - 1), seen);
            mbar_expect_tx(&full[e - 1], CH); bulk_g2s(ring + (size_t)(e - 1) * SLOT_D, bw_t - (size_t)(NE - e + 1) * CH, CH, &full[e - 1]);
          }
        } else if (t < P - 1) {
          if (NW < NE && bpflag) seen = 0;
          wait_tape_flag(bpflag + (NE - 1), Pfull - t, seen);
          mbar_expect_tx(&full[NE - 1], CH); bulk_g2s(ring + (size_t)(NE - 1) * SLOT_D, bw_t + (size_t)(NE - 1) * CH, CH, &full[NE - 1]);
        }
      }
      const double ze = M.z[e];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const double cA = sl[BW_CA * LDA + j * NT], cB = sl[BW_CB * LDA + j * NT];
        const double E1 = sl[BW_E1 * LDA + j * NT], vf = sl[BW_VF * LDA + j * NT];
        const double* kk = kb + sli[j * NT];
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const double pd = fma(cA, kk[l * LDA], cB * kk[l * LDA + 1]);
          __stcs(dpc + (size_t)l * LDA + j * NT, pd);
          Vd[l][j][e] = fma(vf, fma(ze, dwl[l], -pd), E1 * drl[l]);
        }
      }
      dpc += strideKL;
    }
    if (NE & 1) { double* sw = kb0; kb0 = kb1; kb1 = sw; }   // an odd column count flips the buffer parity every period
    par ^= 1;
    bw_t -= (size_t)NE * CH;
    dp_t -= (size_t)NE * strideKL;
  }
}

// ======================================================================================
// Forward tangent sweep, TMA-staged tape and ṗ stream.  Slot = forward chunk | ṗ of L lanes.
// smem: ring[S][fw chunk + L*LDA doubles] | Xb[2][L][LDA] | Yb[2][L][LDA] | full[S]
// A slot is only read before the column's barrier, so it is free right after it: chunk c+S is
// issued after the barrier of column c (prefetch distance S columns).
// ======================================================================================
template <int NE, int R, int NT, int L, bool SKIP>
__global__ void __launch_bounds__(NT, 1)
k_forward_tangent_tma(const Consts<NE> M, const Tape tp, int K, int Kp, int S, const int* __restrict__ thi,
                      const double* __restrict__ zeros, const double* __restrict__ dpol,
                      const double* __restrict__ dD0, double* __restrict__ dkdpart,
                      double* __restrict__ dD_last) {
  constexpr int LDA = NT * R, U = HANK_GATHER_U, NW = NT / 32;
  constexpr size_t GP = (size_t)NE * LDA;
  constexpr int CH = (int)fw_chunk_bytes<LDA>();
  constexpr int PD_OFF = CH / 8;                 // ṗ lanes follow the chunk (CH is a multiple of 16)
  constexpr int SLOT_D = PD_OFF + L * LDA;
  constexpr uint32_t PDB = (uint32_t)L * LDA * 8;
  extern __shared__ __align__(128) unsigned char smem_tma[];
  const int n_a = M.n_a, P = M.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lane0 = blockIdx.x * L;
  double* ring = reinterpret_cast<double*>(smem_tma);
  double* Xb = ring + (size_t)S * SLOT_D;
  double* Yb = Xb + 2 * L * LDA;
  uint64_t* full = reinterpret_cast<uint64_t*>(Yb + 2 * L * LDA);
  const int nchunks = P * NE;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // chunk c = t*NE + e: the tape chunk at tp.fw + c*CH and ṗ at dpol + (c*Kp + lane0)*LDA
  const size_t strideKL = (size_t)Kp * LDA;
  const double* psrc0 = dpol + (size_t)lane0 * LDA;
  // SKIP: beyond the seed horizon of this CTA's lanes ṗ is zero and was never written — stage the zero page
  const int ce = SKIP ? min(P, thi[lane0 / kThiGroup]) * NE : 0;
  auto issue = [&](int c, int s) {
    double* dst = ring + (size_t)s * SLOT_D;
    mbar_expect_tx(&full[s], (uint32_t)CH + PDB);
    bulk_g2s(dst, tp.fw + (size_t)c * CH, CH, &full[s]);
    if constexpr (SKIP) bulk_g2s(dst + PD_OFF, c < ce ? psrc0 + (size_t)c * strideKL : zeros, PDB, &full[s]);
    else bulk_g2s(dst + PD_OFF, psrc0 + (size_t)c * strideKL, PDB, &full[s]);
  };
  if (tid == 0)
    for (int c = 0; c < S && c < nchunks; ++c) issue(c, c);
  int ci = S, si = 0;

  double Dd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e)
        Dd[l][j][e] = (dD0 && tid + j * NT < n_a && lane0 + l < K) ? dD0[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] : 0.0;

  const double* sl = ring + tid;
  int slot = 0, par = 0, pb = 0, iw = 0;
  for (int t = 0; t < P; ++t) {
    double kacc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) kacc[l] = 0.0;
    double pv[R][NE];  // p_t of the row, picked up column by column for <p_t, Ḋ_t>
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int* sst = reinterpret_cast<const int*>(sl - tid + FW_NF * LDA) + tid + 1;
      double* xb = Xb + (size_t)pb * L * LDA;
      double* yb = Yb + (size_t)pb * L * LDA;
      mbar_wait(&full[slot], par);
      int s0[R], s1[R], s2[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        // all shared-memory loads of the row first: the stores below may alias them as far as the
        // compiler can tell, so loads written after a store would wait for it
        double pd[L];
#pragma unroll
        for (int l = 0; l < L; ++l) pd[l] = sl[PD_OFF + l * LDA + j * NT];
        const double om = sl[FW_OM * LDA + j * NT], dco = sl[FW_DCO * LDA + j * NT], Dn = sl[FW_D * LDA + j * NT];
        pv[j][e] = sl[FW_P * LDA + j * NT];
        s0[j] = sst[j * NT]; s1[j] = sst[j * NT + 1]; s2[j] = sst[j * NT + 2];
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const double xd = fma(om, Dd[l][j][e], dco * pd[l]);
          xb[l * LDA + j * NT + tid] = xd;
          yb[l * LDA + j * NT + tid] = Dd[l][j][e] - xd;
          kacc[l] = fma(pd[l], Dn, kacc[l]);
        }
      }
      __syncthreads();
      if (ci < nchunks) {   // this column's slot is free again: refill it S columns ahead
        if (warp == iw && lane == 0) issue(ci, si);
        ++ci;
        if (++si == S) si = 0;
        if (++iw == NW) iw = 0;
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        double acc[L];
        gather_row<L, LDA, U>(xb, yb, s0[j], s1[j], s2[j], lane, acc);
#pragma unroll
        for (int l = 0; l < L; ++l) Dd[l][j][e] = acc[l];
      }
      pb ^= 1;
      sl += SLOT_D;
      if (++slot == S) { slot = 0; par ^= 1; sl -= (size_t)S * SLOT_D; }
    }
    // ---- Markov mix (in place) and second aggregation term <p_t, Ḋ_t>
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int l = 0; l < L; ++l) {
        double d[NE];
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) {
          double s = 0.0;
#pragma unroll
          for (int e = 0; e < NE; ++e) s = fma(M.Pi[e][e2], Dd[l][j][e], s);
          d[e2] = s;
        }
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) { Dd[l][j][e2] = d[e2]; kacc[l] = fma(pv[j][e2], d[e2], kacc[l]); }
      }
    }
    if constexpr (L == 1 || L == 2 || L == 4) {
      const double s = warp_sum_multi<L>(kacc, lane);
      const int l = lane / (32 / L);
      if ((lane & (32 / L - 1)) == 0 && lane0 + l < K) dkdpart[((size_t)(lane0 + l) * P + t) * NW + warp] = s;
    } else {
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const double s = warp_sum(kacc[l]);
        if (lane == 0 && lane0 + l < K) dkdpart[((size_t)(lane0 + l) * P + t) * NW + warp] = s;
      }
    }
  }
  if (dD_last) {
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (lane0 + l < K)
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (tid + j * NT < n_a)
#pragma unroll
            for (int e = 0; e < NE; ++e)
              dD_last[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] = Dd[l][j][e];
  }
}

}  // namespace hank
