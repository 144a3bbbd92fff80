// hank_launch.cuh — launch-configuration dispatch for the sweep kernels, templated on n_e.
// Included only by hank_ne*.cu, each of which instantiates Sweeps<NE> for one n_e so the
// instantiations compile in parallel.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "hank_ctx.h"
#include "hank_tangent.cuh"
#include "hank_tangent_tma.cuh"
#include "hank_primal_cluster.cuh"
#include "hank_primal_dsmem.cuh"

namespace hank {

template <int NE>
static Consts<NE> make_consts(const hank_ctx* c, int P) {
  Consts<NE> M;
  for (int e = 0; e < NE; ++e) {
    M.z[e] = c->h_z[e];
    for (int e2 = 0; e2 < NE; ++e2) M.Pi[e][e2] = c->h_Pi[e * NE + e2];
  }
  M.beta = c->beta; M.gamma = c->gamma; M.bc = c->bc; M.yexp = -1.0 / c->gamma;
  M.n_a = c->n_a; M.P = P;
  M.gamma_int = (std::fabs(c->gamma) < 0x1.8p62 && (double)(long long)c->gamma == c->gamma) ? 1 : 0;
  return M;
}

template <typename KernelT>
static int set_smem(hank_ctx* c, KernelT k, size_t bytes) {
  if ((int)bytes > c->smem_max)
    return set_error(c, 1, "shared memory needed (" + std::to_string(bytes) + " B) exceeds the device limit");
  return cuda_check(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                    "cudaFuncSetAttribute");
}

#define HANK_LAUNCH(kind, kern, grid, block, smem, ...)                             \
  do {                                                                              \
    auto kfn_ = kern;                                                               \
    int rc_ = set_smem(c, kfn_, smem);                                              \
    if (rc_) return rc_;                                                            \
    cudaEvent_t ev_ = prof_begin(c);                                                \
    kfn_<<<grid, block, smem, c->stream>>>(__VA_ARGS__);                            \
    prof_end(c, kind, ev_);                                                         \
    c->launches++;                                                                  \
    return cuda_check(c, cudaGetLastError(), #kern);                                \
  } while (0)

// Launch `kern` as ONE cluster of NE CTAs (one income state per CTA). Returns -1 if the cluster
// cannot be scheduled (caller falls back to the single-CTA kernel).
template <int NE, typename KernelT, typename... Args>
static int launch_cluster(hank_ctx* c, int kind, KernelT kern, int block, size_t smem, const char* name, Args... args) {
  if (NE > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int rc = set_smem(c, kern, smem);
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NE); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NE; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl < 1) { cudaGetLastError(); return -1; }
  if (c->next_launch_ev) {   // fires once every CTA of this launch has begun execution (hank_ctx.h: ev_bps)
    at[1].id = cudaLaunchAttributeLaunchCompletionEvent;
    at[1].val.launchCompletionEvent.event = c->next_launch_ev;
    at[1].val.launchCompletionEvent.flags = 0;
    cfg.numAttrs = 2;
    c->next_launch_ev = nullptr;
  }
  cudaEvent_t ev = prof_begin(c);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  prof_end(c, kind, ev);
  c->launches++;
  return cuda_check(c, e, name);
}

// Launch `kern` on `grid` CTAs in clusters of `cluster` CTAs.  Returns -1 if such a cluster cannot be scheduled.
template <typename KernelT, typename... Args>
static int launch_cluster_grid(hank_ctx* c, int kind, KernelT kern, int grid, int cluster, int block, size_t smem,
                               const char* name, Args... args) {
  if (cluster > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int rc = set_smem(c, kern, smem);
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl < 1) { cudaGetLastError(); return -1; }
  cudaEvent_t ev = prof_begin(c);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  prof_end(c, kind, ev);
  c->launches++;
  return cuda_check(c, e, name);
}

// ---- backward primal -------------------------------------------------------------------
template <int NE, int R, int NT>
static int bp_launch(hank_ctx* c, int P, const double* valueT, const double* r, const double* w) {
  const Consts<NE> M = make_consts<NE>(c, P);
  c->tape_rs_bw_nt = 0;
  c->bp_flags = false;
  // exchange through distributed shared memory (hank_primal_dsmem.cuh); with two or more rows per thread the
  // per-row remote stores cost more than the fence they replace (1000x7: 4.16 vs 3.97 us per period)
  if constexpr (R == 1) if (!c->no_cluster && !c->no_dsmem && NE > 1 && bp_ds_smem<NE, NT * R>() <= (size_t)c->smem_max) {
    const size_t smem_d = bp_ds_smem<NE, NT * R>();
    // a pipelined linearisation (hank_ctx.h): progress counters per income state + "all CTAs resident" event
    int* flags = nullptr;
    if (c->bp_pipe_req && c->d_bpflag) {
      flags = c->d_bpflag;
      if (cudaMemsetAsync(flags, 0, 16 * sizeof(int), c->stream) != cudaSuccess) { cudaGetLastError(); flags = nullptr; }
    }
    c->next_launch_ev = flags ? c->ev_bps : nullptr;
    int rc = c->gamma == 2.0
        ? launch_cluster<NE>(c, KIND_BP, k_backward_primal_ds<NE, R, NT, true>, NT + 32, smem_d, "k_backward_primal_ds", M, c->tape,
                             (const double*)c->d_grid, valueT, r, w, c->d_status, flags)
        : launch_cluster<NE>(c, KIND_BP, k_backward_primal_ds<NE, R, NT, false>, NT + 32, smem_d, "k_backward_primal_ds", M, c->tape,
                             (const double*)c->d_grid, valueT, r, w, c->d_status, flags);
    c->next_launch_ev = nullptr;
    if (rc >= 0) { c->bp_flags = flags != nullptr && rc == 0; return rc; }
  }
  if (!c->no_cluster && NE > 1) {
    const size_t smem_c = (size_t)2 * NT * R * sizeof(double);
    int rc = c->gamma == 2.0
        ? launch_cluster<NE>(c, KIND_BP, k_backward_primal_cl<NE, R, NT, true>, NT, smem_c, "k_backward_primal_cl", M, c->tape,
                             (const double*)c->d_grid, valueT, r, w, c->d_xch, c->d_status)
        : launch_cluster<NE>(c, KIND_BP, k_backward_primal_cl<NE, R, NT, false>, NT, smem_c, "k_backward_primal_cl", M, c->tape,
                             (const double*)c->d_grid, valueT, r, w, c->d_xch, c->d_status);
    if (rc >= 0) return rc;
  }
  const size_t smem = ((size_t)NE * NT * R + NT * R) * sizeof(double);
  if (c->gamma == 2.0)
    HANK_LAUNCH(KIND_BP, (k_backward_primal<NE, R, NT, true>), 1, NT, smem, M, c->tape, c->d_grid, valueT, r, w, c->d_status);
  else
    HANK_LAUNCH(KIND_BP, (k_backward_primal<NE, R, NT, false>), 1, NT, smem, M, c->tape, c->d_grid, valueT, r, w, c->d_status);
}
template <int NE>
int Sweeps<NE>::backward_primal(hank_ctx* c, int P, const double* valueT, const double* r, const double* w) {
  Shape s;
  if (!pick_shape(c->n_a, &s)) return set_error(c, 1, "n_a > 2048 is not supported");
  if (s.NT == 256) return bp_launch<NE, 1, 256>(c, P, valueT, r, w);
  if (s.R == 1) return bp_launch<NE, 1, 512>(c, P, valueT, r, w);
  if (s.R == 2) return bp_launch<NE, 2, 512>(c, P, valueT, r, w);
  return bp_launch<NE, 4, 512>(c, P, valueT, r, w);
}

// ---- both primal sweeps, one launch (pipelined linearisation) ------------------------------
template <int NE, int NT>
static int pb_launch(hank_ctx* c, int P, const double* valueT, const double* r, const double* w, const double* D0) {
  if (c->no_cluster || c->no_dsmem || NE <= 1 || !c->d_bpflag || fp_ds_smem<NE, NT>() > (size_t)c->smem_max) return -1;
  const Consts<NE> M = make_consts<NE>(c, P);
  const size_t smem = fp_ds_smem<NE, NT>() > bp_ds_smem<NE, NT>() ? fp_ds_smem<NE, NT>() : bp_ds_smem<NE, NT>();
  if (cudaMemsetAsync(c->d_bpflag, 0, 16 * sizeof(int), c->stream) != cudaSuccess) { cudaGetLastError(); return -1; }
  c->next_launch_ev = c->ev_bps;
  int rc = c->gamma == 2.0
      ? launch_cluster<NE>(c, KIND_BP, k_primal_ds_both<NE, 1, NT, true>, NT + 32, smem, "k_primal_ds_both", M, c->tape,
                           (const double*)c->d_grid, valueT, r, w, c->d_status, c->d_bpflag, D0, c->d_kdpart)
      : launch_cluster<NE>(c, KIND_BP, k_primal_ds_both<NE, 1, NT, false>, NT + 32, smem, "k_primal_ds_both", M, c->tape,
                           (const double*)c->d_grid, valueT, r, w, c->d_status, c->d_bpflag, D0, c->d_kdpart);
  c->next_launch_ev = nullptr;
  if (rc != 0) return rc;
  c->tape_rs_bw_nt = 0; c->tape_rs_fw_nt = 0;
  c->bp_flags = true; c->fp_cluster = true;
  return 0;
}
template <int NE>
int Sweeps<NE>::primal_both(hank_ctx* c, int P, const double* valueT, const double* r, const double* w, const double* D0) {
  Shape s;
  if (!pick_shape(c->n_a, &s) || s.R != 1) return -1;
  if (s.NT == 256) return pb_launch<NE, 256>(c, P, valueT, r, w, D0);
  return pb_launch<NE, 512>(c, P, valueT, r, w, D0);
}

// ---- forward primal --------------------------------------------------------------------
template <int NE, int R, int NT, int CS>
static int fp_launch(hank_ctx* c, int P, const double* D0, const double* pol, double* KD) {
  const Consts<NE> M = make_consts<NE>(c, P);
  c->tape_rs_fw_nt = 0;
  constexpr int LDA = NT * R;
  if (!c->no_cluster && !c->no_dsmem && NE > 1 && CS == NE && fp_ds_smem<NE, LDA>() <= (size_t)c->smem_max) {
    int rc = launch_cluster<NE>(c, KIND_FP, k_forward_primal_ds<NE, R, NT>, NT, fp_ds_smem<NE, LDA>(), "k_forward_primal_ds", M,
                                c->tape, (const double*)c->d_grid, D0, pol, c->d_kdpart, c->d_status);
    if (rc >= 0) { c->fp_cluster = true; return rc; }
  }
  if (!c->no_cluster && NE > 1 && CS == NE) {   // (the CS < NE instantiation only exists as the single-CTA fallback)
    const size_t smem_c = (size_t)3 * LDA * sizeof(double) + ((size_t)LDA + LDA + 4) * sizeof(int);
    int rc = launch_cluster<NE>(c, KIND_FP, k_forward_primal_cl<NE, R, NT>, NT, smem_c, "k_forward_primal_cl", M, c->tape,
                                (const double*)c->d_grid, D0, pol, c->d_xch, c->d_kdpart, c->d_status);
    if (rc >= 0) { c->fp_cluster = true; return rc; }
  }
  c->fp_cluster = false;
  const size_t smem = ((size_t)2 * CS * LDA + LDA) * sizeof(double) + ((size_t)CS * LDA + (size_t)CS * (LDA + 4)) * sizeof(int);
  if (smem > (size_t)c->smem_max) return -2;  // caller retries with fewer columns staged at once
  HANK_LAUNCH(KIND_FP, (k_forward_primal<NE, R, NT, CS>), 1, NT, smem, M, c->tape, c->d_grid, D0, pol, c->d_kdpart, KD, c->d_status);
}
template <int NE>
static constexpr int small_cs() { return NE < 4 ? NE : 4; }
template <int NE>
int Sweeps<NE>::forward_primal(hank_ctx* c, int P, const double* D0, const double* pol, double* KD) {
  Shape s;
  if (!pick_shape(c->n_a, &s)) return set_error(c, 1, "n_a > 2048 is not supported");
  const size_t lda = c->lda;
  const size_t full = ((size_t)2 * NE * lda + lda) * 8 + ((size_t)NE * lda + (size_t)NE * (lda + 4)) * 4;
  const bool fits = full <= (size_t)c->smem_max || !c->no_cluster;
  constexpr int C4 = small_cs<NE>();
#define FP(R_, NT_) do { if (fits) { int rc_ = fp_launch<NE, R_, NT_, NE>(c, P, D0, pol, KD); if (rc_ != -2) return rc_; } \
                         return fp_launch<NE, R_, NT_, C4>(c, P, D0, pol, KD); } while (0)
  if (s.NT == 256) FP(1, 256);
  if (s.R == 1) FP(1, 512);
  if (s.R == 2) FP(2, 512);
  FP(4, 512);
#undef FP
}

// ---- tangent kernels: (NT, R, L) by leading dimension and lane count ---------------------
// The register file bounds the lanes a CTA can carry (state = L*R*NE doubles per thread):
// 512 threads x R rows hold L*R <= 4; 256 threads x 2R rows hold 6 lanes at LDA 512 (3 at 1024),
// which cuts the tape bytes staged per lane by a third.  Deeper lanes are only used once every SM
// has a CTA.
struct TangentCfg { int NT, R, L; int NC = 0, GC = 0, LA = 0; };   // NC > 0: rows split over a cluster (hank_tangent_rowsplit.cuh)
// waves(K, L) * relative cost of one wave with L lanes per CTA (measured at 500x7: a 6-lane wave
// costs 1.6x a 4-lane wave, a 4-lane wave 1.9x a 1-lane wave; profiles/r01_notes.md)
static double cfg_cost(int K, int L, int sm, double wave_cost) {
  const int ctas = (K + L - 1) / L;
  return (double)((ctas + sm - 1) / sm) * wave_cost;
}
// Row-split cluster shapes for passes with few lanes: one lane per cluster of 8 CTAs while every lane can have a
// cluster of its own, then wider row blocks (see below) while one wave of clusters still covers the pass.  How many clusters
// can be resident at once (an 8-CTA cluster has to fit inside one GPC) is asked of the driver once per shape.
template <int NE>
static bool rowsplit_cfg(const hank_ctx* c, int K, TangentCfg* out) {
  if (c->no_rowsplit || K < 1) return false;
  const int ne = c->n_e;
  const int nc = c->lda == 256 ? 4 : 8, nt = c->lda / nc;
  // a whole period per exchange when the period's tape rows fit the ring twice over, else column by column
  const bool whole = (size_t)2 * ne * (52 * nt) + (size_t)2 * ne * nt * 8 + 1024 <= (size_t)c->smem_max && c->lda <= 1024;
  // (otherwise four columns per exchange: with the cluster-scope release a hand-shake costs ~0.45 us, so few large groups
  // beat the column-by-column look-ahead pipeline that the relaxed hand-shakes favoured)
  const TangentCfg one = whole ? TangentCfg{nt, 1, 1, nc, ne, 0} : TangentCfg{nt, 1, 1, nc, 4, 0};
  hank_ctx* cm = const_cast<hank_ctx*>(c);
  if (c->rs_cap[0] < 0) cm->rs_cap[0] = Sweeps<NE>::rs_max_clusters(cm, one.NC, one.NT, one.L, one.GC);
  const int max_k = c->rs_max_k > 0 ? c->rs_max_k : c->rs_cap[0];
  if (K <= max_k) { *out = one; return true; }
  if (c->rs_no_multi) return false;
  // more lanes than 8-CTA clusters fit: 500 rows -> one lane per CLUSTER OF 2 (a whole period per exchange);
  // 1000 / 2000 rows -> 2 lanes per cluster of 4, four / two columns per exchange (what fits the ring beside the
  // exchange buffers: a hand-shake costs ~700 cycles whatever it carries); while one wave of clusters covers the pass
  TangentCfg mid; int lm;
  if (c->lda == 512) { mid = TangentCfg{256, 1, 1, 2, ne, 0}; lm = 1; }
  else if (c->lda == 1024) { mid = TangentCfg{256, 1, 2, 4, 4, 0}; lm = 2; }   // 4 columns per exchange
  else if (c->lda == 2048) { mid = TangentCfg{512, 1, 2, 4, 2, 0}; lm = 2; }   // 2 columns per exchange
  else return false;
  if (c->rs_cap[1] < 0) cm->rs_cap[1] = Sweeps<NE>::rs_max_clusters(cm, mid.NC, mid.NT, mid.L, mid.GC);
  // 2000 rows: the one-CTA kernels have no TMA ring at this size (a 106 KB column chunk), so the row-split clusters also
  // carry passes of more than one wave of clusters (clusters are independent: the rest queue behind the resident ones)
  if ((K + lm - 1) / lm <= c->rs_cap[1] || (c->lda == 2048 && c->rs_cap[1] > 0)) { *out = mid; return true; }
  return false;
}
template <int NE>
static TangentCfg tangent_cfg(const hank_ctx* c, int K, bool allow_rowsplit = true, bool forward = false) {
  const int sm = c->sm_count;
  TangentCfg rs;
  if (allow_rowsplit && rowsplit_cfg<NE>(c, K, &rs)) return rs;
  switch (c->lda) {
    case 256: return {256, 1, K >= 4 * sm ? 4 : (K >= 2 * sm ? 2 : 1)};
    case 512: {
      // Seed-horizon passes run their lanes longest first, and a BACKWARD CTA's work shrinks with its horizon: a few
      // CTAs more than one wave are the shortest ones and start behind the first to finish (300 unit-seed lanes of a
      // 2-GPU Jacobian build: one wave of 2-lane CTAs plus two short ones, not half a wave of 4-lane CTAs).  The
      // forward sweep has no short CTAs (every lane runs all periods), so it keeps the strict wave; the two sweeps may
      // use different lanes per CTA because the pass fixes the lane stride of the policy tangents (pass_Kp).
      const int wave = (c->pass_thi && !forward) ? sm + sm / 6 : sm;
      if (K <= wave) return {512, 1, 1};
      if (K <= 2 * wave) return {512, 1, 2};
      const double c4 = cfg_cost(K, 4, sm, 1.0), c6 = cfg_cost(K, 6, sm, 1.6);
      if (c6 < c4 && !c->no_wide) return {256, 2, 6};
      return {512, 1, 4};
    }
    case 1024: {
      if (K <= sm) return {512, 2, 1};
      const double c2 = cfg_cost(K, 2, sm, 1.0), c3 = cfg_cost(K, 3, sm, 1.6);
      // (the 3-lane shape only pays off backward: its forward kernel, 4 rows x 3 lanes fully unrolled at 255 registers,
      // stalls on instruction fetch — 6.6 vs 4.5 ms at K = 444, profiles/r02_notes.md; the lane stride is common)
      if (c3 < c2 && !c->no_wide && !forward) return {256, 4, 3};
      return {512, 2, 2};
    }
    default: return {512, 4, 1};
  }
}
template <int NE>
int Sweeps<NE>::lanes_per_cta(hank_ctx* c, int K) { return tangent_cfg<NE>(c, K).L; }

template <int NE, int R, int NT, int L>
static int bt_launch(hank_ctx* c, int P, int K, const double* dr, const double* dw, const double* dvalT,
                     double* dpol, double* dvf) {
  const Consts<NE> M = make_consts<NE>(c, P);
  const int grid = (K + L - 1) / L;
  const int Kp = c->pass_Kp ? c->pass_Kp : (K + kThiGroup - 1) / kThiGroup * kThiGroup;   // lane stride of the policy tangents: every lanes-per-CTA count divides 12, so the two sweeps may differ in theirs
  c->Kp_last = Kp; c->dpol_rs = false;
  constexpr int LDA = NT * R;
  // one ring slot per income state (compile-time slot addresses, hank_tangent_tma.cuh) where the n_e chunks fit
  if constexpr (bt_ring_ne_smem<NE, LDA, L>() <= 232448 && L != 6 && L != 3)
    if (!c->no_tma && !c->no_ring_ne && dvalT == nullptr && dvf == nullptr && bt_ring_ne_smem<NE, LDA, L>() <= (size_t)c->smem_max) {
      const size_t smem_r = bt_ring_ne_smem<NE, LDA, L>();
      // next to a backward primal sweep still in flight (pipelined linearisation): start once its CTAs are resident
      // and follow its progress counters; everything after this sweep is ordered behind the primal's completion
      const bool pipe = c->bp_pending && c->bp_flags;
      // A pass of this kind can follow a primal sweep in flight.  Whether the NEXT linearisation should be pipelined
      // (tangent_pass turns this into pipe_hint) depends on what bounds the step.  More than one wave of CTAs: the
      // backward tangent sweep is the long pole — overlap it with the backward primal, and launch both primal sweeps as
      // one kernel so that the cluster keeps its SMs in between (2).  About one wave: the forward primal cannot start
      // until tangent CTAs leave; overlapping the backward primal moves everything 0.7 ms earlier (1).  Seed-horizon
      // passes (sorted, short) and passes that leave SMs free anyway are bounded by backward primal -> forward primal
      // -> forward tangent: serial order, no counters, no extra stream hops (0; measured 0.1 ms slower otherwise).
      c->last_bt_mode = c->pass_thi ? 0 : grid > c->sm_count ? 2 : grid > c->sm_count - 2 * NE ? 1 : 0;
      const int* flags = pipe ? c->d_bpflag : nullptr;
      if (pipe) { int rcw = cuda_check(c, cudaStreamWaitEvent(c->stream, c->ev_bps, 0), "cudaStreamWaitEvent(ev_bps)"); if (rcw) return rcw; }
      else { int rcw = join_bp(c); if (rcw) return rcw; }
      auto launch = [&]() -> int {
        if (c->pass_thi)
          HANK_LAUNCH(KIND_BT, (k_backward_tangent_ring_ne<NE, R, NT, L, true>), grid, NT, smem_r, M, c->tape, K, Kp, c->pass_thi, dr, dw, dpol, flags);
        HANK_LAUNCH(KIND_BT, (k_backward_tangent_ring_ne<NE, R, NT, L, false>), grid, NT, smem_r, M, c->tape, K, Kp, (const int*)nullptr, dr, dw, dpol, flags);
      };
      const int rcl = launch();
      if (pipe) { int rcw = join_bp(c); if (rcw) return rcw; }
      return rcl;
    }
  { int rcw = join_bp(c); if (rcw) return rcw; }
  if constexpr (LDA <= 1024) if (!c->no_tma) {  // TMA-staged tape ring (hank_tangent_tma.cuh)
    const size_t slot = bw_chunk_bytes<LDA>();
    const size_t fixed = (size_t)2 * L * LDA * 8 + (size_t)2 * L * P * 8 + (size_t)((P + 1) & ~1) * 8 + 16 * 8 + 128;
    int S = fixed < (size_t)c->smem_max ? (int)(((size_t)c->smem_max - fixed) / slot) : 0;
    if (S > 8) S = 8;
    if (S >= 2) {
      const size_t smem_t = fixed + (size_t)S * slot;
      if constexpr (L != 6 && L != 3) if (c->pass_thi)   // (horizon passes never run the wide shapes backward)
        HANK_LAUNCH(KIND_BT, (k_backward_tangent_tma<NE, R, NT, L, true>), grid, NT, smem_t, M, c->tape, K, Kp, S, c->pass_thi, dr, dw, dvalT, dpol, dvf);
      HANK_LAUNCH(KIND_BT, (k_backward_tangent_tma<NE, R, NT, L, false>), grid, NT, smem_t, M, c->tape, K, Kp, S, (const int*)nullptr, dr, dw, dvalT, dpol, dvf);
    }
  }
  const size_t smem = (size_t)2 * L * LDA * sizeof(double);
  HANK_LAUNCH(KIND_BT, (k_backward_tangent<NE, R, NT, L>), grid, NT, smem, M, c->tape, K, Kp, c->pass_thi, dr, dw, dvalT, dpol, dvf);
}
template <int NE, int R, int NT, int L>
static int ft_launch(hank_ctx* c, int P, int K, const double* dpol, double* dkdpart) {
  const Consts<NE> M = make_consts<NE>(c, P);
  const int grid = (K + L - 1) / L;
  const int Kp = c->pass_Kp ? c->pass_Kp : (K + kThiGroup - 1) / kThiGroup * kThiGroup;
  constexpr int LDA = NT * R;
  if constexpr (LDA <= 1024) if (!c->no_tma) {
    const size_t slot = fw_chunk_bytes<LDA>() + (size_t)L * LDA * 8;
    const size_t fixed = (size_t)4 * L * LDA * 8 + 16 * 8 + 128;
    int S = fixed < (size_t)c->smem_max ? (int)(((size_t)c->smem_max - fixed) / slot) : 0;
    if (S > 8) S = 8;
    if (S >= 2) {
      const size_t smem_t = fixed + (size_t)S * slot;
      if (c->pass_thi)
        HANK_LAUNCH(KIND_FT, (k_forward_tangent_tma<NE, R, NT, L, true>), grid, NT, smem_t, M, c->tape, K, Kp, S, c->pass_thi, (const double*)c->d_zero, dpol, nullptr, dkdpart, nullptr);
      HANK_LAUNCH(KIND_FT, (k_forward_tangent_tma<NE, R, NT, L, false>), grid, NT, smem_t, M, c->tape, K, Kp, S, (const int*)nullptr, (const double*)nullptr, dpol, nullptr, dkdpart, nullptr);
    }
  }
  const size_t smem = (size_t)4 * L * LDA * sizeof(double);
  HANK_LAUNCH(KIND_FT, (k_forward_tangent<NE, R, NT, L>), grid, NT, smem, M, c->tape, K, Kp, c->pass_thi, dpol, nullptr, dkdpart, nullptr);
}

#define TANGENT_DISPATCH(CFG, FN, ...)                                                    \
  do {                                                                                    \
    const TangentCfg g = CFG;                                                             \
    if (g.NT == 256 && g.R == 1) {                                                        \
      if (g.L == 4) return FN<NE, 1, 256, 4>(__VA_ARGS__);                                \
      if (g.L == 2) return FN<NE, 1, 256, 2>(__VA_ARGS__);                                \
      return FN<NE, 1, 256, 1>(__VA_ARGS__);                                              \
    }                                                                                     \
    if (g.NT == 256 && g.R == 2) return FN<NE, 2, 256, 6>(__VA_ARGS__);                   \
    if (g.NT == 256 && g.R == 4) return FN<NE, 4, 256, 3>(__VA_ARGS__);                   \
    if (g.R == 1) {                                                                       \
      if (g.L == 4) return FN<NE, 1, 512, 4>(__VA_ARGS__);                                \
      if (g.L == 2) return FN<NE, 1, 512, 2>(__VA_ARGS__);                                \
      return FN<NE, 1, 512, 1>(__VA_ARGS__);                                              \
    }                                                                                     \
    if (g.R == 2) {                                                                       \
      if (g.L == 2) return FN<NE, 2, 512, 2>(__VA_ARGS__);                                \
      return FN<NE, 2, 512, 1>(__VA_ARGS__);                                              \
    }                                                                                     \
    return FN<NE, 4, 512, 1>(__VA_ARGS__);                                                \
  } while (0)

template <int NE>
int Sweeps<NE>::backward_tangent(hank_ctx* c, int P, int K, const double* dr, const double* dw,
                                 const double* dvalT, double* dpol, double* dvf) {
  if (c->lda > 2048) return set_error(c, 1, "n_a > 2048 is not supported");
  // With seed horizons (descending over the lanes) a CTA's work shrinks with its index, so the block
  // scheduler packs the short CTAs behind the long ones: 4-lane CTAs finish in one full-length sweep where
  // the 6-lane shape would need one (slower) full-length wave as well.
  // single-step callers (hank_egm_step, hank_vfi) seed V̇ and read it back: one-CTA kernels only
  const bool allow_rs = dvalT == nullptr && dvf == nullptr;
  TangentCfg cfg = tangent_cfg<NE>(c, K, allow_rs);
  c->last_bt_mode = 0;
  if (cfg.NC > 0) {
    { int rcw = join_bp(c); if (rcw) return rcw; }
    const int rc = Sweeps<NE>::backward_tangent_rs(c, cfg.NC, cfg.NT, cfg.L, cfg.GC, P, K, dr, dw, dpol);
    if (rc >= 0) return rc;
    return set_error(c, 1, "row-split cluster launch of the backward tangent sweep failed (cluster not schedulable)");
  }
  if (c->pass_thi && c->lda == 512 && cfg.L == 6) cfg = {512, 1, 4};
  if (c->pass_thi && c->lda == 1024 && cfg.L == 3) cfg = {512, 2, 2};
  TANGENT_DISPATCH(cfg, bt_launch, c, P, K, dr, dw, dvalT, dpol, dvf);
}
template <int NE>
int Sweeps<NE>::forward_tangent(hank_ctx* c, int P, int K, const double* dpol, double* dkdpart, int* nw_out) {
  if (c->lda > 2048) return set_error(c, 1, "n_a > 2048 is not supported");
  const TangentCfg cfg = tangent_cfg<NE>(c, K, true, true);
  if (cfg.NC > 0) {
    *nw_out = cfg.NC;
    const int rc = Sweeps<NE>::forward_tangent_rs(c, cfg.NC, cfg.NT, cfg.L, cfg.GC, P, K, dpol, dkdpart);
    if (rc >= 0) return rc;
    return set_error(c, 1, "row-split cluster launch of the forward tangent sweep failed (cluster not schedulable)");
  }
  *nw_out = cfg.NT / 32;
  TANGENT_DISPATCH(cfg, ft_launch, c, P, K, dpol, dkdpart);
}

}  // namespace hank
