// hank_ctx.h — the context object behind the C ABI (include/hankb200.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "hank_kernels.cuh"

struct hank_ctx {
  int device = 0;
  int n_a = 0, n_e = 0, T = 0, P = 0, G = 0;   // n_e: income states the kernels run with (>= ne_user, see hank_ctx_create)
  int ne_user = 0;              // income states of the caller's arrays; G = n_a * ne_user
  int lda = 0, Gp = 0;          // padded leading dimension NT*R >= n_a and n_e*lda
  int P_alloc = 0;              // periods the tape / paths are allocated for
  double beta = 0, gamma = 0, bc = 0;
  std::vector<double> h_grid, h_z, h_Pi;  // h_Pi row-major [e][e2]
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // high-priority side stream: forward primal sweep overlapping the backward tangent
  cudaStream_t stream3 = nullptr;   // copy stream: tangent seeds upload overlapping the primal sweep
  cudaEvent_t ev_bp = nullptr, ev_fp = nullptr, ev_v = nullptr, ev_x = nullptr;   // ev_x: overflow lanes of a Jacobian pass done (stream3)
  bool fp_pending = false, no_overlap = false;
  // Pipelined linearisation: the backward primal sweep runs on stream2 and publishes, per income state, how many
  // periods of the backward tape are complete (d_bpflag); a backward tangent sweep of the one-CTA ring kernel may start
  // as soon as the primal's CTAs are resident (ev_bps, a launch-completion event) and follows it through the flags.
  // Every other reader of the backward tape waits for ev_bpd (join_bp).
  int* d_bpflag = nullptr;
  cudaEvent_t ev_bps = nullptr, ev_bpd = nullptr, next_launch_ev = nullptr;
  bool bp_pending = false, bp_flags = false, bp_pipe_req = false, no_pipe = false;
  // What the last backward tangent launch says about the next linearisation (pipe_hint, set by tangent_pass from the main
  // wave's last_bt_mode): 0 serial order; 1 backward primal on the side stream with progress counters, tangent sweep next
  // to it; 2 the same with both primal sweeps in one launch (k_primal_ds_both)
  int last_bt_mode = 0, pipe_hint = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int smem_max = 0, sm_count = 0;

  double *d_grid = nullptr, *d_valueT = nullptr, *d_D0 = nullptr;
  double *d_Pi = nullptr, *d_scatter = nullptr;   // Π row-major; 3 grids of scratch for the scatter lottery
  double *d_r = nullptr, *d_w = nullptr;
  hank::Tape tape{};
  bool have_terminal = false, have_D0 = false, have_backward = false, have_forward = false;

  // tangent lanes
  int Kcap = 0, K_last = 0;
  double *d_dr = nullptr, *d_dw = nullptr, *d_dpol = nullptr;
  double *d_dvalT = nullptr, *d_dvalue_first = nullptr;  // egm_step / vfi lanes
  int egm_cap = -1;              // lanes the egm scratch is sized for
  double *d_kdpart = nullptr, *d_KD = nullptr, *d_dkdpart = nullptr, *d_dKD = nullptr;
  int* d_status = nullptr;
  int* h_status = nullptr;  // pinned

  // Krusell-Smith residual layer
  bool ks_ready = false, linearized = false;
  double alpha = 0, delta = 0, ssKS = 0;
  double *d_x = nullptr, *d_Z = nullptr, *d_F = nullptr, *d_V = nullptr, *d_JV = nullptr;
  int Vcap = 0;
  double *h_pin = nullptr; size_t h_pin_bytes = 0;  // pinned staging for host-pointer calls

  // equilibrium equations as device bytecode (hank_eq_configure, hank_eq.cu); off: the built-in Krusell-Smith block
  bool eq_on = false;
  int n_endog = 4, n_exog = 1, eq_ir = 2, eq_iw = 3;   // x is n_endog x P; rows of r and w; Z is n_exog x P
  int *d_eq_off = nullptr, *d_eq_code = nullptr;
  double *d_eq_consts = nullptr, *d_eq_ss = nullptr;

  // dense solve (Newton)
  void* solver = nullptr;  // cusolverDnHandle_t
  double *d_Jinv = nullptr, *d_newton = nullptr, *d_lu_work = nullptr;
  int* d_newton_i = nullptr;   // pivots + info
  size_t newton_bytes = 0, jinv_bytes = 0, lu_work_bytes = 0;
  // cache of the preconditioner (J̅ or J̅⁻¹ in d_Jinv) across hank_newton_solve calls, keyed on a hash of J̅
  bool jbar_valid = false, jbar_inverse = false; uint64_t jbar_key = 0; int jbar_n = 0;

  // NCCL
  void* nccl_comm = nullptr;
  double* d_gather = nullptr; size_t gather_bytes = 0;   // scratch of the host-pointer all-gather
  int nranks = 1, rank = 0;

  std::string err;
  int64_t launches = 0;

  // per-kernel CUDA-event timing of the sweep kernels (hank_profile / hank_kernel_times)
  bool profile = false;
  int* d_jac_idx = nullptr; int jac_idx_cap = 0;   // lane<->column maps of hank_ks_jacobian_columns
  bool no_cluster = false;       // HANK_NO_CLUSTER=1: single-CTA primal sweeps
  bool no_dsmem = false;         // HANK_NO_DSMEM=1: cluster primal sweeps exchange through global memory
  bool fp_cluster = false;       // last forward primal ran on the cluster (per-column KD partials)
  double* d_xch = nullptr;       // [2][NE][lda] cluster exchange buffer
  bool no_wide = false;          // HANK_NO_WIDE=1: never use the 256-thread / 6-lane tangent shape
  bool no_tma = false;           // HANK_NO_TMA=1: use the register-prefetch tangent kernels
  bool no_skip = false;          // HANK_NO_SKIP=1: unit-seed Jacobian lanes sweep all periods
  bool no_ring_ne = false;       // HANK_NO_RING_NE=1: backward tangent always through the runtime-sized ring
  bool no_rowsplit = false;      // HANK_NO_ROWSPLIT=1: never split a lane group's rows over a cluster
  bool no_rs_st = false, no_rs_push = false, no_rs_ce = false;   // HANK_NO_RS_ST / _PUSH / _CE=1: one-lane row-split sweeps without the st.async / bulk-push / thread-per-(e,row) kernels
  bool rs_relaxed = false;       // HANK_RS_RELAXED=1: relaxed cluster hand-shakes in the row-split kernels (racy; A/B only)
  bool rs_no_multi = false;      // HANK_RS_NO_MULTI=1: no multi-lane row-split clusters (mid lane counts run one CTA per lane)
  int rs_max_k = 0;              // HANK_RS_MAXK: lane count up to which 1-lane row-split clusters are used (0: sm_count / NC)
  int rs_cap[2] = {-1, -1};      // resident clusters of the 1-lane / 4-lane row-split shape (-1: not asked yet)
  // row-block-major copies of the tape for the row-split kernels, and the layout of the policy tangents in d_dpol
  unsigned char *tape_rs_bw = nullptr, *tape_rs_fw = nullptr;
  int tape_rs_bw_nt = 0, tape_rs_fw_nt = 0;   // rows per block the copies were made for (0: stale; a pass with another
                                              // cluster shape at the same linearisation has to remake them)
  bool dpol_rs = false; int dpol_rs_L = 0, dpol_rs_NC = 0, dpol_rs_ncl = 0;   // d_dpol is [t][cluster][rank][e][l][NT]
  int Kp_last = 0;               // lane stride of the policy tangents written by the last backward tangent sweep
  // Seed horizons of the pass in flight (hank_ks_jacobian_columns): lanes come in groups of kThiGroup
  // whose seeds are zero from period pass_thi[group] on, so the backward tangent starts there and the
  // forward tangent reads zeros instead of policy tangents beyond it.  Null for generic seeds.
  const int* pass_thi = nullptr;
  int pass_Kp = 0;               // lane stride of the pass (0: lanes per CTA decide)
  int* d_thi = nullptr; int thi_cap = 0;
  double* d_zero = nullptr;      // 32 KB of zeros: TMA source for policy tangents beyond the horizon
  struct Rec { int kind; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> ev_pool;
  double kern_ms[4] = {0, 0, 0, 0};
  int64_t kern_n[4] = {0, 0, 0, 0};
};

namespace hank {

struct Shape {  // launch configuration derived from (n_a, K, smem)
  int NT, R;
};
inline bool pick_shape(int n_a, Shape* s) {
  if (n_a <= 256) { *s = {256, 1}; return true; }
  if (n_a <= 512) { *s = {512, 1}; return true; }
  if (n_a <= 1024) { *s = {512, 2}; return true; }
  if (n_a <= 2048) { *s = {512, 4}; return true; }
  return false;
}

// Per-n_e launchers (explicitly instantiated in hank_ne*.cu). All return a hank_status.
template <int NE>
struct Sweeps {
  static int backward_primal(hank_ctx* c, int P, const double* valueT, const double* r, const double* w);
  static int backward_tangent(hank_ctx* c, int P, int K, const double* dr, const double* dw,
                              const double* dvalT, double* dpol, double* dvalue_first);
  static int forward_primal(hank_ctx* c, int P, const double* D0, const double* pol, double* KD);
  // both primal sweeps in one cluster launch with progress counters (pipelined linearisation); -1: shape not available
  static int primal_both(hank_ctx* c, int P, const double* valueT, const double* r, const double* w, const double* D0);
  static int forward_tangent(hank_ctx* c, int P, int K, const double* dpol, double* dkdpart, int* nw_out);
  static int lanes_per_cta(hank_ctx* c, int K);
  // rows of a lane group split over a cluster of NC CTAs (hank_tangent_rowsplit.cuh); -1: shape not available
  static int backward_tangent_rs(hank_ctx* c, int NC, int NT, int L, int GC, int P, int K, const double* dr,
                                 const double* dw, double* dpol);
  static int forward_tangent_rs(hank_ctx* c, int NC, int NT, int L, int GC, int P, int K, const double* dpol,
                                double* dkdpart);
  static int rs_max_clusters(hank_ctx* c, int NC, int NT, int L, int GC);   // clusters of that shape resident at once
};

enum { KIND_BP = 0, KIND_BT = 1, KIND_FP = 2, KIND_FT = 3 };
cudaEvent_t prof_begin(hank_ctx* c);
void prof_end(hank_ctx* c, int kind, cudaEvent_t a);
int set_error(hank_ctx* c, int code, const std::string& msg);
int cuda_check(hank_ctx* c, cudaError_t e, const char* what);
// the backward primal sweep of a pipelined linearisation (stream2) must be complete before c->stream goes on
static inline int join_bp(hank_ctx* c) {
  if (!c->bp_pending) return 0;
  c->bp_pending = false;
  return cuda_check(c, cudaStreamWaitEvent(c->stream, c->ev_bpd, 0), "cudaStreamWaitEvent(ev_bpd)");
}
void newton_release(hank_ctx* c);   // destroys the cuSOLVER handle (hank_newton.cu)
// generic equations (hank_eq.cu)
int eq_extract_rw(hank_ctx* c, const double* x, double* r, double* w);
int eq_extract_drdw(hank_ctx* c, int K, const double* V, double* dr, double* dw);
int eq_residual(hank_ctx* c, const double* x, const double* KD, const double* Z, double* F);
// V: [K][n] seeds, or null with unit_cols[K] (unit seed e_col per lane); dKD rows: lane l, or col_lane[l] (-1: none)
int eq_residual_tangent(hank_ctx* c, int K, const double* x, const double* KD, const double* Z, const double* V,
                        const int* unit_cols, const double* dKD, const int* col_lane, double* JV);
void eq_release(hank_ctx* c);

}  // namespace hank
