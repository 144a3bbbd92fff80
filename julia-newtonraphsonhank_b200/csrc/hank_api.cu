// hank_api.cu — context management and the C ABI of include/hankb200.h.
// Reference functions replaced are cited in the header next to each declaration.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "hank_ctx.h"
#include "../../include/hankb200.h"

namespace hank {

int set_error(hank_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_check(hank_ctx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return HANK_OK;
  return set_error(c, HANK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

cudaEvent_t prof_begin(hank_ctx* c) {
  if (!c->profile) return nullptr;
  cudaEvent_t e = nullptr;
  if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
  else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  cudaEventRecord(e, c->stream);
  return e;
}
void prof_end(hank_ctx* c, int kind, cudaEvent_t a) {
  if (!c->profile || !a) return;
  cudaEvent_t e = nullptr;
  if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
  else if (cudaEventCreate(&e) != cudaSuccess) { c->ev_pool.push_back(a); return; }
  cudaEventRecord(e, c->stream);
  c->recs.push_back({kind, a, e});
}

#define CK(call)                                                   \
  do {                                                             \
    int rc__ = hank::cuda_check(c, (call), #call);                 \
    if (rc__) return rc__;                                         \
  } while (0)
#define RC(call)                 \
  do {                           \
    int rc__ = (call);           \
    if (rc__) return rc__;       \
  } while (0)

// ---- small kernels ---------------------------------------------------------------------
__global__ void k_extract_rw(const double* __restrict__ x, int P, double* r, double* w) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P) { r[t] = x[4 * t + 2]; w[t] = x[4 * t + 3]; }
}
// 1/(1+r): what the backward primal sweep writes per period, in place before a tangent sweep starts next to it
__global__ void k_fill_rho(const double* __restrict__ r, int P, double* rho) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P) { const double opr = 1.0 + r[t]; rho[t] = 1.0 / opr; }
}
// V: n x K column-major (lane-major); dr, dw: [K][P]
__global__ void k_extract_drdw(const double* __restrict__ V, int P, int K, double* dr, double* dw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * P) {
    int l = i / P, t = i - l * P;
    const double* v = V + (size_t)l * 4 * P;
    dr[i] = v[4 * t + 2]; dw[i] = v[4 * t + 3];
  }
}
__global__ void k_reduce_partials(const double* __restrict__ part, int NW, int count, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    double s = 0.0;
    for (int w = 0; w < NW; ++w) s += part[(size_t)i * NW + w];
    out[i] = s;
  }
}
// Compiled KS residuals (KrusellSmith.yaml:90-94; ModelParser.jl:217-259), equation-fastest.
__global__ void k_ks_residual(int P, double alpha, double delta, double ssKS, const double* __restrict__ x,
                              const double* __restrict__ KD, const double* __restrict__ Z, double* F) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P) return;
  const double Y = x[4 * t], KS = x[4 * t + 1], r = x[4 * t + 2], w = x[4 * t + 3];
  const double Kl = t == 0 ? ssKS : x[4 * (t - 1) + 1];
  const double Ka = pow(Kl, alpha), Ka1 = pow(Kl, alpha - 1.0);
  F[4 * t + 0] = Y - (Z[t] * Ka);
  F[4 * t + 1] = (r + delta) - ((alpha * Z[t]) * Ka1);
  F[4 * t + 2] = w - (((1.0 - alpha) * Z[t]) * Ka);
  F[4 * t + 3] = KS - KD[t];
}
// Tangent lanes of the residuals: JV[:, l] for V[:, l] and K̇D[l] (dKD: [K][P]).
__global__ void k_ks_residual_tangent(int P, int K, double alpha, double ssKS, const double* __restrict__ x,
                                      const double* __restrict__ Z, const double* __restrict__ V,
                                      const double* __restrict__ dKD, double* JV) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * P) return;
  const int l = i / P, t = i - l * P;
  const double Kl = t == 0 ? ssKS : x[4 * (t - 1) + 1];
  const double Ka1 = pow(Kl, alpha - 1.0), Ka2 = pow(Kl, alpha - 2.0);
  const double* d = V + (size_t)l * 4 * P;
  const double dKl = t == 0 ? 0.0 : d[4 * (t - 1) + 1];
  const double dKa = dKl * alpha * Ka1, dKa1 = dKl * (alpha - 1.0) * Ka2;
  double* o = JV + (size_t)l * 4 * P;
  o[4 * t + 0] = d[4 * t + 0] - (Z[t] * dKa);
  o[4 * t + 1] = d[4 * t + 2] - ((alpha * Z[t]) * dKa1);
  o[4 * t + 2] = d[4 * t + 3] - (((1.0 - alpha) * Z[t]) * dKa);
  o[4 * t + 3] = d[4 * t + 1] - dKD[(size_t)l * P + t];
}
// Unit seeds for Jacobian columns: lane l perturbs r or w (v = 2, 3) at period tcol.
__global__ void k_unit_seeds(int P, int K, int ne, int ir, int iw, const int* __restrict__ lane_col, double* dr, double* dw) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= K) return;
  const int col = lane_col[l], t = col / ne, v = col - t * ne;
  if (v == ir) dr[(size_t)l * P + t] = 1.0;
  if (v == iw) dw[(size_t)l * P + t] = 1.0;
}
// Jacobian columns with unit seeds e_col: direct residual terms + household term −K̇D.
__global__ void k_ks_jac_columns(int P, const int* __restrict__ col_ids, int ncols, double alpha, double ssKS,
                                 const double* __restrict__ x, const double* __restrict__ Z,
                                 const int* __restrict__ col_lane, const double* __restrict__ dKD, double* J) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols * P) return;
  const int c = i / P, t = i - c * P;
  const int col = col_ids[c], tc = col >> 2, v = col & 3;
  double* o = J + (size_t)c * 4 * P;
  double o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  if (t == tc) { if (v == 0) o0 = 1.0; if (v == 2) o1 = 1.0; if (v == 3) o2 = 1.0; if (v == 1) o3 = 1.0; }
  if (v == 1 && t == tc + 1) {  // KS(-1) terms
    const double Kl = x[4 * tc + 1];
    const double Ka1 = pow(Kl, alpha - 1.0), Ka2 = pow(Kl, alpha - 2.0);
    const double dKa = 1.0 * alpha * Ka1, dKa1 = 1.0 * (alpha - 1.0) * Ka2;
    o0 = 0.0 - (Z[t] * dKa); o1 = 0.0 - ((alpha * Z[t]) * dKa1); o2 = 0.0 - (((1.0 - alpha) * Z[t]) * dKa);
  }
  const int l = col_lane[c];
  if (l >= 0) o3 -= dKD[(size_t)l * P + t];
  o[4 * t + 0] = o0; o[4 * t + 1] = o1; o[4 * t + 2] = o2; o[4 * t + 3] = o3;
  (void)ssKS;
}
// make_endogenous_transition's bracket rule (ForwardIteration.jl:46-75).
__global__ void k_lottery(const double* __restrict__ grid, int n_a, int G, const double* __restrict__ pol,
                          int32_t* m_out, double* om_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G) return;
  const double p = pol[i];
  const int m = lower_bound(grid, n_a, p) + 1;
  double om = 1.0;
  if (m > 1 && m <= n_a) om = (p - grid[m - 2]) / (grid[m - 1] - grid[m - 2]);
  m_out[i] = m;
  if (om_out) om_out[i] = om;
}

// Scatter formulation of the forward sweep for policies that are NOT monotone in a (the reference's
// make_endogenous_transition accepts any policy, ForwardIteration.jl:37-78).  One CTA, FP64 global atomics,
// runtime n_e; correctness path only (the gather kernels reject such policies).  D, tmp, Dn: [n_e][lda].
__global__ void __launch_bounds__(1024, 1)
k_forward_scatter(int n_a, int n_e, int lda, int P, const double* __restrict__ grid, const double* __restrict__ Pi,
                  const double* __restrict__ D0, const double* __restrict__ pol, int fw_chunk_bytes, unsigned char* fw,
                  int* mbr, double* scratch, double* KD) {
  __shared__ double red[33];
  const int Gp = n_e * lda, tid = threadIdx.x, NT = blockDim.x;
  double* D = scratch; double* tmp = scratch + Gp; double* Dn = scratch + 2 * Gp;
  for (int i = tid; i < Gp; i += NT) D[i] = D0[i];
  __syncthreads();
  for (int t = 0; t < P; ++t) {
    for (int i = tid; i < Gp; i += NT) tmp[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < Gp; i += NT) {
      const int e = i / lda, a = i - e * lda;
      if (a >= n_a) continue;
      const double p = pol[(size_t)t * Gp + i], d = D[i];
      const int m = lower_bound(grid, n_a, p) + 1;
      mbr[(size_t)t * Gp + i] = m;
      if (m == 1) atomicAdd(&tmp[e * lda], d);
      else if (m > n_a) atomicAdd(&tmp[e * lda + n_a - 1], d);
      else {
        const double om = (p - grid[m - 2]) / (grid[m - 1] - grid[m - 2]);
        atomicAdd(&tmp[e * lda + m - 2], (1.0 - om) * d);
        atomicAdd(&tmp[e * lda + m - 1], om * d);
      }
    }
    __syncthreads();
    double kacc = 0.0;
    for (int i = tid; i < Gp; i += NT) {
      const int e2 = i / lda, a = i - e2 * lda;
      if (a >= n_a) continue;
      double d = 0.0;
      for (int e = 0; e < n_e; ++e) d += Pi[e * n_e + e2] * tmp[e * lda + a];
      Dn[i] = d;
      reinterpret_cast<double*>(fw + ((size_t)t * n_e + e2) * fw_chunk_bytes)[FW_D * lda + a] = d;
      kacc += pol[(size_t)t * Gp + i] * d;
    }
    kacc = warp_sum(kacc);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = kacc;
    __syncthreads();
    if (tid == 0) { double s = 0.0; for (int w = 0; w < NT / 32; ++w) s += red[w]; KD[t] = s; }
    double* sw = D; D = Dn; Dn = sw;
    __syncthreads();
  }
}

// ---- n_e dispatch ----------------------------------------------------------------------
#define NE_DISPATCH(c, CALL)                                                              \
  switch ((c)->n_e) {                                                                     \
    case 3: return Sweeps<3>::CALL;                                                       \
    case 5: return Sweeps<5>::CALL;                                                       \
    case 7: return Sweeps<7>::CALL;                                                       \
    case 9: return Sweeps<9>::CALL;                                                       \
    case 11: return Sweeps<11>::CALL;                                                     \
    default: return set_error(c, HANK_ERR_ARG, "n_e must be one of 3, 5, 7, 9, 11 in this build"); \
  }
static int sw_backward_primal(hank_ctx* c, int P, const double* vT, const double* r, const double* w) {
  NE_DISPATCH(c, backward_primal(c, P, vT, r, w));
}
static int sw_backward_tangent(hank_ctx* c, int P, int K, const double* dr, const double* dw,
                               const double* dvalT, double* dpol, double* dvf) {
  NE_DISPATCH(c, backward_tangent(c, P, K, dr, dw, dvalT, dpol, dvf));
}
static int sw_primal_both(hank_ctx* c, int P, const double* vT, const double* r, const double* w, const double* D0) {
  NE_DISPATCH(c, primal_both(c, P, vT, r, w, D0));
}
static int sw_forward_primal(hank_ctx* c, int P, const double* D0, const double* pol, double* KD) {
  NE_DISPATCH(c, forward_primal(c, P, D0, pol, KD));
}
static int sw_forward_tangent(hank_ctx* c, int P, int K, const double* dpol, double* dkdpart, int* nw) {
  NE_DISPATCH(c, forward_tangent(c, P, K, dpol, dkdpart, nw));
}
static int sw_lanes_per_cta(hank_ctx* c, int K) { NE_DISPATCH(c, lanes_per_cta(c, K)); }
// lane stride of the policy-tangent array for a K-lane pass: K rounded up to a multiple of every lanes-per-CTA count
static int lane_stride(hank_ctx* c, int K) {
  (void)c;
  return (K + kThiGroup - 1) / kThiGroup * kThiGroup;   // what bt_launch / ft_launch use (hank_launch.cuh)
}
static inline size_t bw_chunk(const hank_ctx* c) { return (size_t)52 * c->lda; }
static inline size_t fw_chunk(const hank_ctx* c) { return (size_t)FW_NF * 8 * c->lda + (size_t)4 * (c->lda + 4); }

template <typename T>
static int dalloc(hank_ctx* c, T** p, size_t count) {
  return cuda_check(c, cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)), "cudaMalloc");
}
template <typename T>
static void dfree(T*& p) { if (p) cudaFree(p); p = nullptr; }

// Work queued on the side stream (forward primal sweep + residuals of the last linearisation) must be
// complete before anything on the main stream reads its results or overwrites its inputs.
static int join_side(hank_ctx* c) {
  if (c->fp_pending) {
    CK(cudaStreamWaitEvent(c->stream, c->ev_fp, 0));
    c->fp_pending = false;
    c->bp_pending = false;   // (the forward sweep was queued behind the backward primal sweep on the same stream)
  }
  return join_bp(c);
}

static int check_status(hank_ctx* c) {
  RC(join_side(c));
  CK(cudaMemcpyAsync(c->h_status, c->d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int code = c->h_status[0];
  if (code == 0) return HANK_OK;
  const int a = c->h_status[1], e = c->h_status[2], t = c->h_status[3];
  CK(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(int), c->stream));
  char buf[256];
  const char* what = code == HANK_ERR_DOMAIN ? "DomainError: negative base under a non-integer power"
                   : code == HANK_ERR_KNOTS ? "knot-vectors must be unique and sorted in increasing order"
                   : code == HANK_ERR_NONMONOTONE ? "policy is not monotone in a (gather lottery not applicable)"
                   : "device-side failure";
  snprintf(buf, sizeof buf, "%s at period t=%d, a=%d, e=%d", what, t, a, e);
  return set_error(c, code, buf);
}

static int ensure_lanes(hank_ctx* c, int K) {
  if (K <= c->Kcap) return HANK_OK;
  const size_t per_lane = (size_t)c->P_alloc * c->Gp * sizeof(double);
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t have = (size_t)c->Kcap * per_lane;
  const long long fit = (long long)((0.85 * (double)(free_b + have)) / (double)per_lane) - 16;   // 16 lanes of stride padding
  const int Kmax = (int)std::min<long long>((long long)K, std::max<long long>(fit, 0));
  if (Kmax < 1) return set_error(c, HANK_ERR_CUDA, "not enough device memory for one tangent lane");
  if (Kmax <= c->Kcap) return HANK_OK;
  dfree(c->d_dr); dfree(c->d_dw); dfree(c->d_dpol); dfree(c->d_dkdpart); dfree(c->d_dKD);
  c->Kcap = 0;
  RC(dalloc(c, &c->d_dr, (size_t)Kmax * c->P_alloc));
  RC(dalloc(c, &c->d_dw, (size_t)Kmax * c->P_alloc));
  RC(dalloc(c, &c->d_dpol, (size_t)(Kmax + 16) * c->P_alloc * c->Gp));   // lane stride padding: up to 11 lanes
  CK(cudaMemsetAsync(c->d_dpol, 0, (size_t)(Kmax + 16) * c->P_alloc * c->Gp * sizeof(double), c->stream));
  RC(dalloc(c, &c->d_dkdpart, (size_t)Kmax * c->P_alloc * 16));
  RC(dalloc(c, &c->d_dKD, (size_t)Kmax * c->P_alloc));
  c->Kcap = Kmax;
  return HANK_OK;
}
static int ensure_V(hank_ctx* c, int K) {
  if (K <= c->Vcap) return HANK_OK;
  dfree(c->d_V); dfree(c->d_JV);
  c->Vcap = 0;
  const size_t n = (size_t)c->n_endog * c->P;
  RC(dalloc(c, &c->d_V, n * K));
  RC(dalloc(c, &c->d_JV, n * K));
  c->Vcap = K;
  return HANK_OK;
}

static inline int nblk(size_t n, int b = 256) { return (int)((n + b - 1) / b); }

// scratch of the single-step entry points: incoming value + K lanes, outgoing K lanes
static int ensure_egm(hank_ctx* c, int K) {
  if (K <= c->egm_cap) return HANK_OK;
  dfree(c->d_dvalT); dfree(c->d_dvalue_first);
  c->egm_cap = -1;
  RC(dalloc(c, &c->d_dvalT, (size_t)(K + 1) * c->Gp));
  RC(dalloc(c, &c->d_dvalue_first, (size_t)std::max(K, 1) * c->Gp));
  CK(cudaMemsetAsync(c->d_dvalT, 0, (size_t)(K + 1) * c->Gp * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->d_dvalue_first, 0, (size_t)std::max(K, 1) * c->Gp * sizeof(double), c->stream));
  c->egm_cap = K;
  return HANK_OK;
}

// Copies `rows` columns of n_a doubles between the caller's dense [rows][n_a] layout and the
// device's padded [rows][lda] layout.
static int copy_in(hank_ctx* c, double* dst_padded, const double* src_dense, size_t rows) {
  return cuda_check(c, cudaMemcpy2DAsync(dst_padded, (size_t)c->lda * 8, src_dense, (size_t)c->n_a * 8,
                                         (size_t)c->n_a * 8, rows, cudaMemcpyDefault, c->stream), "cudaMemcpy2DAsync");
}
template <typename T>
static int copy_out(hank_ctx* c, T* dst_dense, const T* src_padded, size_t rows) {
  return cuda_check(c, cudaMemcpy2DAsync(dst_dense, (size_t)c->n_a * sizeof(T), src_padded, (size_t)c->lda * sizeof(T),
                                         (size_t)c->n_a * sizeof(T), rows, cudaMemcpyDefault, c->stream), "cudaMemcpy2DAsync");
}

// Padding income states (hank_ctx_create): copies the last real column of `count` consecutive [n_e][lda] arrays into
// their padding columns (values that keep the padding states' own recursion well-posed; they carry no mass).
static int pad_columns(hank_ctx* c, double* arr, size_t count = 1) {
  for (int e = c->ne_user; e < c->n_e; ++e)
    CK(cudaMemcpy2DAsync(arr + (size_t)e * c->lda, (size_t)c->Gp * 8, arr + (size_t)(c->ne_user - 1) * c->lda, (size_t)c->Gp * 8,
                         (size_t)c->lda * 8, count, cudaMemcpyDeviceToDevice, c->stream));
  return HANK_OK;
}
// `count` consecutive caller arrays [ne_user][n_a] <-> device arrays [n_e][lda] (one 2D copy per array when padded)
static int copy_in_grids(hank_ctx* c, double* dst, const double* src, size_t count) {
  if (c->ne_user == c->n_e) return copy_in(c, dst, src, count * c->n_e);
  for (size_t i = 0; i < count; ++i) RC(copy_in(c, dst + i * c->Gp, src + i * c->G, c->ne_user));
  return HANK_OK;
}
template <typename T>
static int copy_out_grids(hank_ctx* c, T* dst, const T* src, size_t count) {
  if (c->ne_user == c->n_e) return copy_out(c, dst, src, count * c->n_e);
  for (size_t i = 0; i < count; ++i) RC(copy_out(c, dst + i * c->G, src + i * c->Gp, c->ne_user));
  return HANK_OK;
}

// Tangent pass over the current tape for lanes given by dr/dw (already on device, [K][P]).
static int tangent_pass_range(hank_ctx* c, int P, int K, int k0, double* dpol) {
  const size_t o = (size_t)k0 * P;
  RC(sw_backward_tangent(c, P, K, c->d_dr + o, c->d_dw + o, nullptr, dpol, nullptr));
  return HANK_OK;
}
static int tangent_pass(hank_ctx* c, int P, int K) {
  // Unit-seed passes a few lanes larger than one wave of 4-lane CTAs (the 598 household lanes of the T = 300
  // Jacobian on 148 SMs): the forward sweep has no short CTAs to pack behind the long ones, so the lanes beyond the
  // wave would cost a second full sweep (or the slower 6-lane shape).  They go through the row-split clusters
  // instead, which finish a handful of lanes in a third of a sweep.  The cut is at a horizon-group boundary.
  int K1 = K;
  if (c->pass_thi && !c->no_rowsplit && c->lda == 512 && K > 4 * c->sm_count) {
    const int wave = 4 * c->sm_count / kThiGroup * kThiGroup;
    (void)sw_lanes_per_cta(c, 1);   // makes sure the cluster capacity has been asked
    if (K - wave <= std::min(c->rs_cap[0], kThiGroup * 2) && K - wave > 0) K1 = wave;
  }
  const int* thi = c->pass_thi;
  const int Kp_all = c->pass_Kp;
  double* dpol2 = c->d_dpol + (size_t)P * c->n_e * (size_t)K1 * c->lda;   // policy tangents of the lanes beyond the cut
  if (K1 == K) {
    RC(tangent_pass_range(c, P, K, 0, c->d_dpol));
    c->pipe_hint = c->last_bt_mode;
    RC(join_side(c));  // the forward tangent needs the forward tape of the linearisation
    int nw = 16;
    RC(sw_forward_tangent(c, P, K, c->d_dpol, c->d_dkdpart, &nw));
    k_reduce_partials<<<nblk((size_t)K * P), 256, 0, c->stream>>>(c->d_dkdpart, nw, K * P, c->d_dKD);
    c->launches++;
    return cuda_check(c, cudaGetLastError(), "k_reduce_partials");
  }
  // The overflow lanes run on the copy stream NEXT TO the main wave instead of behind it: the main backward sweep's
  // CTAs are sorted longest horizon first, so SMs free up from half its duration on, and the handful of 8-CTA
  // clusters does both of its sweeps in that tail (0.25 + 0.39 ms at 500x7) before the main forward sweep needs
  // every SM again.
  cudaStream_t main_s = c->stream;
  if (c->bp_pending) CK(cudaStreamWaitEvent(c->stream3, c->ev_bpd, 0));   // (pipelined linearisation: only the main wave follows the primal sweep)
  CK(cudaEventRecord(c->ev_v, main_s));                   // seeds, horizons and the backward tape are in place
  CK(cudaStreamWaitEvent(c->stream3, c->ev_v, 0));
  c->pass_Kp = K1;
  RC(tangent_pass_range(c, P, K1, 0, c->d_dpol));         // main wave, backward
  c->pipe_hint = c->last_bt_mode;
  double* part2 = c->d_dkdpart + (size_t)K1 * P * 16;
  c->stream = c->stream3;
  c->pass_thi = thi + K1 / kThiGroup;
  int rc = tangent_pass_range(c, P, K - K1, K1, dpol2);   // overflow lanes, backward
  if (rc == HANK_OK && c->fp_pending) rc = cuda_check(c, cudaStreamWaitEvent(c->stream3, c->ev_fp, 0), "cudaStreamWaitEvent");
  int nw2 = 16;
  if (rc == HANK_OK) rc = sw_forward_tangent(c, P, K - K1, dpol2, part2, &nw2);   // overflow lanes, forward
  if (rc == HANK_OK) {
    k_reduce_partials<<<nblk((size_t)(K - K1) * P), 256, 0, c->stream>>>(part2, nw2, (K - K1) * P, c->d_dKD + (size_t)K1 * P);
    c->launches++;
    rc = cuda_check(c, cudaGetLastError(), "k_reduce_partials");
  }
  if (rc == HANK_OK) rc = cuda_check(c, cudaEventRecord(c->ev_x, c->stream3), "cudaEventRecord");
  c->stream = main_s;
  c->pass_thi = thi;
  if (rc != HANK_OK) { c->pass_Kp = Kp_all; return rc; }
  rc = join_side(c);
  int nw = 16;
  if (rc == HANK_OK) rc = sw_forward_tangent(c, P, K1, c->d_dpol, c->d_dkdpart, &nw);   // main wave, forward
  c->pass_Kp = Kp_all;
  RC(rc);
  k_reduce_partials<<<nblk((size_t)K1 * P), 256, 0, c->stream>>>(c->d_dkdpart, nw, K1 * P, c->d_dKD);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamWaitEvent(main_s, c->ev_x, 0));            // the overflow lanes' K̇D before anyone reads d_dKD
  return HANK_OK;
}

}  // namespace hank

using namespace hank;

extern "C" {

const char* hank_version(void) { return "hankb200 0.1 (sm_100a)"; }

int hank_ctx_create(hank_ctx** out, int device, int n_a, int n_e, int T, const double* grid, const double* z,
                    const double* Pi, double beta, double gamma, double borrow_cons) {
  if (!out) return HANK_ERR_ARG;
  *out = nullptr;
  if (n_a < 2 || n_e < 1 || n_e > kMaxNE || T < 2 || !grid || !z || !Pi) return HANK_ERR_ARG;
  hank_ctx* c = new hank_ctx;
  *out = c;  // returned even on failure so hank_last_error() can be read; caller destroys it
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(c, HANK_ERR_CUDA, "no CUDA device available: libhankb200 has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(c, HANK_ERR_ARG, "bad device index");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->device = device;
  c->smem_max = (int)prop.sharedMemPerBlockOptin;
  c->sm_count = prop.multiProcessorCount;
  // The sweep kernels are instantiated for n_e in {3, 5, 7, 9, 11}.  Any other count up to 11 (an even number of income
  // states, or the Kronecker product of two exogenous processes, ForwardIteration.jl:280-284) runs on the next
  // instantiated count with ABSORBING, ZERO-MASS padding states appended: Pi'[pad][pad] = 1, no transitions between real
  // and padding states, income of the last real state, terminal value of the last real state, no initial mass.  The
  // real states' sums only gain exact zeros (x + 0*v with finite v), so their results are bit-identical to an
  // unpadded run; the padding columns never leave the device (the caller's arrays have ne_user columns).
  int ne_run = n_e;
  while (ne_run < 3 || ne_run % 2 == 0) ++ne_run;
  if (ne_run > 11) return set_error(c, HANK_ERR_ARG, "n_e must be between 1 and 11 in this build");
  c->n_a = n_a; c->ne_user = n_e; c->n_e = ne_run; c->T = T; c->P = T - 1; c->G = n_a * n_e; c->P_alloc = T - 1;
  c->beta = beta; c->gamma = gamma; c->bc = borrow_cons;
  c->h_grid.assign(grid, grid + n_a);
  c->h_z.assign(z, z + n_e);
  c->h_z.resize(ne_run, z[n_e - 1]);
  c->h_Pi.assign((size_t)ne_run * ne_run, 0.0);
  for (int e = 0; e < n_e; ++e)
    for (int e2 = 0; e2 < n_e; ++e2) c->h_Pi[(size_t)e * ne_run + e2] = Pi[e + (size_t)n_e * e2];
  for (int e = n_e; e < ne_run; ++e) c->h_Pi[(size_t)e * ne_run + e] = 1.0;
  for (int a = 1; a < n_a; ++a)
    if (!(grid[a] > grid[a - 1])) return set_error(c, HANK_ERR_ARG, "grid must be strictly increasing");
  n_e = ne_run;   // everything below sizes device storage
  Shape s;
  if (!pick_shape(n_a, &s)) return set_error(c, HANK_ERR_ARG, "n_a > 2048 is not supported in this build");
  c->lda = s.NT * s.R; c->Gp = n_e * c->lda;
  { const char* nt = getenv("HANK_NO_TMA"); c->no_tma = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_WIDE"); c->no_wide = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_CLUSTER"); c->no_cluster = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_DSMEM"); c->no_dsmem = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_SKIP"); c->no_skip = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_ROWSPLIT"); c->no_rowsplit = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_RING_NE"); c->no_ring_ne = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_RS_MAXK"); c->rs_max_k = nt ? atoi(nt) : 0; }
  { const char* nt = getenv("HANK_RS_RELAXED"); c->rs_relaxed = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_RS_ST"); c->no_rs_st = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_RS_PUSH"); c->no_rs_push = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_NO_RS_CE"); c->no_rs_ce = nt && nt[0] == '1'; }
  { const char* nt = getenv("HANK_RS_NO_MULTI"); c->rs_no_multi = nt && nt[0] == '1'; }
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->ev_bp, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_bps, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_bpd, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_fp, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->stream3, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_v, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
    const char* nv = getenv("HANK_NO_OVERLAP"); c->no_overlap = nv && nv[0] == '1';
  }
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  const size_t PG = (size_t)c->P * c->Gp;
  RC(dalloc(c, &c->d_grid, n_a));
  CK(cudaMemcpy(c->d_grid, grid, n_a * sizeof(double), cudaMemcpyHostToDevice));
  RC(dalloc(c, &c->d_Pi, (size_t)n_e * n_e));
  CK(cudaMemcpy(c->d_Pi, c->h_Pi.data(), (size_t)n_e * n_e * sizeof(double), cudaMemcpyHostToDevice));
  RC(dalloc(c, &c->d_valueT, c->Gp)); RC(dalloc(c, &c->d_D0, c->Gp));
  CK(cudaMemset(c->d_valueT, 0, c->Gp * sizeof(double))); CK(cudaMemset(c->d_D0, 0, c->Gp * sizeof(double)));
  RC(dalloc(c, &c->d_r, c->P)); RC(dalloc(c, &c->d_w, c->P));
  Tape& tp = c->tape;
  const size_t bwb = (size_t)c->P * n_e * bw_chunk(c), fwb = (size_t)c->P * n_e * fw_chunk(c);
  RC(dalloc(c, &tp.pol, PG)); RC(dalloc(c, &tp.bw, bwb)); RC(dalloc(c, &tp.rho, c->P)); RC(dalloc(c, &tp.fw, fwb));
  RC(dalloc(c, &tp.mbr, PG)); RC(dalloc(c, &tp.value_first, c->Gp));
  CK(cudaMemset(tp.pol, 0, PG * sizeof(double))); CK(cudaMemset(tp.bw, 0, bwb)); CK(cudaMemset(tp.fw, 0, fwb));
  CK(cudaMemset(tp.mbr, 0, PG * sizeof(int))); CK(cudaMemset(tp.value_first, 0, c->Gp * sizeof(double)));
  RC(dalloc(c, &c->d_kdpart, (size_t)c->P * 16 * kMaxNE)); RC(dalloc(c, &c->d_KD, c->P));
  RC(dalloc(c, &c->d_xch, (size_t)2 * c->Gp));
  RC(dalloc(c, &c->d_zero, (size_t)kZeroBytes / sizeof(double)));
  CK(cudaMemsetAsync(c->d_zero, 0, kZeroBytes, c->stream));
  RC(dalloc(c, &c->d_status, 4));
  RC(dalloc(c, &c->d_bpflag, 16));
  CK(cudaMemset(c->d_bpflag, 0, 16 * sizeof(int)));
  { const char* nt = getenv("HANK_NO_PIPE"); c->no_pipe = nt && nt[0] == '1'; }
  CK(cudaMemset(c->d_status, 0, 4 * sizeof(int)));
  CK(cudaMallocHost((void**)&c->h_status, 4 * sizeof(int)));
  const size_t n = (size_t)4 * c->P;
  RC(dalloc(c, &c->d_x, n)); RC(dalloc(c, &c->d_Z, c->P)); RC(dalloc(c, &c->d_F, n));
  return HANK_OK;
}

void hank_ctx_destroy(hank_ctx* c) {
  if (!c) return;
  if (c->stream) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
  if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
  if (c->stream3) { cudaStreamSynchronize(c->stream3); cudaStreamDestroy(c->stream3); }
  if (c->ev_v) cudaEventDestroy(c->ev_v);
  if (c->ev_x) cudaEventDestroy(c->ev_x);
  if (c->ev_bp) cudaEventDestroy(c->ev_bp);
  if (c->ev_bps) cudaEventDestroy(c->ev_bps);
  if (c->ev_bpd) cudaEventDestroy(c->ev_bpd);
  dfree(c->d_bpflag);
  if (c->ev_fp) cudaEventDestroy(c->ev_fp);
  hank_comm_destroy(c);
  newton_release(c);
  eq_release(c);
  Tape& tp = c->tape;
  dfree(c->d_grid); dfree(c->d_valueT); dfree(c->d_D0); dfree(c->d_Pi); dfree(c->d_scatter); dfree(c->d_r); dfree(c->d_w);
  dfree(tp.pol); dfree(tp.bw); dfree(tp.rho); dfree(tp.fw); dfree(tp.mbr);
  dfree(tp.value_first);
  dfree(c->d_dr); dfree(c->d_dw); dfree(c->d_dpol); dfree(c->d_dvalT); dfree(c->d_dvalue_first);
  dfree(c->d_jac_idx); dfree(c->d_thi); dfree(c->d_zero); dfree(c->d_xch); dfree(c->d_kdpart); dfree(c->d_KD); dfree(c->d_dkdpart); dfree(c->d_dKD); dfree(c->d_status);
  dfree(c->d_x); dfree(c->d_Z); dfree(c->d_F); dfree(c->d_V); dfree(c->d_JV);
  dfree(c->d_Jinv); dfree(c->d_newton); dfree(c->d_newton_i); dfree(c->d_lu_work);
  dfree(c->tape_rs_bw); dfree(c->tape_rs_fw); dfree(c->d_gather);
  for (auto& r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* hank_last_error(hank_ctx* c) { return c ? c->err.c_str() : "null context"; }

int hank_sync(hank_ctx* c) {
  if (!c) return HANK_ERR_ARG;
  CK(cudaSetDevice(c->device));
  return check_status(c);
}
int hank_timer_start(hank_ctx* c) { CK(cudaEventRecord(c->ev0, c->stream)); return HANK_OK; }
int hank_timer_stop(hank_ctx* c, float* ms) {
  RC(join_side(c));
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return HANK_OK;
}
int64_t hank_launch_count(hank_ctx* c) { return c ? c->launches : 0; }
int hank_profile(hank_ctx* c, int enable) {
  if (!c) return HANK_ERR_ARG;
  c->profile = enable != 0;
  return HANK_OK;
}
int hank_kernel_times(hank_ctx* c, double* ms4, int64_t* count4, int reset) {
  if (!c) return HANK_ERR_ARG;
  CK(cudaSetDevice(c->device));
  RC(join_side(c));
  CK(cudaStreamSynchronize(c->stream));
  for (auto& r : c->recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->kern_ms[r.kind] += ms; c->kern_n[r.kind]++; }
    c->ev_pool.push_back(r.a); c->ev_pool.push_back(r.b);
  }
  c->recs.clear();
  for (int k = 0; k < 4; ++k) {
    if (ms4) ms4[k] = c->kern_ms[k];
    if (count4) count4[k] = c->kern_n[k];
    if (reset) { c->kern_ms[k] = 0; c->kern_n[k] = 0; }
  }
  return HANK_OK;
}
int hank_reserve_lanes(hank_ctx* c, int K) {
  CK(cudaSetDevice(c->device));
  RC(ensure_lanes(c, K));
  return HANK_OK;
}

int hank_set_terminal(hank_ctx* c, const double* v) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));   // a side-stream forward sweep / residual of the last linearisation may still be in flight
  RC(copy_in(c, c->d_valueT, v, c->ne_user));
  RC(pad_columns(c, c->d_valueT));
  CK(cudaStreamSynchronize(c->stream));
  c->have_terminal = true; c->have_backward = false; c->linearized = false;
  return HANK_OK;
}
int hank_set_initial_dist(hank_ctx* c, const double* D0) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));   // the side-stream forward primal sweep reads d_D0
  RC(copy_in(c, c->d_D0, D0, c->ne_user));   // (padding columns stay zero: no mass)
  CK(cudaStreamSynchronize(c->stream));
  c->have_D0 = true; c->have_forward = false; c->linearized = false;
  return HANK_OK;
}

// ---- sweeps ------------------------------------------------------------------------------
static int backward_dev(hank_ctx* c, const double* r, const double* w, int K, const double* dr, const double* dw) {
  RC(join_side(c));
  if (!c->have_terminal) return set_error(c, HANK_ERR_STATE, "hank_set_terminal has not been called");
  if (K < 0 || (K > 0 && (!dr || !dw))) return set_error(c, HANK_ERR_ARG, "K > 0 needs dr and dw");
  RC(sw_backward_primal(c, c->P, c->d_valueT, r, w));
  c->have_backward = true; c->have_forward = false; c->K_last = 0;
  if (K > 0) {
    RC(ensure_lanes(c, K));
    if (K > c->Kcap) return set_error(c, HANK_ERR_ARG, "K exceeds the lanes that fit in device memory; chunk the pass");
    RC(sw_backward_tangent(c, c->P, K, dr, dw, nullptr, c->d_dpol, nullptr));
    c->K_last = K;
  }
  return HANK_OK;
}
static int forward_dev(hank_ctx* c, const double* pol, int K, const double* dpol, double* KD, double* dKD) {
  RC(join_side(c));  // a forward primal sweep still in flight on the side stream writes the same tape
  if (!c->have_D0) return set_error(c, HANK_ERR_STATE, "hank_set_initial_dist has not been called");
  RC(sw_forward_primal(c, c->P, c->d_D0, pol, KD));
  if (c->fp_cluster) {  // the cluster kernel leaves per-column, per-warp partials of <p_t, D_t>
    Shape sh; pick_shape(c->n_a, &sh);
    k_reduce_partials<<<nblk(c->P), 256, 0, c->stream>>>(c->d_kdpart, c->n_e * (sh.NT / 32), c->P, KD);
    c->launches++;
    CK(cudaGetLastError());
  }
  c->have_forward = true;
  if (K > 0) {
    int nw = 16;
    RC(sw_forward_tangent(c, c->P, K, dpol, c->d_dkdpart, &nw));
    k_reduce_partials<<<nblk((size_t)K * c->P), 256, 0, c->stream>>>(c->d_dkdpart, nw, K * c->P, dKD);
    c->launches++;
    CK(cudaGetLastError());
  }
  return HANK_OK;
}

int hank_block_dev(hank_ctx* c, const double* r, const double* w, int K, const double* dr, const double* dw,
                   double* KD, double* dKD) {
  CK(cudaSetDevice(c->device));
  RC(backward_dev(c, r, w, K, dr, dw));
  return forward_dev(c, c->tape.pol, K, c->d_dpol, KD, dKD);
}

int hank_backward(hank_ctx* c, const double* r, const double* w, int K, const double* dr, const double* dw) {
  CK(cudaSetDevice(c->device));
  const int P = c->P;
  if (K > 0) RC(ensure_lanes(c, K));
  if (K > c->Kcap && K > 0) return set_error(c, HANK_ERR_ARG, "K exceeds the lanes that fit in device memory");
  CK(cudaMemcpyAsync(c->d_r, r, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_w, w, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (K > 0) {
    CK(cudaMemcpyAsync(c->d_dr, dr, (size_t)K * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_dw, dw, (size_t)K * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  RC(backward_dev(c, c->d_r, c->d_w, K, c->d_dr, c->d_dw));
  c->linearized = false;
  return check_status(c);
}

int hank_forward(hank_ctx* c, double* KD, double* dKD) {
  CK(cudaSetDevice(c->device));
  if (!c->have_backward) return set_error(c, HANK_ERR_STATE, "hank_forward needs a preceding hank_backward");
  const int K = c->K_last, P = c->P;
  RC(forward_dev(c, c->tape.pol, K, c->d_dpol, c->d_KD, c->d_dKD));
  int rc = check_status(c);
  if (rc == HANK_ERR_NONMONOTONE && K == 0) {
    // a policy that is not monotone in a: redo the sweep with the scatter (atomic) lottery
    if (!c->d_scatter) RC(dalloc(c, &c->d_scatter, (size_t)3 * c->Gp));
    k_forward_scatter<<<1, 1024, 0, c->stream>>>(c->n_a, c->n_e, c->lda, P, c->d_grid, c->d_Pi, c->d_D0, c->tape.pol,
                                                 (int)fw_chunk(c), c->tape.fw, c->tape.mbr, c->d_scatter, c->d_KD);
    c->launches++;
    CK(cudaGetLastError());
    c->have_forward = true;
    rc = check_status(c);
  }
  if (rc) return rc;
  CK(cudaMemcpyAsync(KD, c->d_KD, P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (K > 0 && dKD) CK(cudaMemcpyAsync(dKD, c->d_dKD, (size_t)K * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return check_status(c);
}

int hank_forward_policies(hank_ctx* c, const double* policy, int K, const double* dpolicy, double* KD, double* dKD) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));
  const int P = c->P;
  if (K > 0) { RC(ensure_lanes(c, K)); if (K > c->Kcap) return set_error(c, HANK_ERR_ARG, "K exceeds device memory"); }
  RC(copy_in_grids(c, c->tape.pol, policy, (size_t)P));
  RC(pad_columns(c, c->tape.pol, (size_t)P));   // any monotone policy does for the zero-mass padding states
  const size_t Kpf = K > 0 ? (size_t)lane_stride(c, K) : 0;
  c->Kp_last = (int)Kpf; c->dpol_rs = false;
  for (int l = 0; l < K; ++l) {  // caller [K][P][n_e][n_a] -> device [P][n_e][Kp][lda]
    if (c->ne_user == c->n_e)
      CK(cudaMemcpy2DAsync(c->d_dpol + (size_t)l * c->lda, Kpf * c->lda * 8, dpolicy + (size_t)l * P * c->G,
                           (size_t)c->n_a * 8, (size_t)c->n_a * 8, (size_t)P * c->n_e, cudaMemcpyDefault, c->stream));
    else
      for (int e = 0; e < c->ne_user; ++e)   // one copy per income state: rows are the periods
        CK(cudaMemcpy2DAsync(c->d_dpol + ((size_t)e * Kpf + l) * c->lda, (size_t)c->n_e * Kpf * c->lda * 8,
                             dpolicy + (size_t)l * P * c->G + (size_t)e * c->n_a, (size_t)c->G * 8, (size_t)c->n_a * 8, (size_t)P,
                             cudaMemcpyDefault, c->stream));
  }
  c->have_backward = false; c->linearized = false; c->K_last = K;
  RC(forward_dev(c, c->tape.pol, K, c->d_dpol, c->d_KD, c->d_dKD));
  int rc = check_status(c);
  if (rc == HANK_ERR_NONMONOTONE && K == 0) {
    // a policy that is not monotone in a: redo the sweep with the scatter (atomic) lottery
    if (!c->d_scatter) RC(dalloc(c, &c->d_scatter, (size_t)3 * c->Gp));
    k_forward_scatter<<<1, 1024, 0, c->stream>>>(c->n_a, c->n_e, c->lda, P, c->d_grid, c->d_Pi, c->d_D0, c->tape.pol,
                                                 (int)fw_chunk(c), c->tape.fw, c->tape.mbr, c->d_scatter, c->d_KD);
    c->launches++;
    CK(cudaGetLastError());
    c->have_forward = true;
    rc = check_status(c);
  }
  if (rc) return rc;
  CK(cudaMemcpyAsync(KD, c->d_KD, P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (K > 0 && dKD) CK(cudaMemcpyAsync(dKD, c->d_dKD, (size_t)K * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return check_status(c);
}

int hank_block(hank_ctx* c, const double* r, const double* w, int K, const double* dr, const double* dw,
               double* KD, double* dKD) {
  RC(hank_backward(c, r, w, K, dr, dw));
  return hank_forward(c, KD, dKD);
}

int hank_egm_step(hank_ctx* c, const double* value_next, const double* dvalue_next, double r, double w, int K,
                  const double* dr, const double* dw, double* value, double* policy, double* dvalue, double* dpolicy) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));
  if (K < 0 || (K > 0 && (!dr || !dw))) return set_error(c, HANK_ERR_ARG, "K > 0 needs dr and dw");
  if (K > 0) RC(ensure_lanes(c, K));
  if (K > c->Kcap && K > 0) return set_error(c, HANK_ERR_ARG, "K exceeds device memory");
  // scratch for the incoming value and its lanes
  const int Gp = c->Gp;
  RC(ensure_egm(c, K));
  double* d_vn = c->d_dvalT + (size_t)K * Gp;
  RC(copy_in(c, d_vn, value_next, c->ne_user));
  RC(pad_columns(c, d_vn));
  CK(cudaMemcpyAsync(c->d_r, &r, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_w, &w, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // r, w are stack temporaries
  c->have_backward = false; c->have_forward = false; c->linearized = false; c->K_last = 0;
  RC(sw_backward_primal(c, 1, d_vn, c->d_r, c->d_w));
  if (K > 0) {
    if (dvalue_next) RC(copy_in_grids(c, c->d_dvalT, dvalue_next, (size_t)K));
    CK(cudaMemcpyAsync(c->d_dr, dr, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_dw, dw, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RC(sw_backward_tangent(c, 1, K, c->d_dr, c->d_dw, dvalue_next ? c->d_dvalT : nullptr, c->d_dpol, c->d_dvalue_first));
  }
  RC(copy_out(c, value, c->tape.value_first, c->ne_user));
  RC(copy_out(c, policy, c->tape.pol, c->ne_user));
  if (K > 0) {
    RC(copy_out_grids(c, dvalue, (const double*)c->d_dvalue_first, (size_t)K));
    for (int l = 0; l < K; ++l)  // device [1][n_e][K][lda] -> caller [K][n_e][n_a]
      CK(cudaMemcpy2DAsync(dpolicy + (size_t)l * c->G, (size_t)c->n_a * 8, c->d_dpol + (size_t)l * c->lda,
                           (size_t)c->Kp_last * c->lda * 8, (size_t)c->n_a * 8, c->ne_user, cudaMemcpyDefault, c->stream));
  }
  return check_status(c);
}

// Inner VFI of get_xVals (SteadyState.jl:132-141): iterate value_fn from a matrix of ones until
// max|ΔValue| < eps on the primal values, carrying K tangent lanes, entirely on the device (one
// 8-byte read-back per step for the stopping rule).
__global__ void k_fill_d(double* a, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_max_abs_diff(const double* __restrict__ a, const double* __restrict__ b, int n_e, int lda, int n_a,
                               double* out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n_e * lda; i += blockDim.x)
    if (i % lda < n_a) m = fmax(m, fabs(a[i] - b[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    *out = m;
  }
}

int hank_vfi(hank_ctx* c, double r, double w, int K, const double* dr, const double* dw, double eps, int max_iter,
             double* value, double* policy, double* dvalue, double* dpolicy, int* iters) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));
  if (K < 0 || (K > 0 && (!dr || !dw))) return set_error(c, HANK_ERR_ARG, "K > 0 needs dr and dw");
  if (K > 0) RC(ensure_lanes(c, K));
  if (K > c->Kcap && K > 0) return set_error(c, HANK_ERR_ARG, "K exceeds device memory");
  const int Gp = c->Gp;
  RC(ensure_egm(c, K));
  double* d_v = c->d_dvalT + (size_t)K * Gp;  // current value; lanes at d_dvalT[0..K)
  double* d_tol = c->d_KD;                     // 8-byte scratch
  CK(cudaMemcpyAsync(c->d_r, &r, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_w, &w, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (K > 0) {
    CK(cudaMemcpyAsync(c->d_dr, dr, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_dw, dw, K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->d_dvalT, 0, (size_t)K * Gp * sizeof(double), c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  k_fill_d<<<nblk(Gp), 256, 0, c->stream>>>(d_v, (size_t)Gp, 1.0);
  c->launches++;
  c->have_backward = false; c->have_forward = false; c->linearized = false; c->K_last = K;
  auto step = [&]() -> int {
    RC(sw_backward_primal(c, 1, d_v, c->d_r, c->d_w));
    if (K > 0) RC(sw_backward_tangent(c, 1, K, c->d_dr, c->d_dw, c->d_dvalT, c->d_dpol, c->d_dvalue_first));
    return HANK_OK;
  };
  RC(step());
  int it = 0;
  for (; it < max_iter; ++it) {
    k_max_abs_diff<<<1, 1024, 0, c->stream>>>(c->tape.value_first, d_v, c->ne_user, c->lda, c->n_a, d_tol);
    c->launches++;
    CK(cudaMemcpyAsync(d_v, c->tape.value_first, (size_t)Gp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (K > 0) CK(cudaMemcpyAsync(c->d_dvalT, c->d_dvalue_first, (size_t)K * Gp * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    double tol = 0.0;
    CK(cudaMemcpyAsync(&tol, d_tol, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RC(check_status(c));
    if (!(tol >= eps)) break;  // tol < eps, or NaN: stop like the reference's comparison would not
    RC(step());
  }
  if (iters) *iters = it;
  RC(copy_out(c, value, (const double*)c->tape.value_first, c->ne_user));
  RC(copy_out(c, policy, (const double*)c->tape.pol, c->ne_user));
  if (K > 0) {
    RC(copy_out_grids(c, dvalue, (const double*)c->d_dvalue_first, (size_t)K));
    for (int l = 0; l < K; ++l)
      CK(cudaMemcpy2DAsync(dpolicy + (size_t)l * c->G, (size_t)c->n_a * 8, c->d_dpol + (size_t)l * c->lda,
                           (size_t)c->Kp_last * c->lda * 8, (size_t)c->n_a * 8, c->ne_user, cudaMemcpyDefault, c->stream));
  }
  return check_status(c);
}

// ---- accessors ---------------------------------------------------------------------------
int hank_get_policy(hank_ctx* c, int t, int lane, double* out) {
  CK(cudaSetDevice(c->device));
  if (t < 1 || t > c->P || lane < 0 || lane > c->K_last) return set_error(c, HANK_ERR_ARG, "t or lane out of range");
  if (lane == 0) RC(copy_out(c, out, (const double*)(c->tape.pol + (size_t)(t - 1) * c->Gp), c->ne_user));
  else  // tangents are [t][e][K][lda]
  {
    const size_t Kp = (size_t)c->Kp_last;
    if (c->dpol_rs) {
      // written by the row-split backward sweep: [t][cluster][rank][e][l][NT]
      const int L = c->dpol_rs_L, NC = c->dpol_rs_NC, NT = c->lda / NC, cl = (lane - 1) / L, l = (lane - 1) % L;
      const double* base = c->d_dpol + (((size_t)(t - 1) * c->dpol_rs_ncl + cl) * NC) * ((size_t)c->n_e * L * NT);
      const size_t full = (size_t)c->n_a / NT, rem = (size_t)c->n_a % NT;
      for (int e = 0; e < c->ne_user; ++e) {
        const double* src = base + ((size_t)e * L + l) * NT;
        if (full) CK(cudaMemcpy2DAsync(out + (size_t)e * c->n_a, (size_t)NT * 8, src, (size_t)c->n_e * L * NT * 8, (size_t)NT * 8, full,
                                       cudaMemcpyDefault, c->stream));
        if (rem) CK(cudaMemcpyAsync(out + (size_t)e * c->n_a + full * NT, src + full * ((size_t)c->n_e * L * NT), rem * 8,
                                    cudaMemcpyDefault, c->stream));
      }
    } else
    CK(cudaMemcpy2DAsync(out, (size_t)c->n_a * 8, c->d_dpol + ((size_t)(t - 1) * c->n_e * Kp + (lane - 1)) * c->lda,
                         Kp * c->lda * 8, (size_t)c->n_a * 8, c->ne_user, cudaMemcpyDefault, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return HANK_OK;
}
int hank_get_dist(hank_ctx* c, int t, double* out) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));
  if (!c->have_forward) return set_error(c, HANK_ERR_STATE, "no forward sweep has been run");
  if (t < 1 || t > c->P) return set_error(c, HANK_ERR_ARG, "t out of range");
  // D_t sits in field FW_D of each column chunk: [t][e][FW_NF][lda]
  CK(cudaMemcpy2DAsync(out, (size_t)c->n_a * 8, c->tape.fw + (size_t)(t - 1) * c->n_e * fw_chunk(c) + (size_t)FW_D * c->lda * 8,
                       fw_chunk(c), (size_t)c->n_a * 8, c->ne_user, cudaMemcpyDefault, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return HANK_OK;
}
int hank_get_value_first(hank_ctx* c, int lane, double* out) {
  CK(cudaSetDevice(c->device));
  if (lane != 0) return set_error(c, HANK_ERR_ARG, "only the primal value is retained by the sweeps");
  RC(copy_out(c, out, (const double*)c->tape.value_first, c->ne_user));
  CK(cudaStreamSynchronize(c->stream));
  return HANK_OK;
}
int hank_get_brackets(hank_ctx* c, int t, int32_t* m) {
  CK(cudaSetDevice(c->device));
  RC(join_side(c));
  if (!c->have_forward) return set_error(c, HANK_ERR_STATE, "no forward sweep has been run");
  if (t < 1 || t > c->P) return set_error(c, HANK_ERR_ARG, "t out of range");
  RC(copy_out(c, m, (const int32_t*)(c->tape.mbr + (size_t)(t - 1) * c->Gp), c->ne_user));
  CK(cudaStreamSynchronize(c->stream));
  return HANK_OK;
}
int hank_lottery(hank_ctx* c, const double* policy, int32_t* m, double* omega) {
  CK(cudaSetDevice(c->device));
  const int G = c->G;
  double* d_p = nullptr; int32_t* d_m = nullptr; double* d_o = nullptr;
  auto run = [&]() -> int {
    RC(dalloc(c, &d_p, G)); RC(dalloc(c, &d_m, G)); RC(dalloc(c, &d_o, G));
    CK(cudaMemcpyAsync(d_p, policy, G * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_lottery<<<nblk(G), 256, 0, c->stream>>>(c->d_grid, c->n_a, G, d_p, d_m, d_o);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(m, d_m, G * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (omega) CK(cudaMemcpyAsync(omega, d_o, G * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return HANK_OK;
  };
  const int rc = run();
  dfree(d_p); dfree(d_m); dfree(d_o);   // also on the error paths
  return rc;
}

// ---- Krusell-Smith F and JVPs ---------------------------------------------------------------
int hank_ks_configure(hank_ctx* c, double alpha, double delta, double ss_start_KS) {
  if (c->eq_on) return set_error(c, HANK_ERR_STATE, "the equations of this context were set by hank_eq_configure");
  c->alpha = alpha; c->delta = delta; c->ssKS = ss_start_KS; c->ks_ready = true; c->linearized = false;
  return HANK_OK;
}

int hank_ks_linearize_dev(hank_ctx* c, const double* x, const double* Z, double* F) {
  CK(cudaSetDevice(c->device));
  if (!c->ks_ready) return set_error(c, HANK_ERR_STATE, "hank_ks_configure has not been called");
  if (!c->have_terminal || !c->have_D0) return set_error(c, HANK_ERR_STATE, "terminal value / initial distribution not set");
  const int P = c->P; const size_t n = (size_t)c->n_endog * P;
  // the previous linearisation's forward sweep and residuals (side stream) read d_x, d_Z and d_KD
  RC(join_side(c));
  if (x != c->d_x) CK(cudaMemcpyAsync(c->d_x, x, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (Z != c->d_Z) CK(cudaMemcpyAsync(c->d_Z, Z, (size_t)c->n_exog * P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (c->eq_on) RC(eq_extract_rw(c, c->d_x, c->d_r, c->d_w));
  else { k_extract_rw<<<nblk(P), 256, 0, c->stream>>>(c->d_x, P, c->d_r, c->d_w); c->launches++; }
  // The forward primal sweep and the residuals only feed the forward tangent / F: they go to the
  // high-priority side stream so that a following backward tangent sweep overlaps them.  Pipelined linearisation
  // (hank_ctx.h): the backward primal sweep goes there too, and a backward tangent sweep of the one-CTA ring kernel
  // launched next may run alongside it, following its progress counters; everything else joins first (join_bp).
  cudaStream_t main_stream = c->stream;
  const bool pipe = !c->no_overlap && !c->no_pipe && c->stream2 && c->d_bpflag && c->pipe_hint > 0;
  bool fused = false;
  if (pipe) {
    k_fill_rho<<<nblk(P), 256, 0, main_stream>>>(c->d_r, P, c->tape.rho);
    c->launches++;
    CK(cudaEventRecord(c->ev_bp, main_stream));
    CK(cudaStreamWaitEvent(c->stream2, c->ev_bp, 0));
    c->stream = c->stream2;
    int rcb = -1;
    // a multi-wave tangent pass followed the last linearisation (pipe_hint 2): both primal sweeps in one launch, so that
    // the cluster's SMs are not taken by that pass's pending CTAs between the two sweeps
    if (c->pipe_hint == 2) {
      rcb = sw_primal_both(c, P, c->d_valueT, c->d_r, c->d_w, c->d_D0);
      if (rcb == 0) {
        fused = true;
        c->have_backward = true; c->K_last = 0;
        k_reduce_partials<<<nblk(P), 256, 0, c->stream>>>(c->d_kdpart, c->n_e * (c->lda / 32), P, c->d_KD);
        c->launches++;
        rcb = cuda_check(c, cudaGetLastError(), "k_reduce_partials");
        c->have_forward = true;
      } else if (rcb > 0) { c->stream = main_stream; return rcb; }
    }
    if (!fused) {
      c->bp_pipe_req = true;
      rcb = backward_dev(c, c->d_r, c->d_w, 0, nullptr, nullptr);
      c->bp_pipe_req = false;
    }
    if (rcb == HANK_OK) rcb = cuda_check(c, cudaEventRecord(c->ev_bpd, c->stream2), "cudaEventRecord(ev_bpd)");
    if (rcb != HANK_OK) { c->stream = main_stream; cudaStreamWaitEvent(main_stream, c->ev_bpd, 0); return rcb; }
  } else {
    RC(backward_dev(c, c->d_r, c->d_w, 0, nullptr, nullptr));
    if (!c->no_overlap) {
      CK(cudaEventRecord(c->ev_bp, main_stream));
      CK(cudaStreamWaitEvent(c->stream2, c->ev_bp, 0));
      c->stream = c->stream2;
    }
  }
  int rc = fused ? HANK_OK : forward_dev(c, c->tape.pol, 0, nullptr, c->d_KD, nullptr);
  if (rc == HANK_OK) {
    if (c->eq_on) rc = eq_residual(c, c->d_x, c->d_KD, c->d_Z, F);
    else {
      k_ks_residual<<<nblk(P), 256, 0, c->stream>>>(P, c->alpha, c->delta, c->ssKS, c->d_x, c->d_KD, c->d_Z, F);
      c->launches++;
      rc = cuda_check(c, cudaGetLastError(), "k_ks_residual");
    }
  }
  if (!c->no_overlap) {
    c->stream = main_stream;
    c->bp_pending = pipe;   // (set here: the forward sweep above ran on the primal's own stream)
    if (rc == HANK_OK) { CK(cudaEventRecord(c->ev_fp, c->stream2)); c->fp_pending = true; }
  }
  RC(rc);
  c->linearized = true;
  return HANK_OK;
}

int hank_ks_jvp_dev(hank_ctx* c, int K, const double* V, double* JV) {
  CK(cudaSetDevice(c->device));
  if (!c->linearized) return set_error(c, HANK_ERR_STATE, "hank_ks_jvp needs a preceding hank_ks_linearize");
  if (K < 1) return set_error(c, HANK_ERR_ARG, "K must be >= 1");
  const int P = c->P; const size_t n = (size_t)c->n_endog * P;
  RC(ensure_lanes(c, K));
  for (int k0 = 0; k0 < K; k0 += c->Kcap) {
    const int kc = std::min(c->Kcap, K - k0);
    const double* Vc = V + (size_t)k0 * n;
    if (c->eq_on) RC(eq_extract_drdw(c, kc, Vc, c->d_dr, c->d_dw));
    else { k_extract_drdw<<<nblk((size_t)kc * P), 256, 0, c->stream>>>(Vc, P, kc, c->d_dr, c->d_dw); c->launches++; }
    RC(tangent_pass(c, P, kc));
    c->K_last = kc;
    if (c->eq_on) RC(eq_residual_tangent(c, kc, c->d_x, c->d_KD, c->d_Z, Vc, nullptr, c->d_dKD, nullptr, JV + (size_t)k0 * n));
    else {
      k_ks_residual_tangent<<<nblk((size_t)kc * P), 256, 0, c->stream>>>(P, kc, c->alpha, c->ssKS, c->d_x, c->d_Z, Vc,
                                                                        c->d_dKD, JV + (size_t)k0 * n);
      c->launches++;
      CK(cudaGetLastError());
    }
  }
  return HANK_OK;
}

int hank_ks_linearize(hank_ctx* c, const double* x, const double* Z, double* F) {
  CK(cudaSetDevice(c->device));
  const int P = c->P; const size_t n = (size_t)c->n_endog * P;
  RC(join_side(c));   // before d_x / d_Z are overwritten
  CK(cudaMemcpyAsync(c->d_x, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_Z, Z, (size_t)c->n_exog * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  RC(hank_ks_linearize_dev(c, c->d_x, c->d_Z, c->d_F));
  RC(join_side(c));
  if (F) CK(cudaMemcpyAsync(F, c->d_F, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  int rc = check_status(c);
  if (rc) c->linearized = false;
  return rc;
}

int hank_ks_jvp(hank_ctx* c, int K, const double* V, double* JV) {
  CK(cudaSetDevice(c->device));
  if (K < 1) return set_error(c, HANK_ERR_ARG, "K must be >= 1");
  const size_t n = (size_t)c->n_endog * c->P;
  RC(ensure_V(c, K));
  CK(cudaMemcpyAsync(c->d_V, V, n * K * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  RC(hank_ks_jvp_dev(c, K, c->d_V, c->d_JV));
  CK(cudaMemcpyAsync(JV, c->d_JV, n * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return check_status(c);
}

// fullFunction(x) and JVP(fullFunction, x, V[:,k]) in one call with host buffers: the upload of the K
// tangent seeds runs on a copy stream underneath the primal backward sweep.
int hank_ks_fjvp(hank_ctx* c, const double* x, const double* Z, int K, const double* V, double* F, double* JV) {
  CK(cudaSetDevice(c->device));
  if (K < 1) return set_error(c, HANK_ERR_ARG, "K must be >= 1");
  const int P = c->P; const size_t n = (size_t)c->n_endog * P;
  RC(ensure_V(c, K));
  RC(ensure_lanes(c, K));
  RC(join_side(c));
  CK(cudaStreamSynchronize(c->stream));   // d_V may still feed the previous call's kernels
  CK(cudaMemcpyAsync(c->d_V, V, n * K * sizeof(double), cudaMemcpyHostToDevice, c->stream3));
  CK(cudaEventRecord(c->ev_v, c->stream3));
  CK(cudaMemcpyAsync(c->d_x, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_Z, Z, (size_t)c->n_exog * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  RC(hank_ks_linearize_dev(c, c->d_x, c->d_Z, c->d_F));
  CK(cudaStreamWaitEvent(c->stream, c->ev_v, 0));
  // A multi-wave pass is cut at CTA-wave boundaries (the first wave leaves n_e SMs to the overlapped forward
  // primal): the columns of a finished wave travel to the host on the copy stream under the next wave.
  const int L = std::max(1, sw_lanes_per_cta(c, K));
  const int first = L * (c->sm_count - (c->no_overlap ? 0 : c->n_e)), wave = L * c->sm_count;
  if (L == 4 && K > first && first > 0) {
    for (int k0 = 0, kb = first; k0 < K; k0 += kb, kb = wave) {
      const int kc = std::min(kb, K - k0);
      RC(hank_ks_jvp_dev(c, kc, c->d_V + (size_t)k0 * n, c->d_JV + (size_t)k0 * n));
      if (k0 + kc < K) {
        CK(cudaEventRecord(c->ev_v, c->stream));
        CK(cudaStreamWaitEvent(c->stream3, c->ev_v, 0));
        CK(cudaMemcpyAsync(JV + (size_t)k0 * n, c->d_JV + (size_t)k0 * n, n * kc * sizeof(double), cudaMemcpyDeviceToHost,
                           c->stream3));
      } else {
        CK(cudaMemcpyAsync(JV + (size_t)k0 * n, c->d_JV + (size_t)k0 * n, n * kc * sizeof(double), cudaMemcpyDeviceToHost,
                           c->stream));
      }
    }
  } else {
    RC(hank_ks_jvp_dev(c, K, c->d_V, c->d_JV));
    CK(cudaMemcpyAsync(JV, c->d_JV, n * K * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (F) CK(cudaMemcpyAsync(F, c->d_F, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream3));
  int rc = check_status(c);
  if (rc) c->linearized = false;
  return rc;
}

// Jacobian columns `cols` (0-based, ascending) at the linearisation point, written to J in that order.
static int jacobian_cols_impl(hank_ctx* c, const std::vector<int>& cols, double* J) {
  const int P = c->P, ne = c->n_endog, n = ne * P;
  const int ncols = (int)cols.size();
  std::vector<int> lane_col, lane_pos, col_lane(ncols, -1), chunk_cols, thi;
  for (int j = 0; j < ncols; ++j)
    if (cols[j] % ne == c->eq_ir || cols[j] % ne == c->eq_iw) { lane_col.push_back(cols[j]); lane_pos.push_back(j); }
  int Kh = (int)lane_col.size();
  if (Kh > 0) RC(ensure_lanes(c, Kh));
  if (c->jac_idx_cap < ncols) {
    dfree(c->d_jac_idx);
    c->jac_idx_cap = 0;
    RC(dalloc(c, &c->d_jac_idx, (size_t)3 * ncols));
    c->jac_idx_cap = ncols;
  }
  int* d_lane_col = c->d_jac_idx; int* d_col_lane = c->d_jac_idx + c->jac_idx_cap; int* d_col_ids = d_col_lane + c->jac_idx_cap;
  CK(cudaMemcpyAsync(d_col_ids, cols.data(), ncols * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  const int chunk = Kh > 0 ? c->Kcap : 1;
  // Y / KS columns need no sweeps: handled by col_lane = -1. Household lanes go in chunks of
  // Kcap; each chunk writes the columns it owns.
  int done_cols = 0;  // positions [0, done_cols) of `cols` written
  for (int k0 = 0; k0 < std::max(Kh, 1); k0 += chunk) {
    const int kc = Kh > 0 ? std::min(chunk, Kh - k0) : 0;
    // positions covered by this chunk: up to (not including) the first household column of the next chunk
    const int col_hi = (k0 + kc < Kh) ? lane_pos[k0 + kc] : ncols;
    const int col_lo = done_cols;
    std::fill(col_lane.begin(), col_lane.end(), -1);
    // Lanes of a chunk run latest seed first.  A unit seed at period s leaves V̇ and ṗ exactly zero after s,
    // so each group of kThiGroup lanes carries a horizon: the backward tangent starts there, the forward
    // tangent stages zeros beyond it, and the long CTAs are scheduled ahead of the short ones.
    chunk_cols.assign(lane_col.begin() + k0, lane_col.begin() + k0 + kc);
    std::reverse(chunk_cols.begin(), chunk_cols.end());
    for (int l = 0; l < kc; ++l) col_lane[lane_pos[k0 + kc - 1 - l]] = l;
    if (kc > 0) {
      CK(cudaMemcpyAsync(d_lane_col, chunk_cols.data(), kc * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemsetAsync(c->d_dr, 0, (size_t)kc * P * sizeof(double), c->stream));
      CK(cudaMemsetAsync(c->d_dw, 0, (size_t)kc * P * sizeof(double), c->stream));
      k_unit_seeds<<<nblk(kc), 256, 0, c->stream>>>(P, kc, ne, c->eq_ir, c->eq_iw, d_lane_col, c->d_dr, c->d_dw);
      c->launches++;
      const int ngroups = (kc + kThiGroup - 1) / kThiGroup + 1;   // + one group of padding lanes
      if (!c->no_skip) {
        thi.assign(ngroups, 0);
        for (int l = 0; l < kc; ++l) thi[l / kThiGroup] = std::max(thi[l / kThiGroup], chunk_cols[l] / ne + 1);
        if (c->thi_cap < ngroups) {
          dfree(c->d_thi); c->thi_cap = 0;
          RC(dalloc(c, &c->d_thi, (size_t)ngroups + 64));
          c->thi_cap = ngroups + 64;
        }
        CK(cudaMemcpyAsync(c->d_thi, thi.data(), ngroups * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        c->pass_thi = c->d_thi;
        c->pass_Kp = (kc + kThiGroup - 1) / kThiGroup * kThiGroup;
      }
      int rc = tangent_pass(c, P, kc);
      c->pass_thi = nullptr; c->pass_Kp = 0;
      RC(rc);
      c->K_last = 0;   // policy tangents beyond the horizons were never written: not retrievable
    }
    CK(cudaMemcpyAsync(d_col_lane, col_lane.data() + col_lo, (col_hi - col_lo) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (c->eq_on)   // direct terms by a unit-seed pass through the equation programs, household term from the lanes' K̇D
      RC(eq_residual_tangent(c, col_hi - col_lo, c->d_x, c->d_KD, c->d_Z, nullptr, d_col_ids + col_lo, c->d_dKD, d_col_lane,
                             J + (size_t)col_lo * n));
    else {
      k_ks_jac_columns<<<nblk((size_t)(col_hi - col_lo) * P), 256, 0, c->stream>>>(
          P, d_col_ids + col_lo, col_hi - col_lo, c->alpha, c->ssKS, c->d_x, c->d_Z, d_col_lane, c->d_dKD, J + (size_t)col_lo * n);
      c->launches++;
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(c->stream));  // host vectors are reused by the next chunk
    done_cols = col_hi;
  }
  return HANK_OK;
}

int hank_ks_jacobian_columns_dev(hank_ctx* c, int col_begin, int col_end, double* J) {
  CK(cudaSetDevice(c->device));
  if (!c->linearized) return set_error(c, HANK_ERR_STATE, "hank_ks_jacobian_columns needs a preceding hank_ks_linearize");
  const int n = c->n_endog * c->P;
  if (col_begin < 1 || col_end > n + 1 || col_end <= col_begin) return set_error(c, HANK_ERR_ARG, "bad column range");
  std::vector<int> cols(col_end - col_begin);
  for (int j = 0; j < (int)cols.size(); ++j) cols[j] = col_begin - 1 + j;
  return jacobian_cols_impl(c, cols, J);
}

int hank_ks_jacobian_column_list_dev(hank_ctx* c, int ncols, const int* cols1, double* J) {
  CK(cudaSetDevice(c->device));
  if (!c->linearized) return set_error(c, HANK_ERR_STATE, "hank_ks_jacobian_column_list needs a preceding hank_ks_linearize");
  const int n = c->n_endog * c->P;
  if (ncols < 1 || !cols1) return set_error(c, HANK_ERR_ARG, "empty column list");
  std::vector<int> cols(ncols);
  for (int j = 0; j < ncols; ++j) {
    if (cols1[j] < 1 || cols1[j] > n || (j > 0 && cols1[j] <= cols1[j - 1]))
      return set_error(c, HANK_ERR_ARG, "column list must be ascending, 1-based and within 1..n");
    cols[j] = cols1[j] - 1;
  }
  return jacobian_cols_impl(c, cols, J);
}

int hank_ks_jacobian_column_list(hank_ctx* c, int ncols, const int* cols1, double* J) {
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_endog * c->P;
  if (ncols < 1) return set_error(c, HANK_ERR_ARG, "empty column list");
  RC(ensure_V(c, ncols));
  RC(hank_ks_jacobian_column_list_dev(c, ncols, cols1, c->d_JV));
  CK(cudaMemcpyAsync(J, c->d_JV, n * ncols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return check_status(c);
}

int hank_ks_jacobian_columns(hank_ctx* c, int col_begin, int col_end, double* J) {
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_endog * c->P;
  if (col_end <= col_begin) return set_error(c, HANK_ERR_ARG, "bad column range");
  const int ncols = col_end - col_begin;
  RC(ensure_V(c, ncols));
  RC(hank_ks_jacobian_columns_dev(c, col_begin, col_end, c->d_JV));
  CK(cudaMemcpyAsync(J, c->d_JV, n * ncols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return check_status(c);
}

}  // extern "C"
