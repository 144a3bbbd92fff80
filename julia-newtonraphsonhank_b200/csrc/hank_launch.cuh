// hank_launch.cuh — launch-configuration dispatch for the sweep kernels, templated on n_e.
// Included only by hank_ne*.cu, each of which instantiates Sweeps<NE> for one n_e so the
// instantiations compile in parallel.
#pragma once
#include <cmath>
#include <cstdio>
#include "hank_ctx.h"

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

// ---- backward primal -------------------------------------------------------------------
template <int NE, int R, int NT>
static int bp_launch(hank_ctx* c, int P, const double* valueT, const double* r, const double* w) {
  const Consts<NE> M = make_consts<NE>(c, P);
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

// ---- lanes per CTA ---------------------------------------------------------------------
// L*R <= 4 (register budget: L*R*NE doubles of state per thread), k̇ staging L*G*8 B must fit.
template <int NE>
int Sweeps<NE>::lanes_per_cta(hank_ctx* c, int K) {
  Shape s;
  if (!pick_shape(c->n_a, &s)) return 0;
  int L = 4 / s.R;
  while (L > 1 && (size_t)L * NE * c->lda * sizeof(double) > (size_t)c->smem_max) L >>= 1;
  // fill the SMs before deepening the lanes per CTA
  while (L > 1 && K < c->sm_count * L) L >>= 1;
  return L;
}

// ---- backward tangent ------------------------------------------------------------------
template <int NE, int R, int NT, int L>
static int bt_launch(hank_ctx* c, int P, int K, const double* dr, const double* dw, const double* dvalT,
                     double* dpol, double* dvf) {
  const Consts<NE> M = make_consts<NE>(c, P);
  const size_t smem = (size_t)L * NE * NT * R * sizeof(double);
  const int grid = (K + L - 1) / L;
  HANK_LAUNCH(KIND_BT, (k_backward_tangent<NE, R, NT, L>), grid, NT, smem, M, c->tape, K, dr, dw, dvalT, dpol, dvf);
}
template <int NE>
int Sweeps<NE>::backward_tangent(hank_ctx* c, int P, int K, const double* dr, const double* dw,
                                 const double* dvalT, double* dpol, double* dvf) {
  Shape s;
  if (!pick_shape(c->n_a, &s)) return set_error(c, 1, "n_a > 2048 is not supported");
  const int L = lanes_per_cta(c, K);
#define BT(R_, NT_, L_) return bt_launch<NE, R_, NT_, L_>(c, P, K, dr, dw, dvalT, dpol, dvf)
  if (s.NT == 256) { if (L == 4) BT(1, 256, 4); if (L == 2) BT(1, 256, 2); BT(1, 256, 1); }
  if (s.R == 1) { if (L == 4) BT(1, 512, 4); if (L == 2) BT(1, 512, 2); BT(1, 512, 1); }
  if (s.R == 2) { if (L == 2) BT(2, 512, 2); BT(2, 512, 1); }
  BT(4, 512, 1);
#undef BT
}

// ---- forward primal --------------------------------------------------------------------
template <int NE, int R, int NT, int CS>
static int fp_launch(hank_ctx* c, int P, const double* D0, const double* pol, double* KD) {
  const Consts<NE> M = make_consts<NE>(c, P);
  constexpr int LDA = NT * R;
  const size_t smem = ((size_t)2 * CS * LDA + LDA) * sizeof(double) + ((size_t)CS * LDA + (size_t)CS * (LDA + 4)) * sizeof(int);
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
  const bool fits = full <= (size_t)c->smem_max;
  constexpr int C4 = small_cs<NE>();
#define FP(R_, NT_) do { if (fits) return fp_launch<NE, R_, NT_, NE>(c, P, D0, pol, KD); \
                         return fp_launch<NE, R_, NT_, C4>(c, P, D0, pol, KD); } while (0)
  if (s.NT == 256) FP(1, 256);
  if (s.R == 1) FP(1, 512);
  if (s.R == 2) FP(2, 512);
  FP(4, 512);
#undef FP
}

// ---- forward tangent -------------------------------------------------------------------
template <int NE, int R, int NT, int L, int CS>
static int ft_launch(hank_ctx* c, int P, int K, const double* pol, const double* dpol, double* dkdpart) {
  const Consts<NE> M = make_consts<NE>(c, P);
  const size_t smem = (size_t)2 * L * CS * NT * R * sizeof(double);
  const int grid = (K + L - 1) / L;
  HANK_LAUNCH(KIND_FT, (k_forward_tangent<NE, R, NT, L, CS>), grid, NT, smem, M, c->tape, K, pol, dpol, nullptr, dkdpart, nullptr);
}
template <int NE>
int Sweeps<NE>::forward_tangent(hank_ctx* c, int P, int K, const double* pol, const double* dpol,
                                double* dkdpart, int* nw_out) {
  Shape s;
  if (!pick_shape(c->n_a, &s)) return set_error(c, 1, "n_a > 2048 is not supported");
  *nw_out = s.NT / 32;
  const int L = lanes_per_cta(c, K);
  const bool fits = (size_t)2 * L * NE * c->lda * 8 <= (size_t)c->smem_max;
  constexpr int C4 = small_cs<NE>();
  if (!fits && (size_t)2 * L * C4 * c->lda * 8 > (size_t)c->smem_max)
    return set_error(c, 1, "forward tangent staging does not fit in shared memory");
#define FT(R_, NT_, L_) do { if (fits) return ft_launch<NE, R_, NT_, L_, NE>(c, P, K, pol, dpol, dkdpart); \
                             return ft_launch<NE, R_, NT_, L_, C4>(c, P, K, pol, dpol, dkdpart); } while (0)
  if (s.NT == 256) { if (L == 4) FT(1, 256, 4); if (L == 2) FT(1, 256, 2); FT(1, 256, 1); }
  if (s.R == 1) { if (L == 4) FT(1, 512, 4); if (L == 2) FT(1, 512, 2); FT(1, 512, 1); }
  if (s.R == 2) { if (L == 2) FT(2, 512, 2); FT(2, 512, 1); }
  FT(4, 512, 1);
#undef FT
}

}  // namespace hank
