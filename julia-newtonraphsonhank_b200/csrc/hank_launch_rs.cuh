// hank_launch_rs.cuh — launchers of the row-split cluster tangent sweeps (hank_tangent_rowsplit.cuh), templated
// on n_e.  Included only by hank_ne*_rs.cu so that these kernels compile in their own translation units.
#pragma once
#include <cstdlib>
#include "hank_launch.cuh"
#include "hank_tangent_rowsplit.cuh"

namespace hank {

// ---- row-split cluster launchers ---------------------------------------------------------
static int rs_ring_slots(const hank_ctx* c, size_t fixed, size_t slot, int min_slots, int max_slots) {
  if (fixed + (size_t)min_slots * slot > (size_t)c->smem_max) return 0;
  const int s = (int)(((size_t)c->smem_max - fixed) / slot);
  return s > max_slots ? max_slots : s;
}
// Row-block-major copies of the tape (k_tape_rowblocks_*), made on first use per linearisation and block size NT
// (a K = 1 pass after a K = 64 pass at the same linearisation uses another cluster shape, hence another layout).
static size_t tape_rs_rng_offset(const hank_ctx* c, bool forward) {
  const size_t ncols = (size_t)c->P_alloc * c->n_e;
  return forward ? ncols * (36 * (size_t)c->lda + 16 * (size_t)(c->lda / 32)) : ncols * 52 * (size_t)c->lda;
}
// behind the ranges: destinations per source (k_rs_st_mark_*, one int per grid point and period), then values received
// per (period, CTA)
static size_t tape_rs_info_offset(const hank_ctx* c, bool forward) { return tape_rs_rng_offset(c, forward) + (size_t)c->P_alloc * 64 * sizeof(int); }
static size_t tape_rs_cnt_offset(const hank_ctx* c, bool forward) {
  return tape_rs_info_offset(c, forward) + (size_t)c->P_alloc * c->n_e * c->lda * sizeof(int);
}
static int ensure_tape_rs(hank_ctx* c, int P, int NT, bool forward) {
  const size_t ncols = (size_t)c->P_alloc * c->n_e;
  // (sized for the smallest block any shape uses, 32 rows: the forward copy carries 4 extra range starts per block)
  const size_t bwb = ncols * 52 * (size_t)c->lda, fwb = ncols * (36 * (size_t)c->lda + 16 * (size_t)(c->lda / 32));
  const size_t rngb = (size_t)c->P_alloc * 64 * sizeof(int);   // source-block ranges per (period, block) behind each copy
  const size_t infob = ncols * (size_t)c->lda * sizeof(int), cntb = rngb;
  if (!c->tape_rs_bw) {
    int rc = cuda_check(c, cudaMalloc((void**)&c->tape_rs_bw, bwb + rngb + infob + cntb), "cudaMalloc(tape_rs_bw)");
    if (rc) return rc;
    rc = cuda_check(c, cudaMalloc((void**)&c->tape_rs_fw, fwb + rngb + infob + cntb), "cudaMalloc(tape_rs_fw)");
    if (rc) return rc;
  }
  const int NCb = c->lda / NT;
  const unsigned rgrid = (unsigned)(((size_t)P * NCb * 32 + 255) / 256);
  const size_t n = (size_t)P * c->n_e * c->lda;
  const unsigned grid = (unsigned)((n + 255) / 256);
  const bool st = NT == 64 && NCb <= 8;   // the shapes of the st.async kernels
  if (!forward && c->tape_rs_bw_nt != NT) {
    k_tape_rowblocks_bw<<<grid, 256, 0, c->stream>>>(c->tape.bw, c->tape_rs_bw, P * c->n_e, c->n_e, c->lda, NT);
    k_rs_ranges<<<rgrid, 256, 0, c->stream>>>(c->tape_rs_bw, P, c->n_e, NCb, NT, c->n_a, 0, reinterpret_cast<int*>(c->tape_rs_bw + bwb));
    c->launches += 2; c->tape_rs_bw_nt = NT;
    if (st) {
      int* info = reinterpret_cast<int*>(c->tape_rs_bw + tape_rs_info_offset(c, false));
      int* cnt = reinterpret_cast<int*>(c->tape_rs_bw + tape_rs_cnt_offset(c, false));
      cudaMemsetAsync(info, 0, n * sizeof(int), c->stream);
      cudaMemsetAsync(cnt, 0, (size_t)P * NCb * sizeof(int), c->stream);
      k_rs_st_mark_bw<<<grid, 256, 0, c->stream>>>(c->tape_rs_bw, P, c->n_e, NCb, NT, c->n_a, info);
      k_rs_st_count_bw<<<grid, 256, 0, c->stream>>>(info, P, c->n_e, NCb, NT, cnt);
      c->launches += 2;
    }
  }
  if (forward && c->tape_rs_fw_nt != NT) {
    k_tape_rowblocks_fw<<<grid, 256, 0, c->stream>>>(c->tape.fw, c->tape_rs_fw, P * c->n_e, c->n_e, c->lda, NT);
    k_rs_ranges<<<rgrid, 256, 0, c->stream>>>(c->tape_rs_fw, P, c->n_e, NCb, NT, c->n_a, 1, reinterpret_cast<int*>(c->tape_rs_fw + fwb));
    c->launches += 2; c->tape_rs_fw_nt = NT;
    if (st) {
      int* info = reinterpret_cast<int*>(c->tape_rs_fw + tape_rs_info_offset(c, true));
      int* cnt = reinterpret_cast<int*>(c->tape_rs_fw + tape_rs_cnt_offset(c, true));
      cudaMemsetAsync(info, 0xFF, n * sizeof(int), c->stream);
      cudaMemsetAsync(cnt, 0, (size_t)P * NCb * sizeof(int), c->stream);
      k_rs_st_mark_fw<<<grid, 256, 0, c->stream>>>(c->tape_rs_fw, P, c->n_e, NCb, NT, c->n_a, info, cnt);
      c->launches++;
    }
  }
  return cuda_check(c, cudaGetLastError(), "k_tape_rowblocks");
}
template <int NE, int NC, int NT, int L, int GC, int LA>
static int bt_rs_launch(hank_ctx* c, int P, int K, const double* dr, const double* dw, double* dpol) {
  const Consts<NE> M = make_consts<NE>(c, P);
  const int ncl = (K + L - 1) / L;
  c->Kp_last = ncl * L;
  c->dpol_rs = true; c->dpol_rs_L = L; c->dpol_rs_NC = NC; c->dpol_rs_ncl = ncl;
  const size_t slot = (size_t)GC * rs_bw_col_bytes<NT>();
  // one lane, a whole period per exchange: the push kernels (hank_tangent_rowsplit.cuh) where their buffers fit
  if constexpr (L == 1 && GC == NE && LA == 0 && NE * NT + 64 <= 1024) {
    const bool no_push = c->no_rs_push, no_st = c->no_rs_st;
    if constexpr (NT == 64 && NC <= 8) {   // every thread sends its own value with st.async
      const int Ss = rs_ring_slots(c, rs_bw_st_smem<NE, NC, NT>(0), rs_bw_st_slot<NE, NC, NT>(), 2, 6);
      if (!no_push && !no_st && Ss >= 2) {
        int rc = ensure_tape_rs(c, P, NT, false);
        if (rc) return rc;
        rc = launch_cluster_grid(c, KIND_BT, k_backward_tangent_rs_st<NE, NC, NT>, ncl * NC, NC, NE * NT + 32, rs_bw_st_smem<NE, NC, NT>(Ss),
                                 "k_backward_tangent_rs_st", M, c->tape, (const unsigned char*)c->tape_rs_bw,
                                 reinterpret_cast<const int*>(c->tape_rs_bw + tape_rs_info_offset(c, false)),
                                 reinterpret_cast<const int*>(c->tape_rs_bw + tape_rs_cnt_offset(c, false)), (const int*)c->d_status, K, Ss, c->pass_thi, dr, dw, dpol);
        if (rc >= 0) return rc;
      }
    }
    const int Sp = rs_ring_slots(c, rs_bw_push_smem<NE, NC, NT>(0), slot + 16, 2, 6);
    if (!no_push && Sp >= 2) {
      int rc = ensure_tape_rs(c, P, NT, false);
      if (rc) return rc;
      rc = launch_cluster_grid(c, KIND_BT, k_backward_tangent_rs_push<NE, NC, NT>, ncl * NC, NC, NE * NT + 64, rs_bw_push_smem<NE, NC, NT>(Sp),
                               "k_backward_tangent_rs_push", M, c->tape, (const unsigned char*)c->tape_rs_bw,
                               reinterpret_cast<const int*>(c->tape_rs_bw + tape_rs_rng_offset(c, false)), K, Sp, c->pass_thi, dr, dw, dpol);
      if (rc >= 0) return rc;
    }
  }
  const int S = rs_ring_slots(c, rs_bw_smem<NT, L, GC, LA>(0), slot + 16, LA + 2, GC == 1 ? 3 * NE : 6);
  if (S < LA + 2) return -1;
  int rc = ensure_tape_rs(c, P, NT, false);
  if (rc) return rc;
  return launch_cluster_grid(c, KIND_BT, k_backward_tangent_rs<NE, NC, NT, L, GC, LA>, ncl * NC, NC, NT + 64,
                             rs_bw_smem<NT, L, GC, LA>(S), "k_backward_tangent_rs", M, c->tape, (const unsigned char*)c->tape_rs_bw, K, S,
                             c->pass_thi, dr, dw, dpol, c->rs_relaxed ? 0 : 1);
}
template <int NE, int NC, int NT, int L, int GC, int LA>
static int ft_rs_launch(hank_ctx* c, int P, int K, const double* dpol, double* dkdpart) {
  const Consts<NE> M = make_consts<NE>(c, P);
  const int ncl = (K + L - 1) / L;
  // ṗ comes from this pass's row-split backward sweep (same shape), or from the caller in the column-major layout
  const bool pd_rs = c->dpol_rs && c->dpol_rs_L == L && c->dpol_rs_NC == NC && c->dpol_rs_ncl == ncl;
  const int Kp = c->pass_Kp ? c->pass_Kp : (K + kThiGroup - 1) / kThiGroup * kThiGroup;   // (the caller-layout stride of hank_forward_policies)
  const size_t slot = (size_t)GC * rs_fw_col_bytes<NT, L>();
  // one lane, a whole period per exchange: a thread per (income state, row) instead of per row (hank_tangent_rowsplit.cuh)
  if constexpr (L == 1 && GC == NE && LA == 0 && NE * NT + 64 <= 1024) {
    const bool no_push = c->no_rs_push, no_st = c->no_rs_st;
    if constexpr (NT == 64 && NC <= 8) {   // every thread sends its own masses with st.async
      const int Ss = rs_ring_slots(c, rs_fw_st_smem<NE, NC, NT>(0), rs_fw_st_slot<NE, NC, NT>(), 2, 6);
      if (!no_push && !no_st && Ss >= 2) {
        int rc = ensure_tape_rs(c, P, NT, true);
        if (rc) return rc;
        rc = launch_cluster_grid(c, KIND_FT, k_forward_tangent_rs_st<NE, NC, NT>, ncl * NC, NC, NE * NT + 32, rs_fw_st_smem<NE, NC, NT>(Ss),
                                 "k_forward_tangent_rs_st", M, (const unsigned char*)c->tape_rs_fw,
                                 reinterpret_cast<const int*>(c->tape_rs_fw + tape_rs_info_offset(c, true)),
                                 reinterpret_cast<const int*>(c->tape_rs_fw + tape_rs_cnt_offset(c, true)), (const int*)c->d_status, K, Kp, Ss, c->pass_thi,
                                 (const double*)c->d_zero, dpol, pd_rs ? 1 : 0, dkdpart);
        if (rc >= 0) return rc;
      }
    }
    const int Sp = rs_ring_slots(c, rs_fw_push_smem<NE, NC, NT>(0), slot + 16, 2, 6);
    if (!no_push && Sp >= 2) {
      int rc = ensure_tape_rs(c, P, NT, true);
      if (rc) return rc;
      rc = launch_cluster_grid(c, KIND_FT, k_forward_tangent_rs_push<NE, NC, NT>, ncl * NC, NC, NE * NT + 64, rs_fw_push_smem<NE, NC, NT>(Sp),
                               "k_forward_tangent_rs_push", M, (const unsigned char*)c->tape_rs_fw,
                               reinterpret_cast<const int*>(c->tape_rs_fw + tape_rs_rng_offset(c, true)), K, Kp, Sp, c->pass_thi,
                               (const double*)c->d_zero, dpol, pd_rs ? 1 : 0, dkdpart);
      if (rc >= 0) return rc;
    }
    const bool no_ce = c->no_rs_ce;
    const int Sc = rs_ring_slots(c, rs_fw_ce_smem<NE, NT>(0), slot + 16, 2, 6);
    if (!no_ce && Sc >= 2) {
      int rc = ensure_tape_rs(c, P, NT, true);
      if (rc) return rc;
      rc = launch_cluster_grid(c, KIND_FT, k_forward_tangent_rs_ce<NE, NC, NT>, ncl * NC, NC, NE * NT + 64, rs_fw_ce_smem<NE, NT>(Sc),
                               "k_forward_tangent_rs_ce", M, (const unsigned char*)c->tape_rs_fw, K, Kp, Sc, c->pass_thi,
                               (const double*)c->d_zero, dpol, pd_rs ? 1 : 0, dkdpart, c->rs_relaxed ? 0 : 1);
      if (rc >= 0) return rc;   // (-1: this cluster cannot be scheduled with the larger CTAs; use the kernel below)
    }
  }
  const int S = rs_ring_slots(c, rs_fw_smem<NT, L, GC, LA>(0), slot + 16, LA + 2, GC == 1 ? 3 * NE : 6);
  if (S < LA + 2) return -1;
  int rc = ensure_tape_rs(c, P, NT, true);
  if (rc) return rc;
  return launch_cluster_grid(c, KIND_FT, k_forward_tangent_rs<NE, NC, NT, L, GC, LA>, ncl * NC, NC, NT + 64,
                             rs_fw_smem<NT, L, GC, LA>(S), "k_forward_tangent_rs", M, (const unsigned char*)c->tape_rs_fw, K, Kp, S,
                             c->pass_thi, (const double*)c->d_zero, dpol, pd_rs ? 1 : 0, dkdpart, c->rs_relaxed ? 0 : 1);
}
// clusters of this shape that can be resident at once (the smaller of the two sweeps' answers)
template <int NE, int NC, int NT, int L, int GC, int LA>
static int rs_cap_shape(hank_ctx* c) {
  int cap = 1 << 30;
  auto ask = [&](auto kern, size_t smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); cap = 0; return; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * 64); cfg.blockDim = dim3(NT + 64); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    cap = n < cap ? n : cap;
  };
  const int Sb = rs_ring_slots(c, rs_bw_smem<NT, L, GC, LA>(0), (size_t)GC * rs_bw_col_bytes<NT>() + 16, LA + 2, GC == 1 ? 3 * NE : 6);
  const int Sf = rs_ring_slots(c, rs_fw_smem<NT, L, GC, LA>(0), (size_t)GC * rs_fw_col_bytes<NT, L>() + 16, LA + 2, GC == 1 ? 3 * NE : 6);
  if (Sb < LA + 2 || Sf < LA + 2) return 0;
  ask(k_backward_tangent_rs<NE, NC, NT, L, GC, LA>, rs_bw_smem<NT, L, GC, LA>(Sb));
  ask(k_forward_tangent_rs<NE, NC, NT, L, GC, LA>, rs_fw_smem<NT, L, GC, LA>(Sf));
  return cap;
}
// shapes instantiated per n_e (rowsplit_cfg only returns these)
#define ROWSPLIT_DISPATCH(g, FN, ...)                                                                   \
  do {                                                                                                  \
    if (g.NC == 4 && g.NT == 64 && g.L == 1 && g.GC == NE) return FN<NE, 4, 64, 1, NE, 0>(__VA_ARGS__);   \
    if (g.NC == 8 && g.NT == 64 && g.L == 1 && g.GC == NE) return FN<NE, 8, 64, 1, NE, 0>(__VA_ARGS__);   \
    if (g.NC == 8 && g.NT == 128 && g.L == 1 && g.GC == NE) return FN<NE, 8, 128, 1, NE, 0>(__VA_ARGS__); \
    if (g.NC == 8 && g.NT == 256 && g.L == 1 && g.GC == 4) return FN<NE, 8, 256, 1, (NE < 4 ? NE : 4), 0>(__VA_ARGS__); \
    if (g.NC == 2 && g.NT == 256 && g.L == 1 && g.GC == NE) return FN<NE, 2, 256, 1, NE, 0>(__VA_ARGS__);  \
    if (g.NC == 4 && g.NT == 256 && g.L == 2 && g.GC == 4) return FN<NE, 4, 256, 2, (NE < 4 ? NE : 4), 0>(__VA_ARGS__); \
    if (g.NC == 4 && g.NT == 512 && g.L == 2 && g.GC == 2) return FN<NE, 4, 512, 2, 2, 0>(__VA_ARGS__);   \
  } while (0)
template <int NE>
int Sweeps<NE>::rs_max_clusters(hank_ctx* c, int NC, int NT, int L, int GC) {
  const TangentCfg g{NT, 1, L, NC, GC, 0};
  ROWSPLIT_DISPATCH(g, rs_cap_shape, c);
  return 0;
}
template <int NE>
int Sweeps<NE>::backward_tangent_rs(hank_ctx* c, int NC, int NT, int L, int GC, int P, int K, const double* dr,
                                    const double* dw, double* dpol) {
  const TangentCfg g{NT, 1, L, NC, GC, 0};
  ROWSPLIT_DISPATCH(g, bt_rs_launch, c, P, K, dr, dw, dpol);
  return -1;
}
template <int NE>
int Sweeps<NE>::forward_tangent_rs(hank_ctx* c, int NC, int NT, int L, int GC, int P, int K, const double* dpol,
                                   double* dkdpart) {
  const TangentCfg g{NT, 1, L, NC, GC, 0};
  ROWSPLIT_DISPATCH(g, ft_rs_launch, c, P, K, dpol, dkdpart);
  return -1;
}


}  // namespace hank
