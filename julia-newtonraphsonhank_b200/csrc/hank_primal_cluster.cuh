// hank_primal_cluster.cuh — primal sweeps on a thread-block cluster, one CTA per income state.
//
// The single-CTA primal sweeps (hank_kernels.cuh) are issue/latency bound on one SM
// (profiles/r01_notes.md: 7.9 and 11.6 us per period at 500x7).  Everything that couples the
// asset dimension — the endogenous-grid bracket search, the lottery gather — stays inside one
// income-state column, so column e is given to CTA e of a cluster of NE CTAs (NE <= 8 portable,
// <= 16 with the non-portable attribute).  The only coupling across columns is the Markov mix,
// which needs one G-sized array from all columns per period (∂V/∂a' backward, the post-lottery
// masses forward): each CTA publishes its column to a double-buffered exchange array in global
// memory (L2-resident, 8*G bytes) and the cluster meets at ONE barrier.cluster per period
// (release/acquire at cluster scope orders the global writes; readers use ld.global.cg).
// Results are identical to the single-CTA kernels: same operations in the same order per point.
#pragma once
#include <cooperative_groups.h>
#include "hank_kernels.cuh"

namespace hank {
namespace cg = cooperative_groups;

// ======================================================================================
// Backward primal sweep, cluster version (see k_backward_primal for the per-point algebra).
// smem: ks[LDA] knots of this column | g[LDA].  xch: [2][NE][LDA] exchange of ∂V/∂a.
// ======================================================================================
template <int NE, int R, int NT, bool G2>
__global__ void __launch_bounds__(NT, 1)
k_backward_primal_cl(const Consts<NE> M, const Tape tp, const double* __restrict__ grid,
                     const double* __restrict__ valueT, const double* __restrict__ rpath,
                     const double* __restrict__ wpath, double* __restrict__ xch, int* __restrict__ status) {
  constexpr int LDA = NT * R;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int e = (int)cluster.block_rank();
  const int n_a = M.n_a, P = M.P;
  double* ks = smem;
  double* g = smem + LDA;
  const int tid = threadIdx.x;
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
  __syncthreads();
  const double* vsrc = valueT;  // [NE][LDA]
  double rn = rpath[P - 1], wn = wpath[P - 1];
  double Vlast[R];
  for (int t = P - 1; t >= 0; --t) {
    const double r = rn, w = wn;
    if (t > 0) { rn = rpath[t - 1]; wn = wpath[t - 1]; }
    const double opr = 1.0 + r, rho = 1.0 / opr;
    if (tid == 0 && e == 0) tp.rho[t] = rho;
    double* bwf = bw_fields<LDA>(tp, NE, t, e) + tid;
    int* bwi = bw_idx<LDA>(tp, NE, t, e) + tid;
    double* polt = tp.pol + (size_t)t * GP + (size_t)e * LDA + tid;
    // ---- phase 1: Euler inversion for this column
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        double vrow[NE];
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) vrow[e2] = __ldcg(vsrc + (size_t)e2 * LDA + a);
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
    double* vdst = xch + (size_t)(t & 1) * GP + (size_t)e * LDA;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const InterpPoint q = egm_interp_point<G2, LDA>(M, ks, g, n_a, a, w * ze, opr, status, e, t);
        polt[j * NT] = q.p;
        bwi[j * NT] = q.i;
        bwf[BW_CA * LDA + j * NT] = q.cA;
        bwf[BW_CB * LDA + j * NT] = q.cB;
        bwf[BW_VF * LDA + j * NT] = q.vf;
        bwf[BW_E1 * LDA + j * NT] = q.E1;
        vdst[a] = q.vnew;
        Vlast[j] = q.vnew;
      }
    }
    cluster.sync();   // publishes this period's ∂V/∂a column; also orders the reuse of ks
    vsrc = xch + (size_t)(t & 1) * GP;
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int a = tid + j * NT;
    if (a < n_a) tp.value_first[(size_t)e * LDA + a] = Vlast[j];
  }
}

// ======================================================================================
// Forward primal sweep, cluster version (see k_forward_primal).
// smem: X[LDA] | Y[LDA] | g[LDA] | ms[LDA] (int) | st[LDA+4] (int)
// xch: [2][NE][LDA] exchange of the post-lottery masses. kdpart: [P][NE*NT/32].
// ======================================================================================
template <int NE, int R, int NT>
__global__ void __launch_bounds__(NT, 1)
k_forward_primal_cl(const Consts<NE> M, const Tape tp, const double* __restrict__ grid,
                    const double* __restrict__ D0, const double* __restrict__ pol_in,
                    double* __restrict__ xch, double* __restrict__ kdpart, int* __restrict__ status) {
  constexpr int LDA = NT * R, NS = LDA + 4, NW = NT / 32;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int e = (int)cluster.block_rank();
  const int n_a = M.n_a, P = M.P;
  double* X = smem;
  double* Y = X + LDA;
  double* g = Y + LDA;
  int* ms = reinterpret_cast<int*>(g + LDA);
  int* st = ms + LDA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
  __syncthreads();
  for (int t = 0; t < P; ++t) {
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
    // ---- phase C: gather in ascending source order; publish this column's masses
    double* tdst = xch + (size_t)(t & 1) * GP + (size_t)e * LDA;
    int* so = fw_start<LDA>(tp, NE, t, e);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const int s0 = st[a + 1], s1 = st[a + 2], s2 = st[a + 3];
        so[a + 1] = s0;
        if (a == n_a - 1) { so[a + 2] = s1; so[a + 3] = s2; }
        const double acc = lottery_gather_point(X, Y, s0, s1, s2);
        tdst[a] = acc;
      }
    }
    cluster.sync();
    // ---- Markov mix for this column and its share of <p_t, D_t>
    const double* tsrc = xch + (size_t)(t & 1) * GP;
    double kacc = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        double d = 0.0;
#pragma unroll
        for (int e1 = 0; e1 < NE; ++e1) d += pi_col[e1] * __ldcg(tsrc + (size_t)e1 * LDA + a);
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
}

}  // namespace hank
