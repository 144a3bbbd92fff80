// hank_tangent.cuh — batched tangent-lane sweeps, register-prefetch variant.
//
// Used where the TMA-staged ring of hank_tangent_tma.cuh does not fit in shared memory (LDA = 2048)
// or when HANK_NO_TMA=1.  Same recursion, same per-column pipeline; the tape coefficients of column
// e+1 are requested with plain loads into registers before the barrier of column e.
//
// Both kernels are linear recursions over the periods with coefficients from the primal tape.
// One CTA carries L lanes; the recursion state (V̇ or Ḋ, L*R*NE doubles per thread) lives in
// registers for the whole sweep and the only per-period HBM traffic is the policy tangent ṗ
// (8 B per point and lane: written once by the backward sweep, read once by the forward sweep).
#pragma once
#include "hank_kernels.cuh"

namespace hank {

// ======================================================================================
// Backward tangent sweep: per period t = P..1 and lane l
//   ĖV = V̇⁺ Πᵀ;  k̇ = a1·ĖV + kr·ṙ − ρ z ẇ;  ṗ = cA·k̇[i] + cB·k̇[i+1];  V̇ = E1·ṙ + vf·(z ẇ − ṗ)
// CTA b carries lanes [b*L, b*L+L). smem: kb[2][L][LDA].
// dr/dw: [K][P]; dvalT: [K][NE][LDA] or null; dpol: [P][NE][Kp][LDA].
// ======================================================================================
template <int NE, int R, int NT, int L>
__global__ void __launch_bounds__(NT, 1)
k_backward_tangent(const Consts<NE> M, const Tape tp, int K, int Kp, const int* __restrict__ thi,
                   const double* __restrict__ dr, const double* __restrict__ dw, const double* __restrict__ dvalT,
                   double* __restrict__ dpol, double* __restrict__ dvalue_first) {
  constexpr int LDA = NT * R;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  const int n_a = M.n_a, P = M.P;
  const int tid = threadIdx.x;
  const int lane0 = blockIdx.x * L;
  bool rowok[R];
#pragma unroll
  for (int j = 0; j < R; ++j) rowok[j] = tid + j * NT < n_a;

  double Vd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e)
        Vd[l][j][e] = (dvalT && rowok[j] && lane0 + l < K)
            ? dvalT[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] : 0.0;

  // coefficient registers of the column in flight: [R][a1,kr,cA,cB,E1,vf] + idx
  double cf[R][BW_NF]; int ci[R];
  auto load_col = [&](int t, int e, double (&c)[R][BW_NF], int (&ix)[R]) {
    const double* bwt = bw_fields<LDA>(tp, NE, t, e) + tid;
    const int* idxt = bw_idx<LDA>(tp, NE, t, e) + tid;
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int f = 0; f < BW_NF; ++f) c[j][f] = rowok[j] ? __ldg(bwt + f * LDA + j * NT) : 0.0;
      ix[j] = rowok[j] ? __ldg(idxt + j * NT) : 0;
    }
  };
  const int Pe = thi ? min(P, thi[lane0 / kThiGroup]) : P;   // seeds are zero from period Pe on
  if (Pe > 0) load_col(Pe - 1, 0, cf, ci);
  int pb = 0;
  for (int t = Pe - 1; t >= 0; --t) {
    const double rho = __ldg(tp.rho + t);
    double drl[L], dwl[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const bool on = lane0 + l < K;
      drl[l] = on ? __ldg(dr + (size_t)(lane0 + l) * P + t) : 0.0;
      dwl[l] = on ? __ldg(dw + (size_t)(lane0 + l) * P + t) : 0.0;
    }
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
    // ---- column pipeline
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      double cn[R][BW_NF]; int cin[R];
      if (e + 1 < NE) load_col(t, e + 1, cn, cin);
      else if (t > 0) load_col(t - 1, 0, cn, cin);
      double* kb = smem + (size_t)pb * L * LDA;
      const double cw = -(rho * M.z[e]);
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (rowok[j]) {
#pragma unroll
          for (int l = 0; l < L; ++l)
            kb[l * LDA + j * NT + tid] = fma(cf[j][BW_A1], Vd[l][j][e], fma(cf[j][BW_KR], drl[l], cw * dwl[l]));
        }
      __syncthreads();
      const double ze = M.z[e];
      double* dpc = dpol + (((size_t)t * NE + e) * Kp + lane0) * LDA + tid;
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (rowok[j]) {
          const double* kk = kb + ci[j];
#pragma unroll
          for (int l = 0; l < L; ++l) {
            const double pd = fma(cf[j][BW_CA], kk[l * LDA], cf[j][BW_CB] * kk[l * LDA + 1]);
            if (lane0 + l < K) __stcs(dpc + (size_t)l * LDA + j * NT, pd);
            Vd[l][j][e] = fma(cf[j][BW_VF], fma(ze, dwl[l], -pd), cf[j][BW_E1] * drl[l]);
          }
        }
      pb ^= 1;
      if (e + 1 < NE || t > 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
#pragma unroll
          for (int f = 0; f < BW_NF; ++f) cf[j][f] = cn[j][f];
          ci[j] = cin[j];
        }
      }
    }
  }
  if (dvalue_first) {
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (lane0 + l < K)
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (rowok[j])
#pragma unroll
            for (int e = 0; e < NE; ++e)
              dvalue_first[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] = Vd[l][j][e];
  }
}

// Gather of one destination row from the two staged source ranges: the first U sources of each
// range with predicated loads, longer ranges (the mass piling up at the borrowing constraint)
// finished by the whole warp, one row at a time.  Must be called by all 32 lanes of a warp.
template <int L, int LDA, int U>
__device__ __forceinline__ void gather_row(const double* __restrict__ xb, const double* __restrict__ yb, int s0,
                                           int s1, int s2, int lane, double (&acc)[L]) {
  const int n1 = s1 - s0, n2 = s2 - s1;
  // Most rows have one source per range: the second round is skipped by warps in which no row needs it.
  if constexpr (L > 4) {
    // wide shapes: absent sources load as zeros and the adds form a short tree (measured 15 % faster
    // at L = 6, 2 % slower at L = 4 than the predicated chain below)
#pragma unroll
    for (int d = 0; d < U; ++d) {
      if (d > 0 && !__any_sync(0xffffffffu, (d < n1) | (d < n2))) break;
      double x[L], y[L];
#pragma unroll
      for (int l = 0; l < L; ++l) x[l] = d < n1 ? xb[l * LDA + s0 + d] : 0.0;
#pragma unroll
      for (int l = 0; l < L; ++l) y[l] = d < n2 ? yb[l * LDA + s1 + d] : 0.0;
#pragma unroll
      for (int l = 0; l < L; ++l) acc[l] = d == 0 ? x[l] + y[l] : acc[l] + (x[l] + y[l]);
    }
  } else {
#pragma unroll
    for (int l = 0; l < L; ++l) acc[l] = 0.0;
#pragma unroll
    for (int d = 0; d < U; ++d) {
      if (d > 0 && !__any_sync(0xffffffffu, (d < n1) | (d < n2))) break;
      if (d < n1) {
#pragma unroll
        for (int l = 0; l < L; ++l) acc[l] += xb[l * LDA + s0 + d];
      }
      if (d < n2) {
#pragma unroll
        for (int l = 0; l < L; ++l) acc[l] += yb[l * LDA + s1 + d];
      }
    }
  }
  // a few more sources: finish in this thread; long ranges: the whole warp, one row at a time
  constexpr int kSerial = 8;
  const int mx = max(n1, n2) - U;
  if (mx > 0 && mx <= kSerial) {
    for (int b = s0 + U; b < s1; ++b)
#pragma unroll
      for (int l = 0; l < L; ++l) acc[l] += xb[l * LDA + b];
    for (int b = s1 + U; b < s2; ++b)
#pragma unroll
      for (int l = 0; l < L; ++l) acc[l] += yb[l * LDA + b];
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
      for (int b = b0 + U + lane; b < b1; b += 32) v += xb[l * LDA + b];
      for (int b = b1 + U + lane; b < b2; b += 32) v += yb[l * LDA + b];
      v = warp_sum(v);
      if (lane == src) acc[l] += v;
    }
  }
}

// ======================================================================================
// Forward tangent sweep: per period t = 1..P and lane l
//   ẋ = ω Ḋ + (D/Δg) ṗ, ẏ = Ḋ − ẋ;  ṫmp[row] = Σ_{m=row} ẋ + Σ_{m=row+1} ẏ;  Ḋ⁺ = ṫmp Π
//   K̇D_t = <ṗ_t, D_t> + <p_t, Ḋ_t>
// smem: Xb[2][L][LDA] | Yb[2][L][LDA].  dpol: [P][NE][Kp][LDA].  dkdpart: [K][P][NT/32]
// ======================================================================================
template <int NE, int R, int NT, int L>
__global__ void __launch_bounds__(NT, 1)
k_forward_tangent(const Consts<NE> M, const Tape tp, int K, int Kp, const int* __restrict__ thi,
                  const double* __restrict__ dpol, const double* __restrict__ dD0, double* __restrict__ dkdpart,
                  double* __restrict__ dD_last) {
  constexpr int LDA = NT * R, U = 2;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  const int n_a = M.n_a, P = M.P;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = NT / 32;
  const int lane0 = blockIdx.x * L;
  double* Xb = smem;
  double* Yb = smem + (size_t)2 * L * LDA;
  bool rowok[R];
#pragma unroll
  for (int j = 0; j < R; ++j) rowok[j] = tid + j * NT < n_a;
  bool laneon[L];
#pragma unroll
  for (int l = 0; l < L; ++l) laneon[l] = lane0 + l < K;

  double Dd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e)
        Dd[l][j][e] = (dD0 && rowok[j] && laneon[l]) ? dD0[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] : 0.0;

  const int pe = thi ? min(P, thi[lane0 / kThiGroup]) : P;   // ṗ is zero (and unwritten) from period pe on
  // column in flight: ω, D/Δg, D_t, p_t, the row's source-range starts, and ṗ of the L lanes
  double cf[R][FW_NF], pdv[L][R]; int sv[R][3];
  auto load_col = [&](int t, int e, double (&c)[R][FW_NF], double (&pd)[L][R], int (&s)[R][3]) {
    const double* fwt = fw_fields<LDA>(tp, NE, t, e) + tid;
    const int* st = fw_start<LDA>(tp, NE, t, e) + tid + 1;
    const double* dpc = dpol + (((size_t)t * NE + e) * Kp + lane0) * LDA + tid;
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int f = 0; f < FW_NF; ++f) c[j][f] = rowok[j] ? __ldg(fwt + f * LDA + j * NT) : 0.0;
#pragma unroll
      for (int q = 0; q < 3; ++q) s[j][q] = rowok[j] ? __ldg(st + j * NT + q) : 0;
#pragma unroll
      for (int l = 0; l < L; ++l) pd[l][j] = (rowok[j] && laneon[l] && t < pe) ? __ldcs(dpc + (size_t)l * LDA + j * NT) : 0.0;
    }
  };
  load_col(0, 0, cf, pdv, sv);
  int pb = 0;
  for (int t = 0; t < P; ++t) {
    double kacc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) kacc[l] = 0.0;
    double pv[R][NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      double cn[R][FW_NF], pdn[L][R]; int sn[R][3];
      if (e + 1 < NE) load_col(t, e + 1, cn, pdn, sn);
      else if (t + 1 < P) load_col(t + 1, 0, cn, pdn, sn);
      double* xb = Xb + (size_t)pb * L * LDA;
      double* yb = Yb + (size_t)pb * L * LDA;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        pv[j][e] = cf[j][FW_P];
        if (rowok[j]) {
#pragma unroll
          for (int l = 0; l < L; ++l) {
            const double xd = fma(cf[j][FW_OM], Dd[l][j][e], cf[j][FW_DCO] * pdv[l][j]);
            xb[l * LDA + j * NT + tid] = xd;
            yb[l * LDA + j * NT + tid] = Dd[l][j][e] - xd;
            kacc[l] = fma(pdv[l][j], cf[j][FW_D], kacc[l]);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < R; ++j) {
        double acc[L];
        gather_row<L, LDA, U>(xb, yb, sv[j][0], sv[j][1], sv[j][2], lane, acc);  // starts are 0 beyond n_a
#pragma unroll
        for (int l = 0; l < L; ++l) Dd[l][j][e] = acc[l];
      }
      pb ^= 1;
      if (e + 1 < NE || t + 1 < P) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
#pragma unroll
          for (int f = 0; f < FW_NF; ++f) cf[j][f] = cn[j][f];
#pragma unroll
          for (int q = 0; q < 3; ++q) sv[j][q] = sn[j][q];
#pragma unroll
          for (int l = 0; l < L; ++l) pdv[l][j] = pdn[l][j];
        }
      }
    }
    // ---- Markov mix (in place) and second aggregation term <p_t, Ḋ_t>
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (rowok[j]) {
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
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const double s = warp_sum(kacc[l]);
      if (lane == 0 && laneon[l]) dkdpart[((size_t)(lane0 + l) * P + t) * NW + warp] = s;
    }
  }
  if (dD_last) {
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (laneon[l])
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (rowok[j])
#pragma unroll
            for (int e = 0; e < NE; ++e)
              dD_last[(size_t)(lane0 + l) * GP + e * LDA + j * NT + tid] = Dd[l][j][e];
  }
}

}  // namespace hank
