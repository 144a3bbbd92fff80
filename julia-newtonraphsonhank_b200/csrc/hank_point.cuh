// hank_point.cuh — the per-grid-point algebra of the primal sweeps, written ONCE.
//
// The three backward primal kernels (one CTA: hank_kernels.cuh; cluster with a global-memory exchange:
// hank_primal_cluster.cuh; cluster with a distributed-shared-memory exchange: hank_primal_dsmem.cuh) and the three
// forward primal kernels differ only in how a column's values travel between income states.  What happens AT a
// grid point — Euler inversion, endogenous-grid interpolation with Flat extrapolation, the borrowing-constraint
// clamp, the marginal value, the lottery bracket and weight, and the linearisation coefficients the tangent
// sweeps replay — is the functions below; a semantic fix lands here and nowhere else.
// Operation order follows the reference line by line (the parity bar is 1e-10 relative on policies).
#pragma once
// (included by hank_kernels.cuh right after Consts / raise / pow_c / pow_v2 / lower_bound_fixed, which it uses)

namespace hank {

// ---- EGM step 1 at (a, e): c = (β·EV)^(-1/γ), endogenous knot k = ρ((c − w z) + a')   (KrusellSmith.jl:59-62)
// `ev` is Σ_e2 V⁺[a, e2]·Π[e, e2] accumulated by the caller in ascending e2 (value_next * Π').
// Tangent coefficients: ċ = Ḃ·yexp·B^(yexp−1);  k̇ = ρ·(ċ − ẇ z) + S·ρ̇,  ρ̇ = −(ρ/(1+r))·ṙ
//   a1 multiplies ĖV, kr multiplies ṙ (the ẇ coefficient −ρ z is formed by the tangent kernels).
struct EulerPoint { double knot, a1, kr; };
template <bool G2, int NE>
__device__ __forceinline__ EulerPoint egm_euler_point(const Consts<NE>& M, double ev, double wz, double ga, double rho,
                                                      double opr, int* status, int a, int e, int t) {
  const double B = M.beta * ev;
  if (B < 0.0) raise(status, 2, a, e, t);
  const double c = pow_c<G2>(B, M.yexp);
  const double S = (c - wz) + ga;
  EulerPoint o;
  o.knot = rho * S;
  o.a1 = rho * (M.beta * (M.yexp * (G2 ? c * c * c : c / B)));
  o.kr = -(S * (rho / opr));
  return o;
}

// ---- EGM steps 2-4 at (a, e): a'(x) by gridded-linear interpolation of (knots -> grid) with Flat extrapolation and
// the left-interval rule of Interpolations.jl, clamp at the borrowing constraint, consumption on the grid, marginal
// value V = (1+r)·c^-γ   (KrusellSmith.jl:65-80).  `ks` = this column's knots, `g` = the exogenous grid (both in
// shared memory).  Tangent coefficients:
//   δ̇ = (1/den)(−k̇_i) + (−num/den²)(k̇_{i+1} − k̇_i);  q̇ = δ̇ (g_{i+1} − g_i)  ->  ṗ = cA·k̇_i + cB·k̇_{i+1}  (0 when clamped / extrapolated)
//   V̇ = ṙ·cg^-γ + (1+r)·(−γ)·cg^(−γ−1)·ċg,  ċg = ṙ a + ẇ z − ṗ            ->  V̇ = vf·(z ẇ − ṗ) + E1·ṙ
struct InterpPoint { int i; double p, vnew, cA, cB, vf, E1; };
template <bool G2, int LDA, int NE>
__device__ __forceinline__ InterpPoint egm_interp_point(const Consts<NE>& M, const double* __restrict__ ks,
                                                        const double* __restrict__ g, int n_a, int a, double wz,
                                                        double opr, int* status, int e, int t) {
  const double x = g[a];
  if (a > 0 && !(ks[a] > ks[a - 1])) raise(status, 3, a, e, t);
  int i; double num, den; bool interior = true;
  const double k0 = ks[0], kl = ks[n_a - 1];
  if (x > kl) { i = n_a - 2; den = kl - ks[i]; num = den; interior = false; }
  else if (x < k0) { i = 0; den = ks[1] - k0; num = 0.0; interior = false; }
  else {
    int lb = lower_bound_fixed<LDA>(ks, n_a, x);  // searchsortedfirst - 1 (0-based count)
    i = min(max(lb, 1), n_a - 1) - 1;             // find_knot_index clamp, left knot 0-based
    num = x - ks[i]; den = ks[i + 1] - ks[i];
  }
  const double delta = num / den;
  const double gi = g[i], gi1 = g[i + 1];
  const double q = (1.0 - delta) * gi + delta * gi1;
  const bool cons = q < M.bc;
  InterpPoint o;
  o.i = i;
  o.p = cons ? M.bc : q;
  const double cg = (opr * x + wz) - o.p;
  if (cg < 0.0 && !M.gamma_int) raise(status, 2, a, e, t);
  double cgp, cgp1;
  pow_v2<G2>(cg, M.gamma, cgp, cgp1);
  o.vnew = opr * cgp;
  const bool live = interior && !cons;
  const double dg = gi1 - gi, id = 1.0 / den, nd2 = delta * id;
  o.cA = live ? (nd2 - id) * dg : 0.0;
  o.cB = live ? -(nd2 * dg) : 0.0;
  o.vf = opr * ((-M.gamma) * cgp1);
  o.E1 = cgp + o.vf * x;
  return o;
}

// ---- Young's lottery at a source point: bracket m = searchsortedfirst(grid, p) (1-based, bit-exact integer work),
// weight ω on grid[m-1] (ForwardIteration.jl:37-60), and the tangent coefficient D/Δg of ω̇·D = (D/Δg)·ṗ.
struct LotteryPoint { int m; double om, dco; };
template <int LDA>
__device__ __forceinline__ LotteryPoint lottery_point(const double* __restrict__ g, int n_a, double p, double D) {
  LotteryPoint o;
  o.m = lower_bound_fixed<LDA>(g, n_a, p) + 1;
  if (o.m == 1) { o.om = 1.0; o.dco = 0.0; }
  else if (o.m > n_a) { o.om = 0.0; o.dco = 0.0; }
  else {
    const double dgm = g[o.m - 1] - g[o.m - 2];
    o.om = (p - g[o.m - 2]) / dgm;
    o.dco = D / dgm;
  }
  return o;
}

// ---- source-range starts per destination row from the (monotone) brackets of a column: st[row] = first source a
// whose bracket is >= row; rows beyond the last bracket get n_a.  Called by the thread of source row a after the
// brackets `ms` of the column are in shared memory.
__device__ __forceinline__ void lottery_starts_point(const int* __restrict__ ms, int* __restrict__ st, int n_a, int a,
                                                     int* status, int e, int t) {
  const int hi = ms[a];
  const int lo = a == 0 ? 0 : ms[a - 1];
  if (hi < lo) raise(status, 6, a, e, t);
  for (int row = lo + 1; row <= hi; ++row) st[row] = a;
  if (a == n_a - 1)
    for (int row = max(hi, lo) + 1; row <= n_a + 2; ++row) st[row] = n_a;
}

// ---- destination row a: masses of its two contiguous source ranges in ascending source order (the order of
// Julia's CSC SpMV in transition_step, ForwardIteration.jl:63-99)
__device__ __forceinline__ double lottery_gather_point(const double* __restrict__ X, const double* __restrict__ Y,
                                                       int s0, int s1, int s2) {
  double acc = 0.0;
  for (int b = s0; b < s1; ++b) acc += X[b];
  for (int b = s1; b < s2; ++b) acc += Y[b];
  return acc;
}

}  // namespace hank
