// hank_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Single-threaded FP64 CPU restatement of the sequence-space HANK household
// block of vasudeva-ram/Julia-NewtonRaphsonHANK (reference files cited per
// function as `File.jl:lines`).  It is the checker the CUDA path is compared
// against in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Nothing under julia-newtonraphsonhank_b200/ may
// import, link or call it.
//
// PARITY UNPINNED: the reference ships no golden vectors and Julia is not
// installed in this image, so this restatement cannot be checked against the
// reference's own output.  It follows the reference source line by line, plus
// the documented semantics of three third-party packages that are not vendored
// in /root/reference: Interpolations.jl 0.16.2 (gridded linear + Flat),
// ForwardDiff 1.3.2 / DiffRules 1.15.1 (dual arithmetic; the vendored
// ForwardDiff.jl/src copy is used as the spec) and IterativeSolvers 0.9.4
// (gmres!).  tests/test_oracle_*.py pin it instead against an independent numpy
// restatement, finite differences and structural invariants.
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
// -ffp-contract=off keeps every a*b+c as two roundings, like Julia's generic code.
//
// Layouts (Julia column-major):
//   grid matrices  n_a x n_e, a fastest:      idx = e*n_a + a           (0-based here)
//   Pi             n_e x n_e column-major:    Pi[e + n_e*e2] = Π[e,e2]  (row-stochastic)
//   x              n_endog x P, var fastest:  x[v + 4*t], v = Y,KS,r,w
//   lanes          lane-major:                d*[l*len + i]
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>

namespace {

struct Model {
  int n_a, n_e, T;
  std::vector<double> grid, z, Pi;  // Pi column-major as passed
  double beta, gamma, borrow_cons;
  int G() const { return n_a * n_e; }
  int P() const { return T - 1; }
  double pi(int e, int e2) const { return Pi[e + (size_t)n_e * e2]; }
};

// ---- Julia Base `^` for Float64 (base/math.jl, Julia 1.11: pow_body) -------------------
// Integer-valued exponents take the power-by-squaring path (n == -2 -> inv(x)^2,
// n == 3 -> x*x*x, otherwise compensated squaring); everything else is a <1 ulp pow, for
// which glibc's correctly-rounded-in-practice pow() stands in.
inline void two_mul(double a, double b, double& hi, double& lo) {
  hi = a * b;
  lo = std::fma(a, b, -hi);
}
double jl_pow_int(double x, long long n) {
  if (n == 0) return 1.0;
  double y = 1.0, xnlo = 0.0, ynlo = 0.0;
  if (n == 3) return x * x * x;
  if (n < 0) {
    double rx = 1.0 / x;
    if (n == -2) return rx * rx;
    if (std::isfinite(x)) xnlo = -std::fma(x, rx, -1.0) * rx;
    x = rx;
    n = -n;
  }
  while (n > 1) {
    if (n & 1) {
      double err = y * xnlo + x * ynlo;
      double yh, yl;
      two_mul(x, y, yh, yl);
      y = yh;
      ynlo = yl + err;
    }
    double err = x * 2 * xnlo;
    double xh, xl;
    two_mul(x, x, xh, xl);
    x = xh;
    xnlo = xl + err;
    n >>= 1;
  }
  double err = y * xnlo + x * ynlo;
  double r = std::fma(x, y, err);
  return (std::isfinite(x) && std::isfinite(err)) ? r : x * y;
}
// returns NaN-with-domain-flag through *domain_err when x<0 and y non-integer
double jl_pow(double x, double y, int* domain_err) {
  if (x == 1.0) return 1.0;
  if (std::fabs(y) < 0x1.8p62) {
    long long yint = (long long)y;
    if ((double)yint == y) return jl_pow_int(x, yint);
  }
  if (x < 0 && std::isfinite(y)) {
    if (domain_err) *domain_err = 1;
    return std::numeric_limits<double>::quiet_NaN();
  }
  return std::pow(x, y);
}

// Julia searchsortedfirst(v, x): smallest 1-based i with v[i] >= x (isless order), n+1 if none.
// Returned 1-based to keep the reference's index algebra.
inline int searchsortedfirst1(const double* v, int n, double x) {
  int lo = 0, hi = n + 1;  // Base sort.jl: lo = first-1, hi = last+1
  while (lo < hi - 1) {
    int m = lo + ((hi - lo) >> 1);
    if (v[m - 1] < x) lo = m; else hi = m;
  }
  return hi;
}

enum { OK = 0, ERR_DOMAIN = 2, ERR_KNOTS = 3, ERR_ARG = 1 };

// One EGM step with K tangent lanes — KrusellSmith.jl:43-83 (ValueFunction); dual rules per
// ForwardDiff.jl/src/dual.jl:495-581, partials.jl:85-87,219-221; interpolation per
// Interpolations.jl gridded linear + Flat.
// info[0..2] = (kind, a, e) of the first error, 1-based a/e.
int egm_step(const Model& m, const double* vnext, const double* dvnext, double r, double w, int K,
             const double* dr, const double* dw, double* value, double* policy, double* dvalue,
             double* dpolicy, int* info, int32_t* interval_out) {
  const int n_a = m.n_a, n_e = m.n_e, G = m.G();
  const double* g = m.grid.data();
  std::vector<double> knots(G), cmat(G), Bmat(G);
  std::vector<double> dknots((size_t)K * G);
  const double yexp = -1.0 / m.gamma;
  const double rho = 1.0 / (1.0 + r);
  // Step 1 (KrusellSmith.jl:59) and step 2 (:62)
  for (int e = 0; e < n_e; ++e)
    for (int a = 0; a < n_a; ++a) {
      double ev = 0.0;
      for (int e2 = 0; e2 < n_e; ++e2) ev += vnext[e2 * n_a + a] * m.pi(e, e2);
      double B = m.beta * ev;
      int dom = 0;
      double c = jl_pow(B, yexp, &dom);
      if (dom) {
        if (info) { info[0] = ERR_DOMAIN; info[1] = a + 1; info[2] = e + 1; }
        return ERR_DOMAIN;
      }
      Bmat[e * n_a + a] = B;
      cmat[e * n_a + a] = c;
      double S = (c - w * m.z[e]) + g[a];
      knots[e * n_a + a] = rho * S;
    }
  for (int l = 0; l < K; ++l) {
    const double* dv = dvnext ? dvnext + (size_t)l * G : nullptr;
    const double drl = dr ? dr[l] : 0.0, dwl = dw ? dw[l] : 0.0;
    const double drho = -(rho / (1.0 + r)) * drl;  // dual.jl:535-539 (Real / Dual)
    for (int e = 0; e < n_e; ++e)
      for (int a = 0; a < n_a; ++a) {
        double dev = 0.0;
        if (dv)
          for (int e2 = 0; e2 < n_e; ++e2) dev += dv[e2 * n_a + a] * m.pi(e, e2);
        double dB = m.beta * dev;
        double B = Bmat[e * n_a + a], c = cmat[e * n_a + a];
        // dual.jl:563-572: partials(x) * y * f(v, y - 1)
        double dc = (dB == 0.0) ? 0.0 : dB * yexp * jl_pow(B, yexp - 1.0, nullptr);
        double S = (c - w * m.z[e]) + g[a];
        double dS = dc - dwl * m.z[e];
        dknots[(size_t)l * G + e * n_a + a] = S * drho + rho * dS;  // Dual*Dual: vy*dx + vx*dy
      }
  }
  // Step 3 (:65-73): per-column gridded linear interpolation, Flat extrapolation
  const double opr = 1.0 + r;
  for (int e = 0; e < n_e; ++e) {
    const double* k = &knots[e * n_a];
    for (int a = 1; a < n_a; ++a)
      if (!(k[a] > k[a - 1])) {  // Interpolations.jl check_gridded: sorted and unique
        if (info) { info[0] = ERR_KNOTS; info[1] = a + 1; info[2] = e + 1; }
        return ERR_KNOTS;
      }
    for (int a = 0; a < n_a; ++a) {
      const double x = g[a];
      int clampflag = 0;  // -1: x̂ = k[1], +1: x̂ = k[n]
      double xh = x;
      if (x > k[n_a - 1]) { xh = k[n_a - 1]; clampflag = 1; }
      else if (x < k[0]) { xh = k[0]; clampflag = -1; }
      int i1 = searchsortedfirst1(k, n_a, xh) - 1;        // find_knot_index
      i1 = std::min(std::max(i1, 1), n_a - 1);           // clamp to [1, n-1], 1-based
      const int i = i1 - 1;                                // 0-based left knot
      const double l_ = k[i], u_ = k[i + 1];
      const double num = xh - l_, den = u_ - l_;
      const double delta = num / den;
      const double q = (1.0 - delta) * g[i] + delta * g[i + 1];
      const double p = (q < m.borrow_cons) ? m.borrow_cons : q;  // max.(q, bc)  (:76)
      const double cg = (opr * x + w * m.z[e]) - p;               // (:79)
      int dom = 0;
      const double cgp = jl_pow(cg, -m.gamma, &dom);
      if (dom) {
        if (info) { info[0] = ERR_DOMAIN; info[1] = a + 1; info[2] = e + 1; }
        return ERR_DOMAIN;
      }
      const int idx = e * n_a + a;
      policy[idx] = p;
      value[idx] = opr * cgp;  // (:80)
      if (interval_out) interval_out[idx] = i1;
      for (int l = 0; l < K; ++l) {
        const double* dk = &dknots[(size_t)l * G + e * n_a];
        const double drl = dr ? dr[l] : 0.0, dwl = dw ? dw[l] : 0.0;
        // x̂ carries the partial of the knot it was clamped to, else none
        const double dxh = clampflag > 0 ? dk[n_a - 1] : (clampflag < 0 ? dk[0] : 0.0);
        const double dnum = dxh - dk[i], dden = dk[i + 1] - dk[i];
        // _div_partials (partials.jl:85-87): inv(den)*dnum + (-(num/(den*den)))*dden
        const double ddelta = (1.0 / den) * dnum + (-(num / (den * den))) * dden;
        const double dq = (-ddelta) * g[i] + ddelta * g[i + 1];
        const double dp = (q < m.borrow_cons) ? 0.0 : dq;  // DiffRules max: ties pass x's partial
        const double dcg = (drl * x + dwl * m.z[e]) - dp;
        const double dcgp = dcg * (-m.gamma) * jl_pow(cg, -m.gamma - 1.0, nullptr);
        dpolicy[(size_t)l * G + idx] = dp;
        dvalue[(size_t)l * G + idx] = cgp * drl + opr * dcgp;
      }
    }
  }
  return OK;
}

// Lottery brackets — ForwardIteration.jl:46-75. m is Julia's 1-based searchsortedfirst result.
inline void lottery_point(const double* g, int n_a, double p, int& mj, double& omega) {
  mj = searchsortedfirst1(g, n_a, p);
  if (mj == 1 || mj > n_a) omega = 1.0;  // all mass on one node (weight one(eltype))
  else omega = (p - g[mj - 2]) / (g[mj - 1] - g[mj - 2]);
}

// D_new = Λ_exog * (Λ_endog(policy) * D) with K tangent lanes — ForwardIteration.jl:95-99 with
// the CSC SpMV accumulation order of SparseArrays (ascending source column).
void forward_step(const Model& m, const double* policy, const double* D, int K,
                  const double* dpolicy, const double* dD, double* Dn, double* dDn) {
  const int n_a = m.n_a, n_e = m.n_e, G = m.G();
  const double* g = m.grid.data();
  std::vector<double> tmp((size_t)(K + 1) * G, 0.0);
  for (int e = 0; e < n_e; ++e)
    for (int a = 0; a < n_a; ++a) {
      const int col = e * n_a + a;
      int mj; double om;
      lottery_point(g, n_a, policy[col], mj, om);
      if (mj == 1) {
        tmp[e * n_a + 0] += 1.0 * D[col];
        for (int l = 0; l < K; ++l) tmp[(size_t)(l + 1) * G + e * n_a + 0] += 1.0 * dD[(size_t)l * G + col];
      } else if (mj > n_a) {
        tmp[e * n_a + n_a - 1] += 1.0 * D[col];
        for (int l = 0; l < K; ++l) tmp[(size_t)(l + 1) * G + e * n_a + n_a - 1] += 1.0 * dD[(size_t)l * G + col];
      } else {
        const double dg = g[mj - 1] - g[mj - 2];
        const double om1 = 1.0 - om;
        tmp[e * n_a + mj - 2] += om1 * D[col];
        tmp[e * n_a + mj - 1] += om * D[col];
        for (int l = 0; l < K; ++l) {
          const double dom = dpolicy[(size_t)l * G + col] / dg;  // Dual / Real
          const double dDl = dD[(size_t)l * G + col];
          // Dual*Dual partials: vy*dx + vx*dy  (x = matrix entry, y = D)
          tmp[(size_t)(l + 1) * G + e * n_a + mj - 2] += D[col] * (-dom) + om1 * dDl;
          tmp[(size_t)(l + 1) * G + e * n_a + mj - 1] += D[col] * dom + om * dDl;
        }
      }
    }
  // Λ_exog = kron(sparse(Π'), I): column (a,e) scatters Π[e,e2] to row (a,e2)  (:280-284)
  for (int l = 0; l <= K; ++l) {
    double* out = l == 0 ? Dn : dDn + (size_t)(l - 1) * G;
    const double* t = &tmp[(size_t)l * G];
    std::fill(out, out + G, 0.0);
    for (int e = 0; e < n_e; ++e)
      for (int a = 0; a < n_a; ++a)
        for (int e2 = 0; e2 < n_e; ++e2) {
          const double pe = m.pi(e, e2);
          if (pe != 0.0) out[e2 * n_a + a] += pe * t[e * n_a + a];
        }
  }
}

// Krusell-Smith residuals with lag padding — KrusellSmith.yaml:90-94 compiled per
// ModelParser.jl:54-119,217-259; assemble_full_xMat GeneralStructures.jl:329-377;
// shift_lag :441-443.  Output equation-fastest. ss_start_KS pads KS(-1) at t=1.
void ks_residuals(int P, double alpha, double delta, double ss_start_KS, const double* x,
                  const double* KD, const double* Z, int K, const double* dx, const double* dKD,
                  double* F, double* dF) {
  for (int t = 0; t < P; ++t) {
    const double Y = x[4 * t + 0], KS = x[4 * t + 1], r = x[4 * t + 2], w = x[4 * t + 3];
    const double Kl = t == 0 ? ss_start_KS : x[4 * (t - 1) + 1];
    const double Ka = jl_pow(Kl, alpha, nullptr);
    const double Ka1 = jl_pow(Kl, alpha - 1.0, nullptr);
    F[4 * t + 0] = Y - (Z[t] * Ka);
    F[4 * t + 1] = (r + delta) - ((alpha * Z[t]) * Ka1);
    F[4 * t + 2] = w - (((1.0 - alpha) * Z[t]) * Ka);
    F[4 * t + 3] = KS - KD[t];
    if (K > 0) {
      const double Ka2 = jl_pow(Kl, alpha - 2.0, nullptr);
      for (int l = 0; l < K; ++l) {
        const double* d = dx + (size_t)l * 4 * P;
        const double dKl = t == 0 ? 0.0 : d[4 * (t - 1) + 1];
        const double dKa = dKl * alpha * Ka1;            // partials(x) * y * v^(y-1)
        const double dKa1 = dKl * (alpha - 1.0) * Ka2;
        double* o = dF + (size_t)l * 4 * P;
        o[4 * t + 0] = d[4 * t + 0] - (Z[t] * dKa);
        o[4 * t + 1] = d[4 * t + 2] - ((alpha * Z[t]) * dKa1);
        o[4 * t + 2] = d[4 * t + 3] - (((1.0 - alpha) * Z[t]) * dKa);
        o[4 * t + 3] = d[4 * t + 1] - dKD[(size_t)l * P + t];
      }
    }
  }
}

struct KSProblem {
  Model* m;
  double alpha, delta, ss_start_KS;
  const double* value_T;  // ss_end.value
  const double* D0;       // ss_initial.D
  const double* Z;
};

// Backward sweep — BackwardIteration.jl:46-116. policy[t*G + idx] for t = 0..P-1 (period t+1).
int backward(const Model& m, const double* value_T, const double* r, const double* w, int K,
             const double* dr, const double* dw, double* policy, double* dpolicy, int* info,
             double* value_first, double* dvalue_first) {
  const int G = m.G(), P = m.P();
  std::vector<double> v(value_T, value_T + G), vn(G);
  std::vector<double> dv((size_t)K * G, 0.0), dvn((size_t)K * G);
  std::vector<double> drl(K), dwl(K), dpol((size_t)K * G);
  for (int t = P - 1; t >= 0; --t) {
    for (int l = 0; l < K; ++l) { drl[l] = dr[(size_t)l * P + t]; dwl[l] = dw[(size_t)l * P + t]; }
    int rc = egm_step(m, v.data(), K ? dv.data() : nullptr, r[t], w[t], K, drl.data(), dwl.data(),
                      vn.data(), policy + (size_t)t * G, dvn.data(), dpol.data(), info, nullptr);
    if (rc) { if (info) info[3] = t + 1; return rc; }
    for (int l = 0; l < K; ++l)
      std::memcpy(dpolicy + ((size_t)l * P + t) * G, &dpol[(size_t)l * G], sizeof(double) * G);
    v.swap(vn);
    dv.swap(dvn);
  }
  if (value_first) std::memcpy(value_first, v.data(), sizeof(double) * G);
  if (dvalue_first && K) std::memcpy(dvalue_first, dv.data(), sizeof(double) * K * G);
  return OK;
}

// Forward sweep + aggregation — ForwardIteration.jl:253-311 (dot at :306 uses post-transition D).
void forward(const Model& m, const double* D0, const double* policy, int K, const double* dpolicy,
             double* KD, double* dKD, double* D_path, double* dD_last) {
  const int G = m.G(), P = m.P();
  std::vector<double> D(D0, D0 + G), Dn(G), dD((size_t)K * G, 0.0), dDn((size_t)K * G);
  std::vector<double> dpol((size_t)K * G);
  for (int t = 0; t < P; ++t) {
    const double* pol = policy + (size_t)t * G;
    for (int l = 0; l < K; ++l)
      std::memcpy(&dpol[(size_t)l * G], dpolicy + ((size_t)l * P + t) * G, sizeof(double) * G);
    forward_step(m, pol, D.data(), K, dpol.data(), dD.data(), Dn.data(), dDn.data());
    D.swap(Dn);
    dD.swap(dDn);
    double s = 0.0;
    for (int j = 0; j < G; ++j) s += pol[j] * D[j];
    KD[t] = s;
    for (int l = 0; l < K; ++l) {
      double ds = 0.0;
      const double* dp = &dpol[(size_t)l * G];
      const double* dd = &dD[(size_t)l * G];
      for (int j = 0; j < G; ++j) ds += D[j] * dp[j] + pol[j] * dd[j];  // vy*dx + vx*dy
      dKD[(size_t)l * P + t] = ds;
    }
    if (D_path) std::memcpy(D_path + (size_t)t * G, D.data(), sizeof(double) * G);
  }
  if (dD_last && K) std::memcpy(dD_last, dD.data(), sizeof(double) * K * G);
}

// fullFunction(x) and its K-lane JVP — NewtonRaphson.jl:77-83, GeneralStructures.jl:542-550.
// V is lane-major (Julia n x K column-major).
int full_function(const KSProblem& pb, const double* x, int K, const double* V, double* F,
                  double* JV, int* info) {
  const Model& m = *pb.m;
  const int P = m.P(), G = m.G();
  const int n = 4 * P;
  std::vector<double> r(P), w(P), dr((size_t)K * P), dw((size_t)K * P);
  for (int t = 0; t < P; ++t) { r[t] = x[4 * t + 2]; w[t] = x[4 * t + 3]; }
  for (int l = 0; l < K; ++l)
    for (int t = 0; t < P; ++t) {
      dr[(size_t)l * P + t] = V[(size_t)l * n + 4 * t + 2];
      dw[(size_t)l * P + t] = V[(size_t)l * n + 4 * t + 3];
    }
  std::vector<double> policy((size_t)P * G), dpolicy((size_t)K * P * G);
  int rc = backward(m, pb.value_T, r.data(), w.data(), K, dr.data(), dw.data(), policy.data(),
                    dpolicy.data(), info, nullptr, nullptr);
  if (rc) return rc;
  std::vector<double> KD(P), dKD((size_t)K * P);
  forward(m, pb.D0, policy.data(), K, dpolicy.data(), KD.data(), dKD.data(), nullptr, nullptr);
  ks_residuals(P, pb.alpha, pb.delta, pb.ss_start_KS, x, KD.data(), pb.Z, K, V, dKD.data(), F, JV);
  return OK;
}

// ---- IterativeSolvers.jl 0.9.4 gmres! (package not in tree; defaults: restart=min(20,n),
// maxiter=n, reltol=sqrt(eps), abstol=0, modified Gram-Schmidt, no preconditioner) -----------
double nrm2(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return std::sqrt(s);
}
void matvec(const double* A, int n, const double* x, double* y) {  // column-major dense
  std::fill(y, y + n, 0.0);
  for (int j = 0; j < n; ++j) {
    const double xj = x[j];
    if (xj == 0.0) continue;
    const double* c = A + (size_t)j * n;
    for (int i = 0; i < n; ++i) y[i] += c[i] * xj;
  }
}
int gmres(const double* A, int n, double* x, const double* b, int* iters_out) {
  const int restart = std::min(20, n), maxiter = n;
  const double reltol = std::sqrt(std::numeric_limits<double>::epsilon());
  std::vector<double> Vb((size_t)n * (restart + 1)), H((size_t)(restart + 1) * restart, 0.0);
  std::vector<double> Ax(n), nullvec(restart + 1), rhs(restart + 1);
  auto Hm = [&](int i, int j) -> double& { return H[(size_t)j * (restart + 1) + i]; };
  auto init = [&]() {  // V[:,1] = b - A x, normalised
    matvec(A, n, x, Ax.data());
    double* v1 = Vb.data();
    for (int i = 0; i < n; ++i) v1[i] = b[i] - Ax[i];
    double beta = nrm2(v1, n);
    double ib = 1.0 / beta;
    for (int i = 0; i < n; ++i) v1[i] *= ib;
    return beta;
  };
  double res_beta, accumulator, current;
  auto init_residual = [&](double beta) {
    accumulator = 1.0; res_beta = beta;
    std::fill(nullvec.begin(), nullvec.end(), 0.0); nullvec[0] = 1.0;
  };
  double beta = init();
  current = beta;
  init_residual(beta);
  const double tol = std::max(reltol * current, 0.0);
  int k = 1, iteration = 0;
  auto solve_and_update = [&](int kk) {  // least squares via Givens on H[1:kk, 1:kk-1]
    const int w = kk - 1;
    std::fill(rhs.begin(), rhs.end(), 0.0);
    rhs[0] = beta;
    std::vector<double> Hc(H);
    auto Hcm = [&](int i, int j) -> double& { return Hc[(size_t)j * (restart + 1) + i]; };
    for (int i = 0; i < w; ++i) {
      double a = Hcm(i, i), bb = Hcm(i + 1, i);
      double rr = std::hypot(a, bb);
      double c = a / rr, s = bb / rr;
      Hcm(i, i) = rr; Hcm(i + 1, i) = 0.0;
      for (int j = i + 1; j < w; ++j) {
        double t1 = Hcm(i, j), t2 = Hcm(i + 1, j);
        Hcm(i, j) = c * t1 + s * t2;
        Hcm(i + 1, j) = -s * t1 + c * t2;
      }
      double t1 = rhs[i], t2 = rhs[i + 1];
      rhs[i] = c * t1 + s * t2;
      rhs[i + 1] = -s * t1 + c * t2;
    }
    for (int i = w - 1; i >= 0; --i) {
      double s = rhs[i];
      for (int j = i + 1; j < w; ++j) s -= Hcm(i, j) * rhs[j];
      rhs[i] = s / Hcm(i, i);
    }
    for (int j = 0; j < w; ++j) {
      const double* vj = &Vb[(size_t)j * n];
      for (int i = 0; i < n; ++i) x[i] += vj[i] * rhs[j];
    }
  };
  while (true) {
    if (iteration >= maxiter || current <= tol) {
      if (k > 1) solve_and_update(k);
      break;
    }
    // expand: V[:,k+1] = A V[:,k]
    double* vk1 = &Vb[(size_t)k * n];
    matvec(A, n, &Vb[(size_t)(k - 1) * n], vk1);
    for (int i = 0; i < k; ++i) {  // modified Gram-Schmidt
      const double* vi = &Vb[(size_t)i * n];
      double h = 0.0;
      for (int j = 0; j < n; ++j) h += vi[j] * vk1[j];
      Hm(i, k - 1) = h;
      for (int j = 0; j < n; ++j) vk1[j] -= h * vi[j];
    }
    double nr = nrm2(vk1, n);
    double inr = 1.0 / nr;
    for (int j = 0; j < n; ++j) vk1[j] *= inr;
    Hm(k, k - 1) = nr;
    // update_residual!
    double d = 0.0;
    for (int i = 0; i < k; ++i) d += nullvec[i] * Hm(i, k - 1);
    nullvec[k] = -(d / Hm(k, k - 1));
    accumulator += nullvec[k] * nullvec[k];
    current = res_beta / std::sqrt(accumulator);
    ++k;
    if (k == restart + 1) {
      solve_and_update(k);
      k = 1;
      if (!(iteration >= maxiter || current <= tol)) {  // done(g, iteration) with the old count
        beta = init();
        init_residual(beta);
        current = beta;  // init_residual! also resets residual.current to the true residual of the new cycle
        std::fill(H.begin(), H.end(), 0.0);
      }
    }
    ++iteration;
  }
  if (iters_out) *iters_out = iteration;
  return OK;
}

// Dense LU with partial pivoting (for the non-reference "lu" inner solver and small systems).
int lu_factor(std::vector<double>& A, int n, std::vector<int>& piv) {
  piv.resize(n);
  for (int k = 0; k < n; ++k) {
    int p = k; double mx = std::fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i) { double v = std::fabs(A[(size_t)k * n + i]); if (v > mx) { mx = v; p = i; } }
    piv[k] = p;
    if (mx == 0.0) return 1;
    if (p != k) for (int j = 0; j < n; ++j) std::swap(A[(size_t)j * n + k], A[(size_t)j * n + p]);
    const double ip = 1.0 / A[(size_t)k * n + k];
    for (int i = k + 1; i < n; ++i) A[(size_t)k * n + i] *= ip;
    for (int j = k + 1; j < n; ++j) {
      const double akj = A[(size_t)j * n + k];
      if (akj == 0.0) continue;
      double* cj = &A[(size_t)j * n];
      const double* ck = &A[(size_t)k * n];
      for (int i = k + 1; i < n; ++i) cj[i] -= ck[i] * akj;
    }
  }
  return 0;
}
void lu_solve(const std::vector<double>& A, int n, const std::vector<int>& piv, double* b) {
  for (int k = 0; k < n; ++k) if (piv[k] != k) std::swap(b[k], b[piv[k]]);
  for (int k = 0; k < n; ++k) { const double bk = b[k]; if (bk != 0.0) for (int i = k + 1; i < n; ++i) b[i] -= A[(size_t)k * n + i] * bk; }
  for (int k = n - 1; k >= 0; --k) { b[k] /= A[(size_t)k * n + k]; const double bk = b[k]; for (int i = 0; i < k; ++i) b[i] -= A[(size_t)k * n + i] * bk; }
}

}  // namespace

extern "C" {

struct hanko_model { Model m; };

hanko_model* hanko_create(int n_a, int n_e, int T, const double* grid, const double* z,
                          const double* Pi, double beta, double gamma, double borrow_cons) {
  if (n_a < 2 || n_e < 1 || T < 2) return nullptr;
  hanko_model* h = new hanko_model;
  h->m.n_a = n_a; h->m.n_e = n_e; h->m.T = T;
  h->m.grid.assign(grid, grid + n_a);
  h->m.z.assign(z, z + n_e);
  h->m.Pi.assign(Pi, Pi + (size_t)n_e * n_e);
  h->m.beta = beta; h->m.gamma = gamma; h->m.borrow_cons = borrow_cons;
  return h;
}
void hanko_destroy(hanko_model* h) { delete h; }

// GeneralStructures.jl:474-483 (make_DoubleExponentialGrid).
void hanko_double_exponential_grid(double amin, double amax, int n, double* out) {
  const double U = std::log(1.0 + std::log(1.0 + amax - amin));
  for (int i = 0; i < n; ++i) {
    // Base.lerpi(j, d, a, b): t = j/d; fma(t, b, fma(-t, a, a)) with a = 0
    // range(0, U, n) is a TwicePrecision StepRangeLen: element i is i*U/(n-1) evaluated in
    // extended precision and rounded once; x87 long double stands in for it.
    const double u = (double)(((long double)i * (long double)U) / (long double)(n - 1));
    out[i] = amin + std::exp(std::exp(u) - 1.0) - 1.0;
  }
}

// GeneralStructures.jl:500-525 (get_RouwenhorstDiscretization) + ForwardIteration.jl:436-442
// (invariant_dist on the small dense chain). Pi_out column-major, row-stochastic.
int hanko_rouwenhorst(int n, double rho, double sigma, double* Pi_out, double* D_out, double* z_out) {
  const double p = (1.0 + rho) / 2.0;
  std::vector<double> Pi = {p, 1 - p, 1 - p, p};  // 2x2 symmetric, column-major
  int cur = 2;
  for (int i = 3; i <= n; ++i) {
    std::vector<double> N((size_t)i * i, 0.0);
    auto at = [&](std::vector<double>& M, int dim, int r, int c) -> double& { return M[(size_t)c * dim + r]; };
    for (int c = 0; c < cur; ++c)
      for (int r = 0; r < cur; ++r) at(N, i, r, c) += p * at(Pi, cur, r, c);
    for (int c = 0; c < cur; ++c)
      for (int r = 0; r < cur; ++r) at(N, i, r, c + 1) += (1 - p) * at(Pi, cur, r, c);
    for (int c = 0; c < cur; ++c)
      for (int r = 0; r < cur; ++r) at(N, i, r + 1, c) += (1 - p) * at(Pi, cur, r, c);
    for (int c = 0; c < cur; ++c)
      for (int r = 0; r < cur; ++r) at(N, i, r + 1, c + 1) += p * at(Pi, cur, r, c);
    for (int r = 1; r < i - 1; ++r)
      for (int c = 0; c < i; ++c) at(N, i, r, c) /= 2;
    Pi.swap(N);
    cur = i;
  }
  // invariant_dist(Π): ΠT = Π'; M = I - ΠT[2:end,2:end]; b = ΠT[2:end,1]; D = [1; M\b]; D/sum(D)
  const int k = n - 1;
  std::vector<double> D(n, 1.0);
  if (k > 0) {
    std::vector<double> M((size_t)k * k), b(k);
    for (int c = 0; c < k; ++c)
      for (int r = 0; r < k; ++r)  // ΠT[r+1, c+1] = Π[c+1, r+1]
        M[(size_t)c * k + r] = (r == c ? 1.0 : 0.0) - Pi[(size_t)(r + 1) * n + (c + 1)];
    for (int r = 0; r < k; ++r) b[r] = Pi[(size_t)(r + 1) * n + 0];  // ΠT[r+1,1] = Π[1,r+1]
    std::vector<int> piv;
    if (lu_factor(M, k, piv)) return ERR_ARG;
    lu_solve(M, k, piv, b.data());
    for (int r = 0; r < k; ++r) D[r + 1] = b[r];
  }
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += D[i];
  for (int i = 0; i < n; ++i) D[i] /= s;
  const double al = 2.0 * (sigma / std::sqrt((double)(n - 1)));
  double zs = 0.0;
  for (int i = 0; i < n; ++i) { z_out[i] = std::exp(al * (double)i); }
  for (int i = 0; i < n; ++i) zs += z_out[i] * D[i];
  for (int i = 0; i < n; ++i) z_out[i] = z_out[i] / zs;
  std::memcpy(Pi_out, Pi.data(), sizeof(double) * n * n);
  std::memcpy(D_out, D.data(), sizeof(double) * n);
  return OK;
}

int hanko_egm_step(hanko_model* h, const double* value_next, const double* dvalue_next, double r,
                   double w, int K, const double* dr, const double* dw, double* value,
                   double* policy, double* dvalue, double* dpolicy, int* info, int32_t* interval) {
  return egm_step(h->m, value_next, dvalue_next, r, w, K, dr, dw, value, policy, dvalue, dpolicy, info, interval);
}

// ForwardIteration.jl:46-75: m (1-based searchsortedfirst) and lottery weight per point
void hanko_lottery(hanko_model* h, const double* policy, int32_t* m_out, double* omega_out) {
  const Model& m = h->m;
  for (int e = 0; e < m.n_e; ++e)
    for (int a = 0; a < m.n_a; ++a) {
      int mj; double om;
      lottery_point(m.grid.data(), m.n_a, policy[e * m.n_a + a], mj, om);
      m_out[e * m.n_a + a] = mj;
      if (omega_out) omega_out[e * m.n_a + a] = om;
    }
}

void hanko_forward_step(hanko_model* h, const double* policy, const double* D, int K,
                        const double* dpolicy, const double* dD, double* Dn, double* dDn) {
  forward_step(h->m, policy, D, K, dpolicy, dD, Dn, dDn);
}

int hanko_backward(hanko_model* h, const double* value_T, const double* r, const double* w, int K,
                   const double* dr, const double* dw, double* policy, double* dpolicy, int* info,
                   double* value_first, double* dvalue_first) {
  return backward(h->m, value_T, r, w, K, dr, dw, policy, dpolicy, info, value_first, dvalue_first);
}

void hanko_forward(hanko_model* h, const double* D0, const double* policy, int K,
                   const double* dpolicy, double* KD, double* dKD, double* D_path, double* dD_last) {
  forward(h->m, D0, policy, K, dpolicy, KD, dKD, D_path, dD_last);
}

void hanko_ks_residuals(int P, double alpha, double delta, double ss_start_KS, const double* x,
                        const double* KD, const double* Z, int K, const double* dx,
                        const double* dKD, double* F, double* dF) {
  ks_residuals(P, alpha, delta, ss_start_KS, x, KD, Z, K, dx, dKD, F, dF);
}

int hanko_ks_fjvp(hanko_model* h, double alpha, double delta, double ss_start_KS,
                  const double* value_T, const double* D0, const double* Z, const double* x, int K,
                  const double* V, double* F, double* JV, int* info) {
  KSProblem pb{&h->m, alpha, delta, ss_start_KS, value_T, D0, Z};
  return full_function(pb, x, K, V, F, JV, info);
}

int hanko_gmres(const double* A, int n, double* x, const double* b, int* iters) {
  return gmres(A, n, x, b, iters);
}

// NewtonRaphsonHANK + y_Iteration — NewtonRaphson.jl:27-46, :65-114.
// solver: 0 = reference (two restarted-GMRES solves per inner step, the M solve is dead work but
// executed), 1 = dense LU of Jbar (factor once). stats: [outer, total_jvps, total_F, last ||y||,
// gmres_iters_total]; inner_counts (optional, up to 100 ints).
int hanko_newton(hanko_model* h, double alpha, double delta, double ss_start_KS,
                 const double* value_T, const double* D0, const double* Z, const double* Jbar,
                 const double* x0, double eps, double eps_inner, int solver, int max_inner,
                 double* x_out,
                 double* stats, int* inner_counts, int* info) {
  KSProblem pb{&h->m, alpha, delta, ss_start_KS, value_T, D0, Z};
  const int n = 4 * h->m.P();
  std::vector<double> x(x0, x0 + n), y(x0, x0 + n);
  std::vector<double> Fx(n), Lxy(n), R(n), M(n), yold(n), rhs(n), dummyF(n);
  std::vector<double> LU; std::vector<int> piv;
  if (solver == 1) { LU.assign(Jbar, Jbar + (size_t)n * n); if (lu_factor(LU, n, piv)) return ERR_ARG; }
  int outer = 1, jvps = 0, fevals = 0; long gm = 0;
  auto norm_diff = [&](const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0; for (int i = 0; i < n; ++i) { double d = a[i] - b[i]; s += d * d; } return std::sqrt(s); };
  while (eps < nrm2(y.data(), n) && outer < 100) {
    // y_Iteration(J̅, x, y, ...)
    std::fill(yold.begin(), yold.end(), 1.0);
    std::fill(M.begin(), M.end(), 1.0);
    std::fill(R.begin(), R.end(), 1.0);
    int rc = full_function(pb, x.data(), 0, nullptr, Fx.data(), nullptr, info);
    if (rc) return rc;
    ++fevals;
    int inner = 0;
    while (eps_inner < norm_diff(y, yold)) {
      rc = full_function(pb, x.data(), 1, y.data(), dummyF.data(), Lxy.data(), info);
      if (rc) return rc;
      ++jvps; ++inner;
      for (int i = 0; i < n; ++i) rhs[i] = Fx[i] - Lxy[i];
      if (solver == 0) {
        int it = 0;
        gmres(Jbar, n, R.data(), rhs.data(), &it); gm += it;
        gmres(Jbar, n, M.data(), Lxy.data(), &it); gm += it;
      } else {
        R = rhs; lu_solve(LU, n, piv, R.data());
      }
      yold = y;
      for (int i = 0; i < n; ++i) y[i] = yold[i] + 0.5 * R[i];
      if (max_inner > 0 && inner >= max_inner) break;
    }
    if (inner_counts && outer - 1 < 100) inner_counts[outer - 1] = inner;
    for (int i = 0; i < n; ++i) x[i] = x[i] - y[i];
    ++outer;
  }
  std::memcpy(x_out, x.data(), sizeof(double) * n);
  if (stats) { stats[0] = outer - 1; stats[1] = jvps; stats[2] = fevals; stats[3] = nrm2(y.data(), n); stats[4] = (double)gm; }
  return OK;
}

int hanko_lu_solve_dense(const double* A, int n, int nrhs, double* B) {
  std::vector<double> LU(A, A + (size_t)n * n); std::vector<int> piv;
  if (lu_factor(LU, n, piv)) return ERR_ARG;
  for (int j = 0; j < nrhs; ++j) lu_solve(LU, n, piv, B + (size_t)j * n);
  return OK;
}

double hanko_jl_pow(double x, double y) { return jl_pow(x, y, nullptr); }

}  // extern "C"
