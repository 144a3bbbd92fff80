// hank_kernels.cuh — hand-written sm_100a kernels of the household block.
//
// Design: the two time recursions are strictly sequential, so every sweep is ONE
// persistent kernel launch that walks all P periods with __syncthreads() between phases; the
// recursion state (∂V/∂a or D, and their tangent lanes) lives in registers, the only shared
// memory traffic is the staging of the endogenous knots / lottery masses that couple threads
// along the asset dimension.  The primal sweep runs once per linearisation point in a single
// CTA and records a "tape" of per-point coefficients; the tangent sweeps are linear recursions
// with those coefficients, one CTA per group of L lanes, all groups in flight at once.
//
// Thread mapping: thread `tid` owns asset rows a = tid + j*NT (j < R) for ALL income states, so
// the Markov mix over e (KrusellSmith.jl:59, ForwardIteration.jl:98) is thread-local in
// registers with Π as immediate kernel-parameter constants.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hank {

constexpr int kMaxNE = 16;

template <int NE>
struct Consts {          // passed by value: lives in the constant bank, indexed at compile time
  double Pi[NE][NE];     // Pi[e][e2] = Π[e,e2] (row-stochastic)
  double z[NE];
  double beta, gamma, bc, yexp;  // yexp = -1/γ
  int n_a, P, gamma_int;         // gamma_int: γ is integer-valued (Julia's integer-power path)
};

// HBM layout: every per-point array is [t][e][LDA] with the compile-time leading dimension
// LDA = NT*R >= n_a (256/512/1024/2048), so one base pointer per period plus immediate offsets
// addresses every (field, e, row) a thread touches — no per-column address registers.
enum { BW_A1 = 0, BW_KR, BW_CA, BW_CB, BW_E1, BW_VF, BW_NF };   // backward tape fields
enum { FW_OM = 0, FW_DCO, FW_D, FW_NF };                        // forward tape fields
struct Tape {
  double* pol;          // [P][NE][LDA]      policy a'(a,e)
  double* bw;           // [P][BW_NF][NE][LDA] EGM linearisation
  int* idx;             // [P][NE][LDA]      left knot of the interpolation interval (0-based)
  double* rho;          // [P]
  double* fw;           // [P][FW_NF][NE][LDA] lottery linearisation: ω, D/Δg, D_t
  int* start;           // [P][NE][LDA+4]    source-range starts per destination row
  int* mbr;             // [P][NE][LDA]      1-based searchsortedfirst brackets
  double* value_first;  // [NE][LDA]         ∂V/∂a after the last backward step (t = 1)
};

enum { ST_CODE = 0, ST_A = 1, ST_E = 2, ST_T = 3 };

__device__ __forceinline__ void raise(int* status, int code, int a, int e, int t) {
  if (atomicCAS(&status[ST_CODE], 0, code) == 0) {
    status[ST_A] = a + 1; status[ST_E] = e + 1; status[ST_T] = t + 1;
  }
}

// x^(-1/2) to <1 ulp: IEEE sqrt+div then one Newton step with an exact residual.
__device__ __forceinline__ double pow_neg_half(double x) {
  double y = 1.0 / sqrt(x);
  double th = x * y, tl = fma(x, y, -th);
  double e = fma(-th, y, 1.0) - tl * y;
  return fma(0.5 * y, e, y);
}

// c = B^(-1/γ) and cgp = cg^(-γ). γ == 2 takes the exact forms Julia's `^` reduces to for the
// second (inv(x)^2, base/math.jl pow_body) and a <1ulp form for the first.
template <bool G2>
__device__ __forceinline__ double pow_c(double B, double yexp) {
  return G2 ? pow_neg_half(B) : pow(B, yexp);
}
template <bool G2>
__device__ __forceinline__ double pow_v(double cg, double gamma) {
  if (G2) { double rx = 1.0 / cg; return rx * rx; }
  return pow(cg, -gamma);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// searchsortedfirst on a sorted shared/global array: number of elements < x (0-based result).
__device__ __forceinline__ int lower_bound(const double* __restrict__ v, int n, double x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (v[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ======================================================================================
// Backward primal sweep: EGM steps t = P..1 (BackwardIteration.jl:90-113 calling
// KrusellSmith.jl:43-83). One CTA. Writes the policy and the tape.
// smem: ks[NE][LDA] knots | g[LDA]
// ======================================================================================
template <int NE, int R, int NT, bool G2>
__global__ void __launch_bounds__(NT, 1)
k_backward_primal(const Consts<NE> M, const Tape tp, const double* __restrict__ grid,
                  const double* __restrict__ valueT, const double* __restrict__ rpath,
                  const double* __restrict__ wpath, int* __restrict__ status) {
  constexpr int LDA = NT * R;
  extern __shared__ double smem[];
  const int n_a = M.n_a, P = M.P;
  double* ks = smem;
  double* g = smem + (size_t)NE * LDA;
  const int tid = threadIdx.x;
  for (int a = tid; a < n_a; a += NT) g[a] = grid[a];
  double V[R][NE];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int a = tid + j * NT;
#pragma unroll
    for (int e = 0; e < NE; ++e) V[j][e] = a < n_a ? valueT[e * LDA + a] : 1.0;
  }
  __syncthreads();
  for (int t = P - 1; t >= 0; --t) {
    const double r = rpath[t], w = wpath[t];
    const double opr = 1.0 + r, rho = 1.0 / opr;
    double* bwt = tp.bw + (size_t)t * BW_NF * NE * LDA + tid;
    double* polt = tp.pol + (size_t)t * NE * LDA + tid;
    int* idxt = tp.idx + (size_t)t * NE * LDA + tid;
    if (tid == 0) tp.rho[t] = rho;
    // ---- phase 1: Euler inversion, endogenous grid (KrusellSmith.jl:59-62)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const double ga = g[a];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          double ev = 0.0;
#pragma unroll
          for (int e2 = 0; e2 < NE; ++e2) ev += V[j][e2] * M.Pi[e][e2];
          const double B = M.beta * ev;
          if (B < 0.0) raise(status, 2, a, e, t);
          const double c = pow_c<G2>(B, M.yexp);
          const double S = (c - w * M.z[e]) + ga;
          ks[e * LDA + a] = rho * S;
          // ċ = Ḃ·yexp·B^(yexp-1);  k̇ = ρ·(ċ − ẇ z) + S·ρ̇,  ρ̇ = −(ρ/(1+r))·ṙ
          bwt[(BW_A1 * NE + e) * LDA + j * NT] = rho * (M.beta * (M.yexp * (c / B)));
          bwt[(BW_KR * NE + e) * LDA + j * NT] = -(S * (rho / opr));
        }
      }
    }
    __syncthreads();
    // ---- phase 2: interpolate a'(·,e) on the exogenous grid, clamp, marginal value (:65-80)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
        const double x = g[a];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const double* k = ks + e * LDA;
          if (a > 0 && !(k[a] > k[a - 1])) raise(status, 3, a, e, t);
          int i; double num, den; bool interior = true;
          const double k0 = k[0], kl = k[n_a - 1];
          if (x > kl) { i = n_a - 2; den = kl - k[i]; num = den; interior = false; }
          else if (x < k0) { i = 0; den = k[1] - k0; num = 0.0; interior = false; }
          else {
            int lb = lower_bound(k, n_a, x);            // searchsortedfirst - 1 (0-based count)
            i = min(max(lb, 1), n_a - 1) - 1;           // find_knot_index clamp, left knot 0-based
            num = x - k[i]; den = k[i + 1] - k[i];
          }
          const double delta = num / den;
          const double gi = g[i], gi1 = g[i + 1];
          const double q = (1.0 - delta) * gi + delta * gi1;
          const bool cons = q < M.bc;
          const double p = cons ? M.bc : q;
          const double cg = (opr * x + w * M.z[e]) - p;
          if (cg < 0.0 && !M.gamma_int) raise(status, 2, a, e, t);
          const double cgp = pow_v<G2>(cg, M.gamma);
          polt[e * LDA + j * NT] = p;
          idxt[e * LDA + j * NT] = i;
          // δ̇ = (1/den)(−k̇_i) + (−num/den²)(k̇_{i+1} − k̇_i);  q̇ = δ̇ (g_{i+1} − g_i)
          const bool live = interior && !cons;
          const double dg = gi1 - gi, id = 1.0 / den, nd2 = num / (den * den);
          bwt[(BW_CA * NE + e) * LDA + j * NT] = live ? (nd2 - id) * dg : 0.0;
          bwt[(BW_CB * NE + e) * LDA + j * NT] = live ? -(nd2 * dg) : 0.0;
          // V̇ = ṙ·cg^-γ + (1+r)·(−γ)·cg^(−γ−1)·ċg,  ċg = ṙ a + ẇ z − ṗ
          const double vf = opr * ((-M.gamma) * (cgp / cg));
          bwt[(BW_VF * NE + e) * LDA + j * NT] = vf;
          bwt[(BW_E1 * NE + e) * LDA + j * NT] = cgp + vf * x;
          V[j][e] = opr * cgp;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int a = tid + j * NT;
    if (a < n_a)
#pragma unroll
      for (int e = 0; e < NE; ++e) tp.value_first[e * LDA + a] = V[j][e];
  }
}

// ======================================================================================
// Backward tangent sweep: K lanes of ForwardDiff partials through the EGM recursion
// (SURVEY.md A.3), as a linear recursion with the taped coefficients. CTA b carries lanes
// [b*L, b*L+L). smem: kds[L][NE][LDA] (k̇ staging for the bracket gather).
// dr/dw: [K][P]; dvalT: [K][NE][LDA] or null (zero terminal tangents, BackwardIteration.jl:85);
// dpol: [K][P][NE][LDA].
// ======================================================================================
template <int NE, int R, int NT, int L>
__global__ void __launch_bounds__(NT, 1)
k_backward_tangent(const Consts<NE> M, const Tape tp, int K, const double* __restrict__ dr,
                   const double* __restrict__ dw, const double* __restrict__ dvalT,
                   double* __restrict__ dpol, double* __restrict__ dvalue_first) {
  constexpr int LDA = NT * R;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  const int n_a = M.n_a, P = M.P;
  const int tid = threadIdx.x;
  double* kds = smem + tid;
  const int lane0 = blockIdx.x * L;
  double Vd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
#pragma unroll
      for (int e = 0; e < NE; ++e)
        Vd[l][j][e] = (dvalT && a < n_a && lane0 + l < K) ? dvalT[(size_t)(lane0 + l) * GP + e * LDA + a] : 0.0;
    }
  for (int t = P - 1; t >= 0; --t) {
    const double* bwt = tp.bw + (size_t)t * BW_NF * GP + tid;
    const int* idxt = tp.idx + (size_t)t * GP + tid;
    const double rho = tp.rho[t];
    double drl[L], dwl[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const bool on = lane0 + l < K;
      drl[l] = on ? __ldg(dr + (size_t)(lane0 + l) * P + t) : 0.0;
      dwl[l] = on ? __ldg(dw + (size_t)(lane0 + l) * P + t) : 0.0;
    }
    // ---- phase 1: k̇ = a1·ĖV + kr·ṙ − ρ z ẇ
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (tid + j * NT < n_a) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const double a1 = __ldg(bwt + (BW_A1 * NE + e) * LDA + j * NT);
          const double kr = __ldg(bwt + (BW_KR * NE + e) * LDA + j * NT);
          const double cw = -(rho * M.z[e]);
#pragma unroll
          for (int l = 0; l < L; ++l) {
            double ev = 0.0;
#pragma unroll
            for (int e2 = 0; e2 < NE; ++e2) ev = fma(M.Pi[e][e2], Vd[l][j][e2], ev);
            kds[(l * NE + e) * LDA + j * NT] = fma(a1, ev, fma(kr, drl[l], cw * dwl[l]));
          }
        }
      }
    }
    __syncthreads();
    // ---- phase 2: q̇ from the two bracketing knots, ṗ, V̇
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (tid + j * NT < n_a) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const int i = __ldg(idxt + e * LDA + j * NT);
          const double cA = __ldg(bwt + (BW_CA * NE + e) * LDA + j * NT);
          const double cB = __ldg(bwt + (BW_CB * NE + e) * LDA + j * NT);
          const double E1 = __ldg(bwt + (BW_E1 * NE + e) * LDA + j * NT);
          const double vf = __ldg(bwt + (BW_VF * NE + e) * LDA + j * NT);
          const double ze = M.z[e];
          const double* kk = smem + i;
#pragma unroll
          for (int l = 0; l < L; ++l) {
            const double pd = fma(cA, kk[(l * NE + e) * LDA], cB * kk[(l * NE + e) * LDA + 1]);
            if (lane0 + l < K)
              __stcs(dpol + ((size_t)(lane0 + l) * P + t) * GP + e * LDA + j * NT + tid, pd);
            Vd[l][j][e] = fma(vf, fma(ze, dwl[l], -pd), E1 * drl[l]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (dvalue_first) {
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (lane0 + l < K)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int a = tid + j * NT;
          if (a < n_a)
#pragma unroll
            for (int e = 0; e < NE; ++e) dvalue_first[(size_t)(lane0 + l) * GP + e * LDA + a] = Vd[l][j][e];
        }
  }
}

// ======================================================================================
// Forward primal sweep: Young lottery + Markov mix + aggregation, t = 1..P
// (ForwardIteration.jl:37-99, :253-311). Gather formulation: the policy is monotone in a, so
// the sources of destination node `row` are two contiguous source ranges
// [start[row], start[row+1]) (weight ω, m == row) and [start[row+1], start[row+2]) (weight 1-ω,
// m == row+1), walked in ascending source order like Julia's CSC SpMV. No atomics.
// Columns are staged CS at a time (CS = NE when the whole grid fits in shared memory).
// smem: X[CS][LDA] | Y[CS][LDA] | g[LDA] | ms[CS][LDA] (int) | st[CS][LDA+4] (int)
// kdpart: [P][NT/32] per-warp partial sums of <p_t, D_t>.
// ======================================================================================
template <int NE, int R, int NT, int CS>
__global__ void __launch_bounds__(NT, 1)
k_forward_primal(const Consts<NE> M, const Tape tp, const double* __restrict__ grid,
                 const double* __restrict__ D0, const double* __restrict__ pol_in,
                 double* __restrict__ kdpart, double* __restrict__ KD, int* __restrict__ status) {
  constexpr int LDA = NT * R, NS = LDA + 4;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  const int n_a = M.n_a, P = M.P;
  double* X = smem;
  double* Y = X + (size_t)CS * LDA;
  double* g = Y + (size_t)CS * LDA;
  int* ms = reinterpret_cast<int*>(g + LDA);
  int* st = ms + (size_t)CS * LDA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = NT / 32;
  for (int a = tid; a < n_a; a += NT) g[a] = grid[a];
  for (int i = tid; i < CS * NS; i += NT) st[i] = n_a;  // safe ranges even for a rejected policy
  double D[R][NE];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int a = tid + j * NT;
#pragma unroll
    for (int e = 0; e < NE; ++e) D[j][e] = a < n_a ? D0[e * LDA + a] : 0.0;
  }
  __syncthreads();
  for (int t = 0; t < P; ++t) {
    const double* polt = pol_in + (size_t)t * GP + tid;
    double* fwt = tp.fw + (size_t)t * FW_NF * GP + tid;
    int* mbt = tp.mbr + (size_t)t * GP + tid;
    double tmp[R][NE];
#pragma unroll
    for (int e0 = 0; e0 < NE; e0 += CS) {
      // ---- phase A: brackets and lottery masses
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int a = tid + j * NT;
        if (a < n_a) {
#pragma unroll
          for (int ee = 0; ee < CS; ++ee) {
            const int e = e0 + ee;
            if (e < NE) {
              const double p = polt[e * LDA + j * NT];
              const int m = lower_bound(g, n_a, p) + 1;  // Julia searchsortedfirst, 1-based
              double om, dco;
              if (m == 1) { om = 1.0; dco = 0.0; }
              else if (m > n_a) { om = 0.0; dco = 0.0; }  // all mass to node n_a via the (1-ω) leg
              else {
                const double dgm = g[m - 1] - g[m - 2];
                om = (p - g[m - 2]) / dgm;
                dco = D[j][e] / dgm;
              }
              X[ee * LDA + a] = om * D[j][e];
              Y[ee * LDA + a] = (1.0 - om) * D[j][e];
              ms[ee * LDA + a] = m;
              fwt[(FW_OM * NE + e) * LDA + j * NT] = om;
              fwt[(FW_DCO * NE + e) * LDA + j * NT] = dco;
              mbt[e * LDA + j * NT] = m;
            }
          }
        }
      }
      __syncthreads();
      // ---- phase B: source-range starts per destination row (rows 1..n_a+2)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int a = tid + j * NT;
        if (a < n_a) {
#pragma unroll
          for (int ee = 0; ee < CS; ++ee) {
            const int e = e0 + ee;
            if (e < NE) {
              const int hi = ms[ee * LDA + a];
              const int lo = a == 0 ? 0 : ms[ee * LDA + a - 1];
              if (hi < lo) raise(status, 6, a, e, t);
              for (int row = lo + 1; row <= hi; ++row) st[ee * NS + row] = a;
              if (a == n_a - 1)
                for (int row = max(hi, lo) + 1; row <= n_a + 2; ++row) st[ee * NS + row] = n_a;
            }
          }
        }
      }
      __syncthreads();
      // ---- phase C: gather
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int a = tid + j * NT;
        if (a < n_a) {
#pragma unroll
          for (int ee = 0; ee < CS; ++ee) {
            const int e = e0 + ee;
            if (e < NE) {
              const int* s = st + ee * NS + a + 1;
              const int s0 = s[0], s1 = s[1], s2 = s[2];
              int* so = tp.start + ((size_t)t * NE + e) * NS;
              so[a + 1] = s0;
              if (a == n_a - 1) { so[a + 2] = s1; so[a + 3] = s2; }
              double acc = 0.0;
              for (int b = s0; b < s1; ++b) acc += X[ee * LDA + b];
              for (int b = s1; b < s2; ++b) acc += Y[ee * LDA + b];
              tmp[j][e] = acc;
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- Markov mix (Λ_exog = kron(Π', I), ForwardIteration.jl:280-284) and aggregation (:306)
    double kacc = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (tid + j * NT < n_a) {
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) {
          double d = 0.0;
#pragma unroll
          for (int e = 0; e < NE; ++e) d += M.Pi[e][e2] * tmp[j][e];
          D[j][e2] = d;
          fwt[(FW_D * NE + e2) * LDA + j * NT] = d;
          kacc += polt[e2 * LDA + j * NT] * d;
        }
      }
    }
    kacc = warp_sum(kacc);
    if (lane == 0) kdpart[(size_t)t * NW + warp] = kacc;
  }
  __syncthreads();
  for (int t = tid; t < P; t += NT) {
    double s = 0.0;
    for (int wv = 0; wv < NW; ++wv) s += kdpart[(size_t)t * NW + wv];
    KD[t] = s;
  }
}

// ======================================================================================
// Forward tangent sweep: Ḋ recursion and K̇D_t = <ṗ_t, D_t> + <p_t, Ḋ_t> for L lanes per CTA
// (SURVEY.md A.4 tangent rules).  smem: Xd[L][CS][LDA] | Yd[L][CS][LDA]
// dkdpart: [K][P][NT/32]
// ======================================================================================
template <int NE, int R, int NT, int L, int CS>
__global__ void __launch_bounds__(NT, 1)
k_forward_tangent(const Consts<NE> M, const Tape tp, int K, const double* __restrict__ pol_in,
                  const double* __restrict__ dpol, const double* __restrict__ dD0,
                  double* __restrict__ dkdpart, double* __restrict__ dD_last) {
  constexpr int LDA = NT * R, NS = LDA + 4;
  constexpr size_t GP = (size_t)NE * LDA;
  extern __shared__ double smem[];
  const int n_a = M.n_a, P = M.P;
  double* Xd = smem;
  double* Yd = smem + (size_t)L * CS * LDA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = NT / 32;
  const int lane0 = blockIdx.x * L;
  double Dd[L][R][NE];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
#pragma unroll
      for (int e = 0; e < NE; ++e)
        Dd[l][j][e] = (dD0 && a < n_a && lane0 + l < K) ? dD0[(size_t)(lane0 + l) * GP + e * LDA + a] : 0.0;
    }
  for (int t = 0; t < P; ++t) {
    const double* fwt = tp.fw + (size_t)t * FW_NF * GP + tid;
    const double* polt = pol_in + (size_t)t * GP + tid;
    double kacc[L];
#pragma unroll
    for (int l = 0; l < L; ++l) kacc[l] = 0.0;
    double tmp[L][R][NE];
#pragma unroll
    for (int e0 = 0; e0 < NE; e0 += CS) {
      // ---- phase A: ẋ = ω Ḋ + (D/Δg) ṗ,  ẏ = Ḋ − ẋ;  first aggregation term <ṗ_t, D_t>
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (tid + j * NT < n_a) {
#pragma unroll
          for (int ee = 0; ee < CS; ++ee) {
            const int e = e0 + ee;
            if (e < NE) {
              const double om = __ldg(fwt + (FW_OM * NE + e) * LDA + j * NT);
              const double dco = __ldg(fwt + (FW_DCO * NE + e) * LDA + j * NT);
              const double Dn = __ldg(fwt + (FW_D * NE + e) * LDA + j * NT);
#pragma unroll
              for (int l = 0; l < L; ++l) {
                const double pd = (lane0 + l < K)
                    ? __ldcs(dpol + ((size_t)(lane0 + l) * P + t) * GP + e * LDA + j * NT + tid) : 0.0;
                const double xd = fma(om, Dd[l][j][e], dco * pd);
                Xd[(l * CS + ee) * LDA + j * NT + tid] = xd;
                Yd[(l * CS + ee) * LDA + j * NT + tid] = Dd[l][j][e] - xd;
                kacc[l] = fma(pd, Dn, kacc[l]);
              }
            }
          }
        }
      }
      __syncthreads();
      // ---- phase C: gather over the taped source ranges
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int a = tid + j * NT;
        if (a < n_a) {
#pragma unroll
          for (int ee = 0; ee < CS; ++ee) {
            const int e = e0 + ee;
            if (e < NE) {
              const int* s = tp.start + ((size_t)t * NE + e) * NS + a + 1;
              const int s0 = __ldg(s), s1 = __ldg(s + 1), s2 = __ldg(s + 2);
#pragma unroll
              for (int l = 0; l < L; ++l) {
                const double* xs = Xd + (l * CS + ee) * LDA;
                const double* ys = Yd + (l * CS + ee) * LDA;
                double acc = 0.0;
                for (int b = s0; b < s1; ++b) acc += xs[b];
                for (int b = s1; b < s2; ++b) acc += ys[b];
                tmp[l][j][e] = acc;
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- Markov mix and second aggregation term <p_t, Ḋ_t>
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (tid + j * NT < n_a) {
#pragma unroll
        for (int e2 = 0; e2 < NE; ++e2) {
          const double p = __ldg(polt + e2 * LDA + j * NT);
#pragma unroll
          for (int l = 0; l < L; ++l) {
            double d = 0.0;
#pragma unroll
            for (int e = 0; e < NE; ++e) d = fma(M.Pi[e][e2], tmp[l][j][e], d);
            Dd[l][j][e2] = d;
            kacc[l] = fma(p, d, kacc[l]);
          }
        }
      }
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const double s = warp_sum(kacc[l]);
      if (lane == 0 && lane0 + l < K) dkdpart[((size_t)(lane0 + l) * P + t) * NW + warp] = s;
    }
  }
  if (dD_last) {
#pragma unroll
    for (int l = 0; l < L; ++l)
      if (lane0 + l < K)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int a = tid + j * NT;
          if (a < n_a)
#pragma unroll
            for (int e = 0; e < NE; ++e) dD_last[(size_t)(lane0 + l) * GP + e * LDA + a] = Dd[l][j][e];
        }
  }
}

}  // namespace hank
