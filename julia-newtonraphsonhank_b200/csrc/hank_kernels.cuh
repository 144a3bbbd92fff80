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

// HBM layout: every per-point array has the compile-time leading dimension LDA = NT*R >= n_a
// (256/512/1024/2048).  The tape is stored as ONE CONTIGUOUS CHUNK PER COLUMN (t, e) so that a
// single TMA bulk copy stages everything the tangent sweep needs for that column:
//   backward chunk: BW_NF x [LDA] doubles (a1, kr, cA, cB, E1, vf) | [LDA] ints (left knot idx)
//   forward  chunk: FW_NF x [LDA] doubles (ω, D/Δg, D_t, p_t)      | [LDA+4] ints (range starts)
// Policy tangents are [t][e][lane][LDA] so the L lanes of a CTA are contiguous per column.
enum { BW_A1 = 0, BW_KR, BW_CA, BW_CB, BW_E1, BW_VF, BW_NF };   // backward tape fields
enum { FW_OM = 0, FW_DCO, FW_D, FW_P, FW_NF };                  // forward tape fields
template <int LDA> __host__ __device__ constexpr size_t bw_chunk_bytes() { return (size_t)BW_NF * LDA * 8 + (size_t)LDA * 4; }
template <int LDA> __host__ __device__ constexpr size_t fw_chunk_bytes() { return (size_t)FW_NF * LDA * 8 + (size_t)(LDA + 4) * 4; }
struct Tape {
  double* pol;          // [P][NE][LDA]   policy a'(a,e)
  unsigned char* bw;    // [P][NE] backward chunks
  double* rho;          // [P]
  unsigned char* fw;    // [P][NE] forward chunks
  int* mbr;             // [P][NE][LDA]   1-based searchsortedfirst brackets
  double* value_first;  // [NE][LDA]      ∂V/∂a after the last backward step (t = 1)
};
template <int LDA> __device__ __forceinline__ double* bw_fields(const Tape& tp, int NE, int t, int e) {
  return reinterpret_cast<double*>(tp.bw + ((size_t)t * NE + e) * bw_chunk_bytes<LDA>());
}
template <int LDA> __device__ __forceinline__ int* bw_idx(const Tape& tp, int NE, int t, int e) {
  return reinterpret_cast<int*>(tp.bw + ((size_t)t * NE + e) * bw_chunk_bytes<LDA>() + (size_t)BW_NF * LDA * 8);
}
template <int LDA> __device__ __forceinline__ double* fw_fields(const Tape& tp, int NE, int t, int e) {
  return reinterpret_cast<double*>(tp.fw + ((size_t)t * NE + e) * fw_chunk_bytes<LDA>());
}
template <int LDA> __device__ __forceinline__ int* fw_start(const Tape& tp, int NE, int t, int e) {
  return reinterpret_cast<int*>(tp.fw + ((size_t)t * NE + e) * fw_chunk_bytes<LDA>() + (size_t)FW_NF * LDA * 8);
}

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
// (cg^-γ, cg^(-γ-1)) with one division
template <bool G2>
__device__ __forceinline__ void pow_v2(double cg, double gamma, double& cgp, double& cgp1) {
  const double rx = 1.0 / cg;
  cgp = G2 ? rx * rx : pow(cg, -gamma);
  cgp1 = cgp * rx;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Seed-horizon groups: lane counts per CTA (1, 2, 3, 4, 6) all divide 12, so a horizon per 12 lanes is
// the same for the backward and the forward kernel whatever shapes they run in.
constexpr int kThiGroup = 12;
constexpr int kZeroBytes = 32768;

// Sums of L per-thread values over the warp with L-1 + 5-log2(L) shuffles instead of 5 L: each
// halving step keeps half of the values and hands the other half to the partner lane.  The total
// of value l ends up in every lane whose top log2(L) lane-index bits spell l (lane (l*32)/L first).
template <int L>
__device__ __forceinline__ double warp_sum_multi(double (&v)[L], int lane) {
  static_assert(L == 1 || L == 2 || L == 4, "power-of-two lane counts only");
  int o = 16;
#pragma unroll
  for (int h = L / 2; h >= 1; h >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = up ? v[i] : v[i + h];
      const double keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  double r = v[0];
#pragma unroll
  for (; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
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
// Same result, branch-free with a fixed trip count (LDA >= n, power of two): independent searches
// interleave instead of serialising on the data-dependent loop.
template <int LDA>
__device__ __forceinline__ int lower_bound_fixed(const double* __restrict__ v, int n, double x) {
  int pos = 0;  // invariant: the first `pos` elements are < x
#pragma unroll
  for (int step = LDA; step >= 1; step >>= 1) {
    const int np = pos + step;
    if (np <= n && v[np - 1] < x) pos = np;
  }
  return pos;
}

}  // namespace hank
#include "hank_point.cuh"
namespace hank {

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
  double rn = rpath[P - 1], wn = wpath[P - 1];
  for (int t = P - 1; t >= 0; --t) {
    const double r = rn, w = wn;
    if (t > 0) { rn = rpath[t - 1]; wn = wpath[t - 1]; }
    const double opr = 1.0 + r, rho = 1.0 / opr;
    double* polt = tp.pol + (size_t)t * NE * LDA + tid;
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
          const EulerPoint u = egm_euler_point<G2>(M, ev, w * M.z[e], ga, rho, opr, status, a, e, t);
          ks[e * LDA + a] = u.knot;
          bw_fields<LDA>(tp, NE, t, e)[BW_A1 * LDA + j * NT + tid] = u.a1;
          bw_fields<LDA>(tp, NE, t, e)[BW_KR * LDA + j * NT + tid] = u.kr;
        }
      }
    }
    __syncthreads();
    // ---- phase 2: interpolate a'(·,e) on the exogenous grid, clamp, marginal value (:65-80)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int a = tid + j * NT;
      if (a < n_a) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const InterpPoint q = egm_interp_point<G2, LDA>(M, ks + e * LDA, g, n_a, a, w * M.z[e], opr, status, e, t);
          polt[e * LDA + j * NT] = q.p;
          bw_idx<LDA>(tp, NE, t, e)[j * NT + tid] = q.i;
          double* f = bw_fields<LDA>(tp, NE, t, e) + j * NT + tid;
          f[BW_CA * LDA] = q.cA;
          f[BW_CB * LDA] = q.cB;
          f[BW_VF * LDA] = q.vf;
          f[BW_E1 * LDA] = q.E1;
          V[j][e] = q.vnew;
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
  // the period's policies are requested one period ahead when they fit in registers
  constexpr bool PRE = R * NE <= 14;
  double pc[PRE ? R : 1][PRE ? NE : 1], pn[PRE ? R : 1][PRE ? NE : 1];
  if (PRE) {
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int e = 0; e < NE; ++e) pc[PRE ? j : 0][PRE ? e : 0] = tid + j * NT < n_a ? pol_in[e * LDA + j * NT + tid] : 0.0;
  }
  for (int t = 0; t < P; ++t) {
    const double* polt = pol_in + (size_t)t * GP + tid;
    int* mbt = tp.mbr + (size_t)t * GP + tid;
    if (PRE && t + 1 < P) {
#pragma unroll
      for (int j = 0; j < R; ++j)
#pragma unroll
        for (int e = 0; e < NE; ++e) pn[PRE ? j : 0][PRE ? e : 0] = tid + j * NT < n_a ? polt[GP + e * LDA + j * NT] : 0.0;
    }
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
              const double p = PRE ? pc[PRE ? j : 0][PRE ? e : 0] : polt[e * LDA + j * NT];
              const LotteryPoint q = lottery_point<LDA>(g, n_a, p, D[j][e]);   // m > n_a: all mass to node n_a via the (1-ω) leg
              X[ee * LDA + a] = q.om * D[j][e];
              Y[ee * LDA + a] = (1.0 - q.om) * D[j][e];
              ms[ee * LDA + a] = q.m;
              fw_fields<LDA>(tp, NE, t, e)[FW_OM * LDA + j * NT + tid] = q.om;
              fw_fields<LDA>(tp, NE, t, e)[FW_DCO * LDA + j * NT + tid] = q.dco;
              fw_fields<LDA>(tp, NE, t, e)[FW_P * LDA + j * NT + tid] = p;
              mbt[e * LDA + j * NT] = q.m;
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
              lottery_starts_point(ms + ee * LDA, st + ee * NS, n_a, a, status, e, t);
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
              int* so = fw_start<LDA>(tp, NE, t, e);
              so[a + 1] = s0;
              if (a == n_a - 1) { so[a + 2] = s1; so[a + 3] = s2; }
              tmp[j][e] = lottery_gather_point(X + ee * LDA, Y + ee * LDA, s0, s1, s2);
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
          fw_fields<LDA>(tp, NE, t, e2)[FW_D * LDA + j * NT + tid] = d;
          kacc += (PRE ? pc[PRE ? j : 0][PRE ? e2 : 0] : polt[e2 * LDA + j * NT]) * d;
        }
      }
    }
    kacc = warp_sum(kacc);
    if (lane == 0) kdpart[(size_t)t * NW + warp] = kacc;
    if (PRE) {
#pragma unroll
      for (int j = 0; j < R; ++j)
#pragma unroll
        for (int e = 0; e < NE; ++e) pc[PRE ? j : 0][PRE ? e : 0] = pn[PRE ? j : 0][PRE ? e : 0];
    }
  }
  __syncthreads();
  for (int t = tid; t < P; t += NT) {
    double s = 0.0;
    for (int wv = 0; wv < NW; ++wv) s += kdpart[(size_t)t * NW + wv];
    KD[t] = s;
  }
}

}  // namespace hank
