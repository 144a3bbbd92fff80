// hank_eq.cu — the model's equilibrium equations as DATA: a postfix program per equation, interpreted on the device.
//
// Replaces, for models other than the built-in Krusell-Smith aggregate block, the reference's compile_residuals
// (ModelParser.jl:217-259: YAML equation strings -> a Julia function evaluating LHS .- RHS over the columns of the
// padded variable matrix) together with assemble_full_xMat (GeneralStructures.jl:329-377: rows = var_names(model) =
// endogenous, heterogeneous, exogenous; `max_lag` steady-state columns in front, `max_lead` behind) and the dual-number
// pass ForwardDiff pushes through both (GeneralStructures.jl:542-550).  The host (hankb200/equations.py, or the Julia
// glue) compiles the strings once; hank_eq_configure uploads the program; after that hank_ks_linearize / hank_ks_jvp /
// hank_ks_jacobian_columns / hank_newton_solve evaluate F(x), J(x)V and J(x) for an n_endog x P system whose
// household block is still the EGM / lottery sweeps (inputs r and w, output KD).
//
// One thread per (period, equation) — or (lane, period, equation) for the tangents — runs the program on a small
// evaluation stack; variables are read straight from x / KD / the exogenous paths with the boundary rule applied per
// access, so the padded matrix is never materialised.  The tangent rules are ForwardDiff's (DiffRules 1.15):
//   d(a*b) = ȧ b + a ḃ,  d(a/b) = (ȧ - (a/b) ḃ)/b,  d(a^b) = ȧ·(b·a^(b-1)) [+ ḃ·a^b·log a only if ḃ != 0],
//   d exp = ȧ·exp,  d log = ȧ/a,  d sqrt = ȧ/(2 sqrt a).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include "hank_ctx.h"
#include "../../include/hankb200.h"

namespace hank {

#define CK(call)                                                   \
  do {                                                             \
    int rc__ = hank::cuda_check(c, (call), #call);                 \
    if (rc__) return rc__;                                         \
  } while (0)

enum { OP_CONST = 0, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG, OP_EXP, OP_LOG, OP_SQRT, OP_COUNT };
constexpr int kEqStack = 16;

struct EqProgram {
  int n_endog, n_exog, n_eq, P;
  const int* off;        // [n_eq + 1]
  const int* code;
  const double* consts;
  const double* ss;      // [2][nv]: boundary values before period 1 / after period P
};

// value of variable v at period index tt (0-based; outside 0..P-1: the steady-state boundary columns)
__device__ __forceinline__ double eq_var(const EqProgram& E, int v, int tt, const double* __restrict__ x,
                                         const double* __restrict__ KD, const double* __restrict__ Z) {
  const int nv = E.n_endog + 1 + E.n_exog;
  if (tt < 0) return E.ss[v];
  if (tt >= E.P) return E.ss[nv + v];
  if (v < E.n_endog) return x[(size_t)tt * E.n_endog + v];
  if (v == E.n_endog) return KD[tt];
  return Z[(size_t)(v - E.n_endog - 1) * E.P + tt];
}
// its tangent in the direction (V, dKD): boundaries and exogenous paths are constants
__device__ __forceinline__ double eq_dvar(const EqProgram& E, int v, int tt, const double* __restrict__ V,
                                          const double* __restrict__ dKD) {
  if (tt < 0 || tt >= E.P) return 0.0;
  if (v < E.n_endog) return V ? V[(size_t)tt * E.n_endog + v] : 0.0;
  if (v == E.n_endog) return dKD ? dKD[tt] : 0.0;
  return 0.0;
}

// F[t*n_eq + i] = LHS_i - RHS_i at period t
__global__ void k_eq_residual(const EqProgram E, const double* __restrict__ x, const double* __restrict__ KD,
                              const double* __restrict__ Z, double* __restrict__ F) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E.P * E.n_eq) return;
  const int t = idx / E.n_eq, i = idx - t * E.n_eq;
  double st[kEqStack];
  int sp = 0;
  for (int pc = E.off[i]; pc < E.off[i + 1];) {
    const int op = E.code[pc++];
    if (op == OP_CONST) { st[sp++] = E.consts[E.code[pc++]]; continue; }
    if (op == OP_VAR) { const int v = E.code[pc], k = E.code[pc + 1]; pc += 2; st[sp++] = eq_var(E, v, t + k, x, KD, Z); continue; }
    if (op >= OP_NEG) {
      const double a = st[sp - 1];
      st[sp - 1] = op == OP_NEG ? -a : op == OP_EXP ? exp(a) : op == OP_LOG ? log(a) : sqrt(a);
      continue;
    }
    const double b = st[--sp], a = st[sp - 1];
    st[sp - 1] = op == OP_ADD ? a + b : op == OP_SUB ? a - b : op == OP_MUL ? a * b : op == OP_DIV ? a / b : pow(a, b);
  }
  F[idx] = st[0];
}

// JV[l*n + t*n_eq + i] = d/dε F_i,t(x + ε V_l, KD + ε dKD_l);  V, JV: [K][n];  dKD: [K][P] or null (no household term)
__global__ void k_eq_residual_tangent(const EqProgram E, int K, const double* __restrict__ x, const double* __restrict__ KD,
                                      const double* __restrict__ Z, const double* __restrict__ V,
                                      const int* __restrict__ unit_cols, const double* __restrict__ dKD,
                                      const int* __restrict__ col_lane, double* __restrict__ JV) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)E.P * E.n_eq;
  if (idx >= n * K) return;
  const int l = (int)(idx / n), rem = (int)(idx - (size_t)l * n), t = rem / E.n_eq, i = rem - t * E.n_eq;
  const double* Vl = V ? V + (size_t)l * E.P * E.n_endog : nullptr;
  const int ucol = unit_cols ? unit_cols[l] : -1;   // Jacobian columns: the seed is the unit vector e_ucol
  const int dl = col_lane ? col_lane[l] : l;
  const double* dKl = (dKD && dl >= 0) ? dKD + (size_t)dl * E.P : nullptr;
  double sv[kEqStack], sd[kEqStack];
  int sp = 0;
  for (int pc = E.off[i]; pc < E.off[i + 1];) {
    const int op = E.code[pc++];
    if (op == OP_CONST) { sv[sp] = E.consts[E.code[pc++]]; sd[sp++] = 0.0; continue; }
    if (op == OP_VAR) {
      const int v = E.code[pc], k = E.code[pc + 1]; pc += 2;
      sv[sp] = eq_var(E, v, t + k, x, KD, Z);
      sd[sp++] = Vl ? eq_dvar(E, v, t + k, Vl, dKl)
                    : (v < E.n_endog ? ((t + k >= 0 && t + k < E.P && (t + k) * E.n_endog + v == ucol) ? 1.0 : 0.0) : eq_dvar(E, v, t + k, nullptr, dKl));
      continue;
    }
    if (op >= OP_NEG) {
      const double a = sv[sp - 1], da = sd[sp - 1];
      if (op == OP_NEG) { sv[sp - 1] = -a; sd[sp - 1] = -da; }
      else if (op == OP_EXP) { const double e = exp(a); sv[sp - 1] = e; sd[sp - 1] = da * e; }
      else if (op == OP_LOG) { sv[sp - 1] = log(a); sd[sp - 1] = da / a; }
      else { const double s = sqrt(a); sv[sp - 1] = s; sd[sp - 1] = da / (2.0 * s); }
      continue;
    }
    --sp;
    const double b = sv[sp], db = sd[sp], a = sv[sp - 1], da = sd[sp - 1];
    double v, d;
    if (op == OP_ADD) { v = a + b; d = da + db; }
    else if (op == OP_SUB) { v = a - b; d = da - db; }
    else if (op == OP_MUL) { v = a * b; d = da * b + a * db; }
    else if (op == OP_DIV) { v = a / b; d = (da - v * db) / b; }
    else {
      v = pow(a, b);
      d = da * (b * pow(a, b - 1.0));
      if (db != 0.0) d += db * (v * log(a));
    }
    sv[sp - 1] = v; sd[sp - 1] = d;
  }
  JV[idx] = sd[0];
}

__global__ void k_eq_extract_rw(const double* __restrict__ x, int P, int ne, int ir, int iw, double* r, double* w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P) { r[t] = x[(size_t)t * ne + ir]; w[t] = x[(size_t)t * ne + iw]; }
}
__global__ void k_eq_extract_drdw(const double* __restrict__ V, int P, int K, int ne, int ir, int iw, double* dr, double* dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * P) {
    const int l = i / P, t = i - l * P;
    const double* v = V + (size_t)l * ne * P;
    dr[i] = v[(size_t)t * ne + ir]; dw[i] = v[(size_t)t * ne + iw];
  }
}
static EqProgram eq_program(const hank_ctx* c) {
  EqProgram E;
  E.n_endog = c->n_endog; E.n_exog = c->n_exog; E.n_eq = c->n_endog; E.P = c->P;
  E.off = c->d_eq_off; E.code = c->d_eq_code; E.consts = c->d_eq_consts; E.ss = c->d_eq_ss;
  return E;
}
static inline unsigned eq_blocks(size_t n) { return (unsigned)((n + 255) / 256); }

int eq_extract_rw(hank_ctx* c, const double* x, double* r, double* w) {
  k_eq_extract_rw<<<eq_blocks(c->P), 256, 0, c->stream>>>(x, c->P, c->n_endog, c->eq_ir, c->eq_iw, r, w);
  c->launches++;
  return cuda_check(c, cudaGetLastError(), "k_eq_extract_rw");
}
int eq_extract_drdw(hank_ctx* c, int K, const double* V, double* dr, double* dw) {
  k_eq_extract_drdw<<<eq_blocks((size_t)K * c->P), 256, 0, c->stream>>>(V, c->P, K, c->n_endog, c->eq_ir, c->eq_iw, dr, dw);
  c->launches++;
  return cuda_check(c, cudaGetLastError(), "k_eq_extract_drdw");
}
int eq_residual(hank_ctx* c, const double* x, const double* KD, const double* Z, double* F) {
  k_eq_residual<<<eq_blocks((size_t)c->P * c->n_endog), 256, 0, c->stream>>>(eq_program(c), x, KD, Z, F);
  c->launches++;
  return cuda_check(c, cudaGetLastError(), "k_eq_residual");
}
int eq_residual_tangent(hank_ctx* c, int K, const double* x, const double* KD, const double* Z, const double* V,
                        const int* unit_cols, const double* dKD, const int* col_lane, double* JV) {
  k_eq_residual_tangent<<<eq_blocks((size_t)K * c->P * c->n_endog), 256, 0, c->stream>>>(eq_program(c), K, x, KD, Z, V, unit_cols, dKD,
                                                                                      col_lane, JV);
  c->launches++;
  return cuda_check(c, cudaGetLastError(), "k_eq_residual_tangent");
}
void eq_release(hank_ctx* c) {
  if (c->d_eq_off) cudaFree(c->d_eq_off);
  if (c->d_eq_code) cudaFree(c->d_eq_code);
  if (c->d_eq_consts) cudaFree(c->d_eq_consts);
  if (c->d_eq_ss) cudaFree(c->d_eq_ss);
  c->d_eq_off = c->d_eq_code = nullptr; c->d_eq_consts = c->d_eq_ss = nullptr;
  c->eq_on = false;
}

}  // namespace hank

using namespace hank;

extern "C" int hank_eq_configure(hank_ctx* c, int n_endog, int n_exog, int ir, int iw, const int* eq_off, const int* code,
                                 int n_const, const double* consts, const double* ss_start, const double* ss_end) {
  if (!c) return HANK_ERR_ARG;
  CK(cudaSetDevice(c->device));
  if (n_endog < 2 || n_endog > 64 || n_exog < 0 || n_exog > 64 || ir < 0 || iw < 0 || ir >= n_endog || iw >= n_endog || ir == iw ||
      !eq_off || !code || n_const < 0 || (n_const > 0 && !consts) || !ss_start || !ss_end)
    return set_error(c, HANK_ERR_ARG, "hank_eq_configure: bad arguments");
  const int nv = n_endog + 1 + n_exog, n_code = eq_off[n_endog];
  // validate the programs on the host: opcodes, operands, stack discipline (a bad program must not reach the device)
  for (int i = 0; i < n_endog; ++i) {
    if (eq_off[i] < 0 || eq_off[i + 1] < eq_off[i]) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: eq_off must be non-decreasing");
    int sp = 0;
    for (int pc = eq_off[i]; pc < eq_off[i + 1];) {
      const int op = code[pc++];
      if (op < 0 || op >= OP_COUNT) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: unknown opcode in equation " + std::to_string(i + 1));
      if (op == OP_CONST) { if (pc >= eq_off[i + 1] || code[pc] < 0 || code[pc] >= n_const) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: constant index out of range"); ++pc; ++sp; }
      else if (op == OP_VAR) {
        if (pc + 1 >= eq_off[i + 1]) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: truncated variable reference");
        if (code[pc] < 0 || code[pc] >= nv || std::abs(code[pc + 1]) >= c->P) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: variable index or shift out of range");
        pc += 2; ++sp;
      } else if (op >= OP_NEG) { if (sp < 1) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: stack underflow"); }
      else { if (sp < 2) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: stack underflow"); --sp; }
      if (sp > kEqStack) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: equation " + std::to_string(i + 1) + " needs a deeper stack than 16");
    }
    if (sp != 1) return set_error(c, HANK_ERR_ARG, "hank_eq_configure: equation " + std::to_string(i + 1) + " does not reduce to one value");
  }
  CK(cudaStreamSynchronize(c->stream));
  eq_release(c);
  std::vector<double> ss(2 * (size_t)nv);
  std::copy(ss_start, ss_start + nv, ss.begin());
  std::copy(ss_end, ss_end + nv, ss.begin() + nv);
  CK(cudaMalloc((void**)&c->d_eq_off, (n_endog + 1) * sizeof(int)));
  CK(cudaMalloc((void**)&c->d_eq_code, std::max(n_code, 1) * sizeof(int)));
  CK(cudaMalloc((void**)&c->d_eq_consts, std::max(n_const, 1) * sizeof(double)));
  CK(cudaMalloc((void**)&c->d_eq_ss, ss.size() * sizeof(double)));
  CK(cudaMemcpy(c->d_eq_off, eq_off, (n_endog + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_eq_code, code, n_code * sizeof(int), cudaMemcpyHostToDevice));
  if (n_const) CK(cudaMemcpy(c->d_eq_consts, consts, n_const * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_eq_ss, ss.data(), ss.size() * sizeof(double), cudaMemcpyHostToDevice));
  // the system is now n_endog x P with n_exog exogenous paths: resize what hank_ctx_create sized for Krusell-Smith
  const size_t n = (size_t)n_endog * c->P;
  if (c->d_x) cudaFree(c->d_x);
  if (c->d_F) cudaFree(c->d_F);
  if (c->d_Z) cudaFree(c->d_Z);
  if (c->d_V) cudaFree(c->d_V);
  if (c->d_JV) cudaFree(c->d_JV);
  c->d_x = c->d_F = c->d_Z = c->d_V = c->d_JV = nullptr; c->Vcap = 0;
  CK(cudaMalloc((void**)&c->d_x, n * sizeof(double)));
  CK(cudaMalloc((void**)&c->d_F, n * sizeof(double)));
  CK(cudaMalloc((void**)&c->d_Z, std::max<size_t>((size_t)n_exog * c->P, 1) * sizeof(double)));
  c->n_endog = n_endog; c->n_exog = n_exog; c->eq_ir = ir; c->eq_iw = iw;
  c->eq_on = true; c->ks_ready = true; c->linearized = false; c->jbar_valid = false;
  return HANK_OK;
}
