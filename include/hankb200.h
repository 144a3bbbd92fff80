/* hankb200.h — C ABI of libhankb200.so: the B200 (sm_100a) household block of a
 * sequence-space HANK model, as a drop-in for the hot path of
 * vasudeva-ram/Julia-NewtonRaphsonHANK.  Citations `File.jl:lines` are into that repository.
 *
 * The reference has no FFI of its own (pure Julia).  Each entry point below replaces the
 * Julia function it cites; INTEGRATION.md shows the `ccall` stubs a maintainer would add.
 *
 * Conventions
 *   - All arrays are Float64 in Julia column-major order, passed as plain pointers:
 *       n_a x n_e matrices (value, policy, D):  idx = (e-1)*n_a + (a-1)   (a fastest)
 *       Pi:  n_e x n_e column-major, Pi[(e-1) + n_e*(e2-1)] = Π[e,e2], row-stochastic
 *       x :  n_endog x P column-major, variable fastest (Y, KS, r, w), P = T-1
 *       paths r, w, Z, KD: length P;  lane tangents: P x K (or n x K) column-major
 *   - Periods t = 1..P and grid indices are 1-based in every argument and error message
 *     (as in Julia); arrays are 0-based in memory as usual.
 *   - Functions without `_dev` take HOST pointers and are synchronous.  `_dev` variants take
 *     DEVICE pointers (same layouts), enqueue on the context's stream and return without
 *     synchronising; call hank_sync() before reading results.
 *   - Every function returns a status (0 = ok); hank_last_error() gives the message, which the
 *     Julia wrapper rethrows with error(...).  find_ss's line search (SteadyState.jl:199) relies
 *     on exceptions from inside F, so domain / knot failures are reported, never silent.
 *   - One context per GPU, not thread-safe; distinct contexts may be used from distinct threads
 *     or processes.  The library owns all device memory; no returned pointer outlives the ctx.
 */
#ifndef HANKB200_H
#define HANKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hank_ctx hank_ctx;

enum hank_status {
  HANK_OK = 0,
  HANK_ERR_ARG = 1,        /* bad argument / unsupported shape                             */
  HANK_ERR_DOMAIN = 2,     /* negative base under a non-integer power (Julia DomainError)  */
  HANK_ERR_KNOTS = 3,      /* endogenous-grid knots not strictly increasing (Interpolations)*/
  HANK_ERR_CUDA = 4,       /* CUDA runtime / cuSOLVER / NCCL failure                       */
  HANK_ERR_STATE = 5,      /* call order violated (e.g. forward before backward)           */
  HANK_ERR_NONMONOTONE = 6,/* policy not monotone in a: gather lottery not applicable      */
  HANK_ERR_NOCONV = 7      /* iteration cap reached                                        */
};

/* ---- context ------------------------------------------------------------------------- */

/* Model constants the household block reads on every value_fn call (KrusellSmith.jl:44-52):
 * wealth grid, productivity grid z, Π, β, γ, borrow_cons, and T (GeneralStructures.jl:166-174).
 * 2 <= n_a <= 2048, 1 <= n_e <= 11.  The sweep kernels exist for n_e in {3,5,7,9,11}; any other count (an even number
 * of income states, or the Kronecker product of several exogenous processes, ForwardIteration.jl:280-284) runs on the
 * next of those with absorbing zero-mass padding states that never leave the device: every array of this interface
 * keeps the caller's n_e, and the real states' results are unchanged.                            */
int hank_ctx_create(hank_ctx** out, int device, int n_a, int n_e, int T, const double* grid,
                    const double* z, const double* Pi, double beta, double gamma,
                    double borrow_cons);
void hank_ctx_destroy(hank_ctx* ctx);
const char* hank_last_error(hank_ctx* ctx);
const char* hank_version(void);
int hank_sync(hank_ctx* ctx);
/* CUDA-event timer on the context's stream (the stream every kernel of this ctx runs on). */
int hank_timer_start(hank_ctx* ctx);
int hank_timer_stop(hank_ctx* ctx, float* ms);
/* Kernels launched by this ctx since creation (for bench.py's gpu_launches). */
int64_t hank_launch_count(hank_ctx* ctx);
/* Per-kernel device times of the four sweep kernels, measured with CUDA events on the ctx stream
 * while enabled: index 0 backward primal, 1 backward tangent, 2 forward primal, 3 forward tangent.
 * hank_kernel_times synchronises, returns accumulated milliseconds and launch counts, and
 * optionally resets the accumulators.                                                        */
int hank_profile(hank_ctx* ctx, int enable);
int hank_kernel_times(hank_ctx* ctx, double* ms4, int64_t* count4, int reset);
/* Max tangent lanes one pass may carry with the memory currently reserved; hank_reserve_lanes
 * grows the reservation (policy tangents are 8*G*P bytes per lane).                          */
int hank_reserve_lanes(hank_ctx* ctx, int K);

/* The SteadyState record's fields the sweeps consume (SteadyState.jl:21-27):
 * ss_ending.value (terminal ∂V/∂a, BackwardIteration.jl:85) and ss_initial.D
 * (ForwardIteration.jl:293). */
int hank_set_terminal(hank_ctx* ctx, const double* value_T /* G */);
int hank_set_initial_dist(hank_ctx* ctx, const double* D0 /* G */);

/* ---- model plug-in: one EGM step, value_fn(value_next, xVals, model) ----------------- */
/* KrusellSmith.jl:43-83 with K ForwardDiff lanes (dual.jl:14-21). d* may be NULL when K = 0.
 * dvalue_next/dvalue/dpolicy are G x K; dr, dw length K.                                     */
int hank_egm_step(hank_ctx* ctx, const double* value_next, const double* dvalue_next, double r,
                  double w, int K, const double* dr, const double* dw, double* value,
                  double* policy, double* dvalue, double* dpolicy);

/* Inner VFI of get_xVals (SteadyState.jl:132-141): iterate the EGM step from a matrix of ones until
 * max|dValue| < eps on the primal values (at most max_iter steps), carrying K lanes.  Outputs as
 * hank_egm_step: the converged Value and the policy of the last step; *iters = steps after the first. */
int hank_vfi(hank_ctx* ctx, double r, double w, int K, const double* dr, const double* dw, double eps,
             int max_iter, double* value, double* policy, double* dvalue, double* dpolicy, int* iters);

/* ---- sweeps -------------------------------------------------------------------------- */
/* BackwardIteration(xVec_endog, exog_paths, model, ss_end) — BackwardIteration.jl:46-116.
 * Only r and w enter the KS household block (KrusellSmith.jl:53-54), so the sweep takes their
 * paths and K tangent lanes of them.  Policies and their tangents stay on the device.        */
int hank_backward(hank_ctx* ctx, const double* r, const double* w /* P */, int K,
                  const double* dr, const double* dw /* P x K */);
/* ForwardIteration(policy_seqs, model, ss_initial) — ForwardIteration.jl:253-311, on the
 * policies left on the device by hank_backward.  KD: P, dKD: P x K.                          */
int hank_forward(hank_ctx* ctx, double* KD, double* dKD);
/* Same, but for caller-supplied policies (P matrices, G x P) and tangents (G x P x K).  A policy
 * that is not monotone in a is handled by a scatter (atomic) sweep when K = 0 and rejected with
 * HANK_ERR_NONMONOTONE when tangent lanes are requested.                                       */
int hank_forward_policies(hank_ctx* ctx, const double* policy, int K, const double* dpolicy,
                          double* KD, double* dKD);
/* hank_backward followed by hank_forward.                                                   */
int hank_block(hank_ctx* ctx, const double* r, const double* w, int K, const double* dr,
               const double* dw, double* KD, double* dKD);
int hank_block_dev(hank_ctx* ctx, const double* r, const double* w, int K, const double* dr,
                   const double* dw, double* KD, double* dKD);

/* Accessors used by parity tests and by callers that need host matrices
 * (SteadyStateJacobian.jl:226-229, SteadyState.jl:225).  t in 1..P; lane 0 = primal,
 * lane l>=1 = tangent lane l of the last pass.                                              */
int hank_get_policy(hank_ctx* ctx, int t, int lane, double* out /* G */);
int hank_get_dist(hank_ctx* ctx, int t, double* out /* G */);         /* primal D_t          */
int hank_get_value_first(hank_ctx* ctx, int lane, double* out /* G */); /* ∂V/∂a at t = 1    */
/* Lottery brackets of the last forward pass: m = searchsortedfirst(grid, p), 1-based Int32
 * (ForwardIteration.jl:52).                                                                  */
int hank_get_brackets(hank_ctx* ctx, int t, int32_t* m /* G */);
/* make_endogenous_transition's bracket rule on an arbitrary policy (ForwardIteration.jl:46-75):
 * m (1-based) and the weight on node m.                                                      */
int hank_lottery(hank_ctx* ctx, const double* policy /* G */, int32_t* m, double* omega);

/* ---- Krusell-Smith full function F(x) and JVPs ---------------------------------------- */
/* Constants of the compiled residuals (KrusellSmith.yaml:90-94, ModelParser.jl:217-259) and the
 * lag padding column (GeneralStructures.jl:350-354): ss_start.vars.KS.                       */
int hank_ks_configure(hank_ctx* ctx, double alpha, double delta, double ss_start_KS);
/* fullFunction(x) (NewtonRaphson.jl:77-83): sweeps at x, residuals F (length n = 4P), and the
 * linearisation (primal tape) later JVPs at the same x reuse.  Z: exogenous path, length P.  */
int hank_ks_linearize(hank_ctx* ctx, const double* x, const double* Z, double* F);
/* JVP(fullFunction, x, V[:,k]) for K directions at the x of the last hank_ks_linearize
 * (GeneralStructures.jl:542-550).  V, JV: n x K column-major.                                */
int hank_ks_jvp(hank_ctx* ctx, int K, const double* V, double* JV);
/* hank_ks_linearize followed by hank_ks_jvp in one call (host buffers); the seed upload overlaps the
 * primal backward sweep.                                                                      */
int hank_ks_fjvp(hank_ctx* ctx, const double* x, const double* Z, int K, const double* V, double* F,
                 double* JV);
int hank_ks_linearize_dev(hank_ctx* ctx, const double* x, const double* Z, double* F);
int hank_ks_jvp_dev(hank_ctx* ctx, int K, const double* V, double* JV);
/* Columns [col_begin, col_end) (1-based, half-open on the right: col_begin..col_end-1) of the
 * sequence-space Jacobian at the linearisation point, as JVPs with unit seeds — the
 * directJVPJacobian pattern (SteadyState.jl:296-320) generalised to any column range.
 * Columns of Y and KS have an identically zero household tangent and skip the sweeps.
 * J: n x (col_end - col_begin).                                                              */
int hank_ks_jacobian_columns(hank_ctx* ctx, int col_begin, int col_end, double* J);
int hank_ks_jacobian_columns_dev(hank_ctx* ctx, int col_begin, int col_end, double* J);
/* Same for an arbitrary ascending list of 1-based columns (host array): the multi-GPU Jacobian build deals the
 * periods round-robin over the ranks so that every rank holds the same mix of seed horizons.  J: n x ncols,
 * device pointer, columns in list order.                                                                     */
int hank_ks_jacobian_column_list_dev(hank_ctx* ctx, int ncols, const int* cols, double* J);
int hank_ks_jacobian_column_list(hank_ctx* ctx, int ncols, const int* cols, double* J /* host */);

/* ---- Newton-Raphson driver ------------------------------------------------------------ */
/* NewtonRaphsonHANK / y_Iteration (NewtonRaphson.jl:27-114) with the sweeps, the residuals and
 * the preconditioner solve on the device.  Jbar: n x n column-major (the steady-state Jacobian
 * matrix, NewtonRaphson.jl:97).  solver: 0 = restarted GMRES(20) with IterativeSolvers 0.9.4
 * defaults (reference-faithful), 1 = LU factorisation of Jbar (exact preconditioner solve),
 * 2 = as 1 with J(x) assembled once per outer iteration from batched unit-seed lanes, so the
 * inner J(x)*y products are GEMVs (same quantity as JVP(fullFunction, x, y)).
 * stats[0..4] = outer iterations, JVPs, F evaluations, final ||y||, GMRES iterations;
 * inner_counts (may be NULL) receives up to 100 inner-iteration counts.                      */
int hank_newton_solve(hank_ctx* ctx, const double* Jbar, const double* x0, const double* Z,
                      double eps, double eps_inner, int solver, double* x_out, double* stats,
                      int* inner_counts);

/* The dense contraction on its own: Ainv = A^-1 for an n x n column-major FP64 matrix (host buffers), by the
 * hand-written blocked Gauss-Jordan inverse with partial pivoting that hank_newton_solve uses for Jbar
 * (replaces the `gmres!(R, Jbar, rhs)` solves of NewtonRaphson.jl:97 by R = Jbar^-1 rhs, and Julia's
 * `Jbar \ rhs` / `inv(Jbar)` for callers that want the inverse).  HANK_ERR_CUDA with a message if a pivot is
 * exactly zero.                                                                                */
int hank_dense_inverse(hank_ctx* ctx, int n, const double* A, double* Ainv);

/* ---- the model's equilibrium equations as device bytecode ---------------------------------- */
/* Replaces the built-in Krusell-Smith aggregate block by the model's own equations: the reference's compile_residuals
 * (ModelParser.jl:217-259) + assemble_full_xMat (GeneralStructures.jl:329-377) + the ForwardDiff pass through both.
 * Variables are numbered like var_names(model) (ModelParser.jl:357): n_endog endogenous, then the household
 * aggregate KD, then n_exog exogenous.  After this call hank_ks_linearize / _jvp / _fjvp / _jacobian_columns /
 * hank_newton_solve work on x of shape n_endog x P (variable fastest) and Z of shape P x n_exog (period fastest per
 * variable: Z[v*P + t]); the household block still takes r = x[ir, :] and w = x[iw, :] (0-based rows) and returns KD.
 * Equation i (one per endogenous variable) is the postfix program code[eq_off[i] .. eq_off[i+1]) computing LHS - RHS:
 *   0 k: push consts[k]    1 v s: push variable v at period t+s (steady-state boundary values ss_start / ss_end of
 *   length n_endog+1+n_exog outside 1..P)    2 + 3 - 4 * 5 / 6 ^ (binary)    7 neg 8 exp 9 log 10 sqrt (unary).
 * hankb200/equations.py (Python) and julia/HankB200.jl (Julia) compile the YAML equation strings to this form.      */
int hank_eq_configure(hank_ctx* ctx, int n_endog, int n_exog, int ir, int iw, const int* eq_off, const int* code,
                      int n_const, const double* consts, const double* ss_start, const double* ss_end);

/* ---- multi-GPU: shard lanes, all-gather the columns ------------------------------------- */
/* NCCL unique id (128 bytes) created on rank 0 and passed to every rank by the host program.  */
int hank_comm_unique_id(void* id128);
int hank_comm_init(hank_ctx* ctx, int nranks, int rank, const void* id128);
/* All-gather equal-sized column blocks: every rank contributes count doubles (device pointer)
 * and receives nranks*count in rank order.                                                    */
int hank_allgather_columns_dev(hank_ctx* ctx, const double* local, size_t count, double* all);
/* Same with HOST buffers (synchronous): what a Julia process holding its column block in a Matrix calls.       */
int hank_allgather_columns(hank_ctx* ctx, const double* local, size_t count, double* all);
int hank_comm_destroy(hank_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HANKB200_H */
