// hank_newton.cu — NewtonRaphsonHANK / y_Iteration (NewtonRaphson.jl:27-114) with every vector
// resident on the device: the sweeps, the residuals, the preconditioner solve and the updates
// run as kernels on the context's stream; the host only reads two norms per inner iteration to
// evaluate the reference's stopping rules.
//
// Preconditioner solve J̅·R = F(x) − J(x)·y (NewtonRaphson.jl:97):
//   solver 0: restarted GMRES(20), IterativeSolvers 0.9.4 defaults (restart = min(20,n),
//             maxiter = n, reltol = sqrt(eps) of the call's initial residual, modified
//             Gram-Schmidt), R reused as the initial guess — the reference's behaviour.  The
//             second solve `gmres!(M, J̅, Λxy)` (:98) only feeds a printed quantity (:101, :108)
//             and is skipped.
//   solver 1: J̅⁻¹ formed once by the hand-written blocked Gauss-Jordan inverse of hank_dense.cu (cached across
//             solves; HANK_CUSOLVER=1 switches back to cuSOLVER getrf/getrs on the identity for A/B runs), then
//             one FP64 GEMV per inner iteration.
//   solver 2: as 1, and J(x) itself is assembled once per outer iteration from batched unit-seed
//             tangent lanes (hank_ks_jacobian_columns_dev, all SMs busy) so that the ~40 inner
//             products J(x)·y become GEMVs instead of 40 strictly sequential single-lane sweeps.
//             Same quantity (JVP(fullFunction, x, y) = J(x)·y), different summation order.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>
#include <cusolverDn.h>
#include "hank_ctx.h"
#include "../../include/hankb200.h"

namespace hank {

#define CK(call)                                                   \
  do {                                                             \
    int rc__ = hank::cuda_check(c, (call), #call);                 \
    if (rc__) return rc__;                                         \
  } while (0)
#define RC(call)                 \
  do {                           \
    int rc__ = (call);           \
    if (rc__) return rc__;       \
  } while (0)

size_t dense_inverse_iwork(int n);                                          // hank_dense.cu
int dense_inverse_dev(hank_ctx* c, double* X, int n, double* out, int* iw);

constexpr int kSplit = 16;   // column splits of the GEMV
constexpr int kRestartMax = 20;

__device__ __forceinline__ double block_sum(double v, double* red) {
  // blockDim.x multiple of 32, <= 1024
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  if (w == 0) {
    s = l < nw ? red[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}

// part[s][i] = Σ_{j in split s} A[i + j n] v[j]   (A column-major)
// `done` (nullable): device flag of the speculative inner loop — once set, queued iterations are no-ops.
__global__ void k_gemv_partial(const double* __restrict__ A, const double* __restrict__ v, int n, double* part,
                               const double* done = nullptr) {
  if (done && *done != 0.0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  const int cs = (n + kSplit - 1) / kSplit;
  const int j0 = s * cs, j1 = min(n, j0 + cs);
  if (i >= n) return;
  double acc = 0.0;
  const double* a = A + (size_t)j0 * n + i;
  // eight loads in flight per thread (the matrix streams from L2); same summation order as the plain loop
#pragma unroll 8
  for (int j = j0; j < j1; ++j, a += n) acc = fma(__ldg(a), __ldg(v + j), acc);
  part[(size_t)s * n + i] = acc;
}
// out = a − Σ_s part[s]   (rhs = F(x) − J(x)·y from the split GEMV)
__global__ void k_sub_partial(const double* a, const double* __restrict__ part, int n, double* out,
                              const double* done = nullptr) {
  if (done && *done != 0.0) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double s = 0.0;
    for (int k = 0; k < kSplit; ++k) s += part[(size_t)k * n + i];
    out[i] = a[i] - s;
  }
}
__global__ void k_sub(const double* a, const double* b, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}
__global__ void k_fill(double* a, int n, double v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_identity(double* A, int n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)n * n) A[i] = (i / n == i % n) ? 1.0 : 0.0;
}
// scal[0] = ||a - b||, scal[1] = ||a||   (single block)
__global__ void k_norms(const double* a, const double* b, int n, double* scal) {
  __shared__ double red[33];
  double d = 0.0, s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { double t = a[i] - b[i]; d += t * t; s += a[i] * a[i]; }
  d = block_sum(d, red);
  s = block_sum(s, red);
  if (threadIdx.x == 0) { scal[0] = sqrt(d); scal[1] = sqrt(s); }
}
// LU mode: R = Σ_s part; y_old = y; y = y_old + α R; norms (single block)
// With eps_inner >= 0 (speculative loop): scal[3] counts the iterations that ran, scal[2] is raised as soon as
// the reference's loop test `eps_inner < ||y - yold||` fails (also for a non-finite step).
__global__ void k_update_from_partial(const double* __restrict__ part, int n, double alpha, double* R, double* y,
                                      double* yold, double* scal, double eps_inner = -1.0) {
  __shared__ double red[33];
  if (eps_inner >= 0.0 && scal[2] != 0.0) return;
  double d = 0.0, s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double r = 0.0;
    for (int k = 0; k < kSplit; ++k) r += part[(size_t)k * n + i];
    R[i] = r;
    const double yo = y[i], yn = yo + alpha * r;
    yold[i] = yo; y[i] = yn;
    const double t = yn - yo;
    d += t * t; s += yn * yn;
  }
  d = block_sum(d, red);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    scal[0] = sqrt(d); scal[1] = sqrt(s);
    if (eps_inner >= 0.0) { scal[3] += 1.0; if (!(eps_inner < sqrt(d))) scal[2] = 1.0; }
  }
}
__global__ void k_update_y(const double* __restrict__ R, int n, double alpha, double* y, double* yold, double* scal) {
  __shared__ double red[33];
  double d = 0.0, s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double yo = y[i], yn = yo + alpha * R[i];
    yold[i] = yo; y[i] = yn;
    const double t = yn - yo;
    d += t * t; s += yn * yn;
  }
  d = block_sum(d, red);
  s = block_sum(s, red);
  if (threadIdx.x == 0) { scal[0] = sqrt(d); scal[1] = sqrt(s); }
}
__global__ void k_xmy(double* x, const double* y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] -= y[i];
}

// ---- the whole inner loop of y_Iteration (NewtonRaphson.jl:94-106) in ONE cooperative launch (batched mode) ---------
// Per iteration: part = J(x)·y by column splits; rhs = F − Σ part; part = J̅⁻¹·rhs; R = Σ part, y_old = y,
// y = y_old + α R; stop when the reference's loop test eps_inner < ||y − y_old|| fails.  Grid = row blocks x kSplit
// column splits, three grid barriers per iteration; same column ranges and the same in-order sums as k_gemv_partial /
// k_sub_partial / k_update_from_partial, so the iterates are the ones the kernel-per-step path produces.
// scal: [0] ||y − y_old||, [1] ||y||, [3] iterations run (added).  npart: [rows blocks][2] norm partials.
constexpr int kInnerRows = 128;    // rows per CTA
constexpr int kInnerQ = 2;         // column halves of a split inside a CTA: kInnerRows x kInnerQ threads
constexpr int kInnerBlock = kInnerRows * kInnerQ;
// Grid barrier on a monotone 64-bit arrival counter (zeroed by the host before the launch; the cooperative launch
// guarantees that every CTA is resident).  Data written by other CTAs is read with ld.global.cg afterwards.
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long nblocks, unsigned long long& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    __threadfence();
    atomicAdd(ctr, 1ULL);
    const unsigned long long want = gen * nblocks;
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
    } while (seen < want);
  }
  __syncthreads();
}
// Σ_k p[k * stride] for k < kSplit, ascending k, with all loads in flight before the first add (a dependent chain of
// L2 round trips otherwise: 16 x ~700 cycles)
__device__ __forceinline__ double sum_partials(const double* p, size_t stride) {
  double v[kSplit];
#pragma unroll
  for (int k = 0; k < kSplit; ++k) v[k] = __ldcg(p + (size_t)k * stride);
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kSplit; ++k) s += v[k];
  return s;
}
__global__ void __launch_bounds__(kInnerBlock, 2)
k_newton_inner(const double* __restrict__ Jx, const double* __restrict__ Jinv, const double* __restrict__ F, int n,
               double alpha, double eps_inner, int max_it, double* y, double* yold, double* R, double* part,
               double* npart, double* scal, unsigned long long* bar_ctr) {
  const unsigned long long nblocks = (unsigned long long)gridDim.x * gridDim.y;
  unsigned long long gen = 0;
  constexpr int kMaxCs = (4096 + kSplit - 1) / kSplit;
  __shared__ double vs[kMaxCs + 8];                    // this split's slice of the multiplied vector
  __shared__ double qs[kInnerQ][kInnerRows];           // the column quarters' partial sums
  __shared__ double red[33];
  const int rb = blockIdx.x, s = blockIdx.y, nrb = gridDim.x;
  const int r = threadIdx.x % kInnerRows, q = threadIdx.x / kInnerRows;
  const int i = rb * kInnerRows + r;
  const int cs = (n + kSplit - 1) / kSplit, j0 = s * cs, j1 = min(n, j0 + cs);
  const int cq = (j1 - j0 + kInnerQ - 1) / kInnerQ, q0 = j0 + q * cq, q1 = min(j1, q0 + cq);
  // part[s][i] = Σ_{j in split} A[i + j n] vs[j - j0]: each quarter of the CTA sums its columns in ascending order with
  // up to 20 loads in flight per thread, the quarters are added in ascending order
  auto gemv_split = [&](const double* A) {
    double acc = 0.0;
    if (i < n) {
      const double* a = A + (size_t)q0 * n + i;
#pragma unroll 20
      for (int j = q0; j < q1; ++j, a += n) acc = fma(__ldg(a), vs[j - j0], acc);
    }
    qs[q][r] = acc;
    __syncthreads();
    if (q == 0 && i < n) {
      double t = qs[0][r];
#pragma unroll
      for (int k = 1; k < kInnerQ; ++k) t += qs[k][r];
      part[(size_t)s * n + i] = t;
    }
  };
  int it = 0;
  for (; it < max_it; ++it) {
    for (int j = j0 + threadIdx.x; j < j1; j += kInnerBlock) vs[j - j0] = __ldcg(y + j);
    __syncthreads();
    gemv_split(Jx);
    grid_barrier(bar_ctr, nblocks, gen);
    for (int j = j0 + threadIdx.x; j < j1; j += kInnerBlock) vs[j - j0] = F[j] - sum_partials(part + j, (size_t)n);   // rhs = F − J(x) y
    grid_barrier(bar_ctr, nblocks, gen);                         // everyone has read the first partials
    gemv_split(Jinv);
    grid_barrier(bar_ctr, nblocks, gen);
    if (s == 0) {                                                // one CTA per row block updates its rows
      double d = 0.0, qq = 0.0;
      if (q == 0 && i < n) {
        const double rr = sum_partials(part + i, (size_t)n);
        R[i] = rr;
        const double yo = __ldcg(y + i), yn = yo + alpha * rr;
        yold[i] = yo; y[i] = yn;
        const double t = yn - yo;
        d = t * t; qq = yn * yn;
      }
      d = block_sum(d, red);
      qq = block_sum(qq, red);
      if (threadIdx.x == 0) { npart[2 * rb] = d; npart[2 * rb + 1] = qq; }
    }
    grid_barrier(bar_ctr, nblocks, gen);
    double d = 0.0, qq = 0.0;
    for (int b0 = 0; b0 < nrb; b0 += 8) {                        // eight row blocks' partials in flight at a time
      double dv[8], qv[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) { dv[b] = b0 + b < nrb ? __ldcg(npart + 2 * (b0 + b)) : 0.0; qv[b] = b0 + b < nrb ? __ldcg(npart + 2 * (b0 + b) + 1) : 0.0; }
#pragma unroll
      for (int b = 0; b < 8; ++b) { d += dv[b]; qq += qv[b]; }
    }
    const double diff = sqrt(d);
    if (rb == 0 && s == 0 && threadIdx.x == 0) { scal[0] = diff; scal[1] = sqrt(qq); }
    if (!(eps_inner < diff)) { ++it; break; }                    // also leaves on a non-finite step
  }
  if (rb == 0 && s == 0 && threadIdx.x == 0) scal[3] += (double)it;
}

// ---- GMRES pieces (single block each; n ~ 10^3) -------------------------------------------
// gs: [0]=beta(residual.β) [1]=accumulator [2]=current [3]=beta of the cycle (init! return)
// V[:,0] = (b − Σ part) / β
__global__ void k_gmres_init(const double* __restrict__ part, const double* __restrict__ b, int n, double* V,
                             double* gs, double* nullvec) {
  __shared__ double red[33];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double ax = 0.0;
    for (int k = 0; k < kSplit; ++k) ax += part[(size_t)k * n + i];
    const double v = b[i] - ax;
    V[i] = v; s += v * v;
  }
  s = block_sum(s, red);
  const double beta = sqrt(s), ib = 1.0 / beta;
  for (int i = threadIdx.x; i < n; i += blockDim.x) V[i] *= ib;
  if (threadIdx.x == 0) { gs[0] = beta; gs[1] = 1.0; gs[2] = beta; gs[3] = beta; nullvec[0] = 1.0; }
}
// Arnoldi step k (1-based): w = Σ part = A V[:,k-1]; MGS against V[:,0..k-1]; H[:,k-1]; residual update
__global__ void k_gmres_step(const double* __restrict__ part, int n, int k, double* V, double* H, double* gs,
                             double* nullvec) {
  __shared__ double red[33];
  double* w = V + (size_t)k * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double a = 0.0;
    for (int s = 0; s < kSplit; ++s) a += part[(size_t)s * n + i];
    w[i] = a;
  }
  __syncthreads();
  const int ld = kRestartMax + 1;
  for (int c = 0; c < k; ++c) {
    const double* vc = V + (size_t)c * n;
    double h = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) h += vc[i] * w[i];
    h = block_sum(h, red);
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] -= h * vc[i];
    if (threadIdx.x == 0) H[(size_t)(k - 1) * ld + c] = h;
    __syncthreads();
  }
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += w[i] * w[i];
  s = block_sum(s, red);
  const double nr = sqrt(s), inr = 1.0 / nr;
  for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] *= inr;
  if (threadIdx.x == 0) {
    H[(size_t)(k - 1) * ld + k] = nr;
    double d = 0.0;
    for (int c = 0; c < k; ++c) d += nullvec[c] * H[(size_t)(k - 1) * ld + c];
    const double nv = -(d / nr);
    nullvec[k] = nv;
    gs[1] += nv * nv;
    gs[2] = gs[0] / sqrt(gs[1]);
  }
}
// Least squares on H[0:kk, 0:kk-1] by Givens rotations, x += V[:,0:kk-1] y
__global__ void k_gmres_update(int n, int kk, const double* __restrict__ V, const double* __restrict__ H,
                               const double* __restrict__ gs, double* x) {
  __shared__ double Hc[(kRestartMax + 1) * kRestartMax];
  __shared__ double rhs[kRestartMax + 1];
  const int ld = kRestartMax + 1, wd = kk - 1;
  for (int i = threadIdx.x; i < ld * kRestartMax; i += blockDim.x) Hc[i] = H[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i <= wd; ++i) rhs[i] = 0.0;
    rhs[0] = gs[3];
    for (int i = 0; i < wd; ++i) {
      const double a = Hc[i * ld + i], b = Hc[i * ld + i + 1];
      const double rr = hypot(a, b), cs = a / rr, sn = b / rr;
      Hc[i * ld + i] = rr; Hc[i * ld + i + 1] = 0.0;
      for (int j = i + 1; j < wd; ++j) {
        const double t1 = Hc[j * ld + i], t2 = Hc[j * ld + i + 1];
        Hc[j * ld + i] = cs * t1 + sn * t2;
        Hc[j * ld + i + 1] = -sn * t1 + cs * t2;
      }
      const double t1 = rhs[i], t2 = rhs[i + 1];
      rhs[i] = cs * t1 + sn * t2; rhs[i + 1] = -sn * t1 + cs * t2;
    }
    for (int i = wd - 1; i >= 0; --i) {
      double s = rhs[i];
      for (int j = i + 1; j < wd; ++j) s -= Hc[j * ld + i] * rhs[j];
      rhs[i] = s / Hc[i * ld + i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double a = x[i];
    for (int j = 0; j < wd; ++j) a += V[(size_t)j * n + i] * rhs[j];
    x[i] = a;
  }
}

struct NewtonBufs {
  double *x, *y, *yold, *Fx, *Lxy, *rhs, *R, *part, *scal, *Vb, *H, *gs, *nullvec, *J;
};

static int gemv_partial(hank_ctx* c, const double* A, const double* v, int n, double* part, const double* done = nullptr) {
  dim3 grid((n + 127) / 128, kSplit);
  k_gemv_partial<<<grid, 128, 0, c->stream>>>(A, v, n, part, done);
  c->launches++;
  return cuda_check(c, cudaGetLastError(), "k_gemv_partial");
}

// gmres!(x, A, b) on the device; returns inner iterations through *iters.
static int device_gmres(hank_ctx* c, const NewtonBufs& B, int n, double* x, const double* b, double* h_scal, long* iters) {
  const int restart = std::min(kRestartMax, n), maxiter = n;
  const double reltol = std::sqrt(std::numeric_limits<double>::epsilon());
  auto read_gs = [&](double* out) -> int {
    CK(cudaMemcpyAsync(out, B.gs, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  };
  RC(gemv_partial(c, B.J, x, n, B.part));
  k_gmres_init<<<1, 1024, 0, c->stream>>>(B.part, b, n, B.Vb, B.gs, B.nullvec);
  c->launches++;
  RC(read_gs(h_scal));
  double current = h_scal[2];
  const double tol = std::max(reltol * current, 0.0);
  int k = 1, iteration = 0;
  while (true) {
    if (iteration >= maxiter || current <= tol) {
      if (k > 1) { k_gmres_update<<<1, 1024, 0, c->stream>>>(n, k, B.Vb, B.H, B.gs, x); c->launches++; }
      break;
    }
    RC(gemv_partial(c, B.J, B.Vb + (size_t)(k - 1) * n, n, B.part));
    k_gmres_step<<<1, 1024, 0, c->stream>>>(B.part, n, k, B.Vb, B.H, B.gs, B.nullvec);
    c->launches++;
    RC(read_gs(h_scal));
    current = h_scal[2];
    ++k;
    if (k == restart + 1) {
      k_gmres_update<<<1, 1024, 0, c->stream>>>(n, k, B.Vb, B.H, B.gs, x);
      c->launches++;
      k = 1;
      if (!(iteration >= maxiter || current <= tol)) {
        RC(gemv_partial(c, B.J, x, n, B.part));
        CK(cudaMemsetAsync(B.H, 0, sizeof(double) * (kRestartMax + 1) * kRestartMax, c->stream));
        k_gmres_init<<<1, 1024, 0, c->stream>>>(B.part, b, n, B.Vb, B.gs, B.nullvec);
        c->launches++;
        // init_residual! of the new cycle resets residual.current to the true residual, which drives the
        // next done() test (IterativeSolvers gmres.jl)
        RC(read_gs(h_scal));
        current = h_scal[2];
      }
    }
    ++iteration;
  }
  *iters += iteration;
  return cuda_check(c, cudaGetLastError(), "gmres");
}

static int status_now(hank_ctx* c) {
  CK(cudaMemcpyAsync(c->h_status, c->d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->h_status[0] == 0) return 0;
  return hank_sync(c);  // formats the message and clears the flag
}

void newton_release(hank_ctx* c) {
  if (c->solver) { cusolverDnDestroy((cusolverDnHandle_t)c->solver); c->solver = nullptr; }
}

}  // namespace hank

using namespace hank;

extern "C" int hank_newton_solve(hank_ctx* c, const double* Jbar, const double* x0, const double* Z, double eps,
                                 double eps_inner, int solver, double* x_out, double* stats, int* inner_counts) {
  if (!c || !Jbar || !x0 || !Z || !x_out) return HANK_ERR_ARG;
  CK(cudaSetDevice(c->device));
  const auto t_entry = std::chrono::steady_clock::now();
  if (!c->ks_ready) return set_error(c, HANK_ERR_STATE, "hank_ks_configure has not been called");
  if (solver < 0 || solver > 2) return set_error(c, HANK_ERR_ARG, "solver must be 0 (gmres), 1 (lu) or 2 (lu, batched J(x))");
  const bool batched = solver == 2;
  if (batched) solver = 1;
  const int P = c->P, n = c->n_endog * P;
  const size_t nn = (size_t)n * n;
  // workspace
  const size_t nvec = 7, extra = (size_t)kSplit * n + 16 + (size_t)(kRestartMax + 1) * n +
                                 (size_t)(kRestartMax + 1) * kRestartMax + 8 + (kRestartMax + 1);
  // workspaces persist across solves (cudaMalloc / cudaFree next to multi-GB lane buffers cost milliseconds)
  const size_t need_vec = (nvec * n + extra) * sizeof(double), need_j = nn * sizeof(double) * 2;
  if (c->newton_bytes < need_vec) {
    if (c->d_newton) cudaFree(c->d_newton);
    c->d_newton = nullptr; c->newton_bytes = 0;
    CK(cudaMalloc((void**)&c->d_newton, need_vec));
    c->newton_bytes = need_vec;
  }
  if (c->jinv_bytes < need_j) {
    if (c->d_Jinv) cudaFree(c->d_Jinv);
    if (c->d_newton_i) cudaFree(c->d_newton_i);
    c->d_Jinv = nullptr; c->d_newton_i = nullptr; c->jinv_bytes = 0; c->jbar_valid = false;
    CK(cudaMalloc((void**)&c->d_Jinv, need_j));
    CK(cudaMalloc((void**)&c->d_newton_i, sizeof(int) * std::max<size_t>(n + 2, dense_inverse_iwork(n))));
    c->jinv_bytes = need_j;
  }
  double* Jx = nullptr;  // J(x), column-major, for the batched mode (reuses the LU scratch half)
  if (batched) Jx = c->d_Jinv + nn;
  NewtonBufs B;
  double* p = c->d_newton;
  B.x = p; p += n; B.y = p; p += n; B.yold = p; p += n; B.Fx = p; p += n; B.Lxy = p; p += n; B.rhs = p; p += n;
  B.R = p; p += n; B.part = p; p += (size_t)kSplit * n; B.scal = p; p += 16;
  B.Vb = p; p += (size_t)(kRestartMax + 1) * n; B.H = p; p += (size_t)(kRestartMax + 1) * kRestartMax;
  B.gs = p; p += 8; B.nullvec = p;
  B.J = c->d_Jinv;
  // J̅ is the same matrix for every solve around one steady state: its inverse is cached, keyed on a 64-bit hash of
  // the caller's buffer (four interleaved multiply-xor lanes over the 8-byte words, ~1 ms for n = 1196)
  uint64_t key = 0x9e3779b97f4a7c15ull ^ (uint64_t)n;
  {
    const uint64_t* wsrc = reinterpret_cast<const uint64_t*>(Jbar);
    uint64_t h0 = 1, h1 = 2, h2 = 3, h3 = 4;
    size_t i = 0;
    for (; i + 4 <= nn; i += 4) {
      h0 = (h0 ^ wsrc[i]) * 0x100000001b3ull; h1 = (h1 ^ wsrc[i + 1]) * 0x9e3779b97f4a7c15ull;
      h2 = (h2 ^ wsrc[i + 2]) * 0xc2b2ae3d27d4eb4full; h3 = (h3 ^ wsrc[i + 3]) * 0x165667b19e3779f9ull;
    }
    for (; i < nn; ++i) h0 = (h0 ^ wsrc[i]) * 0x100000001b3ull;
    key ^= h0 ^ (h1 << 1 | h1 >> 63) ^ (h2 << 2 | h2 >> 62) ^ (h3 << 3 | h3 >> 61);
  }
  const bool want_inverse = solver == 1;
  const bool cached = c->jbar_valid && c->jbar_key == key && c->jbar_n == n && c->jbar_inverse == want_inverse &&
                      getenv("HANK_NO_JBAR_CACHE") == nullptr;
  c->jbar_valid = false;
  if (!cached) CK(cudaMemcpyAsync(B.J, Jbar, nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(B.x, x0, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(B.y, x0, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_Z, Z, (size_t)c->n_exog * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(B.H, 0, sizeof(double) * (kRestartMax + 1) * kRestartMax, c->stream));
  if (solver == 1 && !cached && getenv("HANK_CUSOLVER") == nullptr) {
    // J̅⁻¹ once: blocked Gauss-Jordan with partial pivoting (hank_dense.cu); the second half of the buffer is the
    // work copy, the first half becomes the inverse
    double* W = c->d_Jinv + nn;
    CK(cudaMemcpyAsync(W, B.J, nn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    RC(dense_inverse_dev(c, W, n, B.J, c->d_newton_i));
    int h_info = 0;
    CK(cudaMemcpyAsync(&h_info, c->d_newton_i + 2 * n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h_info > 0)
      return set_error(c, HANK_ERR_CUDA, "LU factorisation of Jbar failed: the matrix is singular (U(" +
                                             std::to_string(h_info) + "," + std::to_string(h_info) + ") = 0)");
  } else if (solver == 1 && !cached) {
    // A/B path: LU (cuSOLVER getrf) then getrs on the identity
    cusolverDnHandle_t h = (cusolverDnHandle_t)c->solver;
    if (!h) {
      if (cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) return set_error(c, HANK_ERR_CUDA, "cusolverDnCreate failed");
      c->solver = h;
    }
    cusolverDnSetStream(h, c->stream);
    double* LU = c->d_Jinv + nn;  // second half holds the factors, first half becomes the inverse
    CK(cudaMemcpyAsync(LU, B.J, nn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    int lwork = 0;
    if (cusolverDnDgetrf_bufferSize(h, n, n, LU, n, &lwork) != CUSOLVER_STATUS_SUCCESS)
      return set_error(c, HANK_ERR_CUDA, "cusolverDnDgetrf_bufferSize failed");
    const size_t need_w = sizeof(double) * std::max(lwork, 1);
    if (c->lu_work_bytes < need_w) {
      if (c->d_lu_work) cudaFree(c->d_lu_work);
      c->d_lu_work = nullptr; c->lu_work_bytes = 0;
      CK(cudaMalloc((void**)&c->d_lu_work, need_w));
      c->lu_work_bytes = need_w;
    }
    double* work = c->d_lu_work; int* ipiv = c->d_newton_i;
    int* info = ipiv + n;   // info[0]: getrf (> 0: U(i,i) is exactly zero, J̅ singular), info[1]: getrs
    cusolverStatus_t s1 = cusolverDnDgetrf(h, n, n, LU, n, work, ipiv, info);
    k_identity<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>(B.J, n);
    c->launches += 2;
    cusolverStatus_t s2 = cusolverDnDgetrs(h, CUBLAS_OP_N, n, n, LU, n, ipiv, B.J, n, info + 1);
    int h_info[2] = {0, 0};
    CK(cudaMemcpyAsync(h_info, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h_info[0] > 0)
      return set_error(c, HANK_ERR_CUDA, "LU factorisation of Jbar failed: the matrix is singular (U(" +
                                             std::to_string(h_info[0]) + "," + std::to_string(h_info[0]) + ") = 0)");
    if (s1 != CUSOLVER_STATUS_SUCCESS || s2 != CUSOLVER_STATUS_SUCCESS || h_info[0] != 0 || h_info[1] != 0)
      return set_error(c, HANK_ERR_CUDA, "LU factorisation of Jbar failed (cuSOLVER error)");
  }
  c->jbar_key = key; c->jbar_n = n; c->jbar_inverse = want_inverse; c->jbar_valid = true;
  double h_scal[4] = {0, 0, 0, 0};
  auto read_scal = [&]() -> int {
    CK(cudaMemcpyAsync(h_scal, B.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->h_status, c->d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->h_status[0] != 0) return hank_sync(c);
    return 0;
  };
  const int nb = (n + 255) / 256;
  int outer = 1; long jvps = 0, fevals = 0, gm = 0;
  // HANK_NEWTON_TRACE=1: host wall-clock split of the solve (setup incl. LU / linearisations / inner loops) on stderr
  const bool trace = getenv("HANK_NEWTON_TRACE") != nullptr;
  auto now = [&]() { if (trace) cudaStreamSynchronize(c->stream); return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  double t_lin = 0.0, t_inner = 0.0;
  const auto t_setup_end = now();
  // ||y|| for the first outer test (y = x0)
  k_norms<<<1, 1024, 0, c->stream>>>(B.y, B.y, n, B.scal);
  c->launches++;
  RC(read_scal());
  double ynorm = h_scal[1];
  while (eps < ynorm && outer < 100) {
    const auto ta = now();
    RC(hank_ks_linearize_dev(c, B.x, c->d_Z, B.Fx));
    ++fevals;
    if (batched) RC(hank_ks_jacobian_columns_dev(c, 1, n + 1, Jx));
    const auto tb = now();
    t_lin += ms_since(ta, tb);
    k_fill<<<nb, 256, 0, c->stream>>>(B.yold, n, 1.0);
    k_fill<<<nb, 256, 0, c->stream>>>(B.R, n, 1.0);
    k_norms<<<1, 1024, 0, c->stream>>>(B.y, B.yold, n, B.scal);
    c->launches += 3;
    RC(read_scal());
    double diff = h_scal[0];
    int inner = 0;
    bool inner_done = false;
    if (batched && eps_inner >= 0.0 && eps_inner < diff && n <= 4096 && getenv("HANK_NEWTON_NO_COOP") == nullptr) {
      // the whole inner loop in one cooperative launch (all CTAs resident: checked against the device's occupancy)
      const dim3 cgrid((n + kInnerRows - 1) / kInnerRows, kSplit);
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_newton_inner, kInnerBlock, 0) == cudaSuccess &&
          (long)per_sm * c->sm_count >= (long)cgrid.x * cgrid.y) {
        double* npart = B.H;   // the GMRES Hessenberg block is unused in this mode: 2 doubles per row block
        CK(cudaMemsetAsync(B.scal + 2, 0, 2 * sizeof(double), c->stream));
        unsigned long long* a_ctr = reinterpret_cast<unsigned long long*>(B.H + 64);
        CK(cudaMemsetAsync(a_ctr, 0, sizeof(unsigned long long), c->stream));
        const double* a_Jx = Jx; const double* a_Ji = B.J; const double* a_F = B.Fx;
        int a_n = n, a_max = 1 << 20; double a_alpha = 0.5, a_eps = eps_inner;
        double *a_y = B.y, *a_yold = B.yold, *a_R = B.R, *a_part = B.part, *a_np = npart, *a_scal = B.scal;
        void* args[] = {&a_Jx, &a_Ji, &a_F, &a_n, &a_alpha, &a_eps, &a_max, &a_y, &a_yold, &a_R, &a_part, &a_np, &a_scal, &a_ctr};
        CK(cudaLaunchCooperativeKernel((const void*)k_newton_inner, cgrid, dim3(kInnerBlock), args, 0, c->stream));
        c->launches++;
        CK(cudaMemcpyAsync(h_scal, B.scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        diff = h_scal[0]; ynorm = h_scal[1];
        inner = (int)h_scal[3]; jvps += inner;
        if (!std::isfinite(diff)) return set_error(c, HANK_ERR_NOCONV, "Newton inner iteration diverged (non-finite step)");
        inner_done = true;
      } else {
        cudaGetLastError();
      }
    }
    if (!inner_done && batched && eps_inner >= 0.0 && eps_inner < diff) {
      // Speculative inner loop: iterations are queued kSpec at a time, the update kernel raises a device flag when
      // the loop test fails and everything queued behind it returns at once, so the host synchronises once per
      // kSpec iterations instead of once per iteration.  Same iterates, same count.
      int kSpec = 8;   // measured: 11.0 / 8.7 / 7.4 / 6.8 / 6.8 ms of inner loops at depth 1 / 2 / 4 / 8 / 16
      if (const char* sp = getenv("HANK_NEWTON_SPEC")) kSpec = std::max(1, std::min(64, atoi(sp)));   // A/B switch
      CK(cudaMemsetAsync(B.scal + 2, 0, 2 * sizeof(double), c->stream));
      while (true) {
        for (int q = 0; q < kSpec; ++q) {
          RC(gemv_partial(c, Jx, B.y, n, B.part, B.scal + 2));
          k_sub_partial<<<nb, 256, 0, c->stream>>>(B.Fx, B.part, n, B.rhs, B.scal + 2);
          RC(gemv_partial(c, B.J, B.rhs, n, B.part, B.scal + 2));
          k_update_from_partial<<<1, 1024, 0, c->stream>>>(B.part, n, 0.5, B.R, B.y, B.yold, B.scal, eps_inner);
          c->launches += 2;
        }
        CK(cudaMemcpyAsync(h_scal, B.scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h_status, c->d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (c->h_status[0] != 0) RC(hank_sync(c));
        diff = h_scal[0]; ynorm = h_scal[1];
        jvps += (long)h_scal[3] - inner; inner = (int)h_scal[3];
        if (!std::isfinite(diff)) return set_error(c, HANK_ERR_NOCONV, "Newton inner iteration diverged (non-finite step)");
        if (h_scal[2] != 0.0) break;
      }
    }
    while (eps_inner < diff) {
      if (batched) {
        RC(gemv_partial(c, Jx, B.y, n, B.part));
        k_sub_partial<<<nb, 256, 0, c->stream>>>(B.Fx, B.part, n, B.rhs);
      } else {
        RC(hank_ks_jvp_dev(c, 1, B.y, B.Lxy));
        k_sub<<<nb, 256, 0, c->stream>>>(B.Fx, B.Lxy, n, B.rhs);
      }
      ++jvps; ++inner;
      c->launches++;
      if (solver == 1) {
        RC(gemv_partial(c, B.J, B.rhs, n, B.part));
        k_update_from_partial<<<1, 1024, 0, c->stream>>>(B.part, n, 0.5, B.R, B.y, B.yold, B.scal);
        c->launches++;
      } else {
        RC(status_now(c));
        RC(device_gmres(c, B, n, B.R, B.rhs, h_scal, &gm));
        k_update_y<<<1, 1024, 0, c->stream>>>(B.R, n, 0.5, B.y, B.yold, B.scal);
        c->launches++;
      }
      RC(read_scal());
      diff = h_scal[0]; ynorm = h_scal[1];
      if (!std::isfinite(diff)) return set_error(c, HANK_ERR_NOCONV, "Newton inner iteration diverged (non-finite step)");
    }
    if (trace) t_inner += ms_since(tb, now());
    if (inner_counts && outer - 1 < 100) inner_counts[outer - 1] = inner;
    k_xmy<<<nb, 256, 0, c->stream>>>(B.x, B.y, n);
    c->launches++;
    ++outer;
  }
  CK(cudaMemcpyAsync(x_out, B.x, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (trace)
    fprintf(stderr, "[hank_newton_solve] setup+LU %.2f ms, linearise%s %.2f ms, inner loops %.2f ms (%ld iterations)\n",
            ms_since(t_entry, t_setup_end), batched ? "+J(x)" : "", t_lin, t_inner, jvps);
  if (stats) { stats[0] = outer - 1; stats[1] = (double)jvps; stats[2] = (double)fevals; stats[3] = ynorm; stats[4] = (double)gm; }
  if (eps < ynorm) return set_error(c, HANK_ERR_NOCONV, "Newton outer iteration cap (100) reached");
  return HANK_OK;
}
