// Krylov loops around the SpMV, kept entirely on the device.
//
//  cask_b200_cg        pcg<double, IdentityPreconditioner>  src/runtime/SparseLinearSolvers.hpp:162-239
//  cask_b200_bicgstab  Eigen::BiCGSTAB (Jacobi precond.)    call sites src/runtime/SparseLinearSolvers.cpp:18-26,62-67
//
// Every scalar of the recurrences (rsold, p.Ap, alpha, beta, rho, omega ...) lives in device memory;
// the host only enqueues batches of iterations and polls a `done` flag that the kernels themselves
// honour, so the loop stops at exactly the iteration the reference would stop at without a host
// round trip per iteration.  Dots are reduced deterministically: one partial per CTA, summed in
// index order by a single CTA (and by NCCL across ranks when row-sharded).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "ctx.cuh"
#include "peer.cuh"

namespace caskb200 {

namespace {

// device scalar slots (doubles)
enum {
  S_RS0 = 0, S_RS1 = 1,  // r.r ping-pong: iteration i reads slot i&1, writes (i+1)&1
  S_PAP = 2, S_TOL2 = 3, S_RS_FINAL = 4,
  // BiCGStab
  B_RHO = 5, B_RHO_OLD = 6, B_ALPHA = 7, B_W = 8, B_R0V = 9, B_TS = 10, B_TT = 11, B_RR = 12, B_R0R = 13,
  B_R0SQ = 14, B_RHSSQ = 15, B_TOL2 = 16, B_TMP = 17,
  S_COUNT = 24  // when row-sharded, slots [S_COUNT, 2*S_COUNT) stage each rank's LOCAL sums: the all-reduce reads
                // them and writes the global value to the slot proper, so repeating it is harmless (iterations
                // enqueued past convergence still run their collectives)
};
// device flag slots (int32)
enum { F_DONE = 0, F_CONVERGED = 1, F_ITERATIONS = 2, F_TRIPS = 3, F_RESTART = 4, F_RESTARTS = 5, F_I = 6, F_MAXIT = 7, F_COUNT = 8 };

constexpr int kVecThreads = 256;
constexpr int kVecItems = 8;  // 8 coalesced 8-byte elements per thread and array: enough loads in flight to stream HBM

__device__ __forceinline__ double cta_sum(double v, double* red) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
  __syncthreads();
  return t;
}

// out[q] = sum_i partials[q*stride + i], i < count, in index order (deterministic). One CTA.
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const double* __restrict__ partials, int count, int stride, int nq, double* __restrict__ scal,
                       int slot0, int slot1, const int32_t* __restrict__ flags) {
  if (flags && (flags[F_DONE] || flags[F_RESTART])) return;
  __shared__ double red[32];
  for (int q = 0; q < nq; q++) {
    double v = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) v += partials[(size_t)q * stride + i];
    const double t = cta_sum(v, red);
    if (threadIdx.x == 0) scal[q == 0 ? slot0 : slot1] = t;
  }
}

// ---- CG -------------------------------------------------------------------------------------
// r = b - Ax (Ax arrives in r), p = r, partial r.r            SparseLinearSolvers.hpp:189-198
__global__ void __launch_bounds__(kVecThreads)
cg_init_kernel(int64_t n, const double* __restrict__ b, double* __restrict__ r, double* __restrict__ p,
               double* __restrict__ partials) {
  __shared__ double red[kVecThreads / 32];
  double acc = 0.0;
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      const double rv = b[base + (int64_t)i * kVecThreads] - r[base + (int64_t)i * kVecThreads];
      r[base + (int64_t)i * kVecThreads] = rv;
      p[base + (int64_t)i * kVecThreads] = rv;
      acc += rv * rv;
    }
  const double t = cta_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// alpha = rsold / p.Ap; x += alpha p; r -= alpha Ap; partial r.r    SparseLinearSolvers.hpp:208-218
__global__ void __launch_bounds__(kVecThreads)
cg_update_xr_kernel(int64_t n, int it, const double* __restrict__ scal, const int32_t* __restrict__ flags,
                    const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ x,
                    double* __restrict__ r, double* __restrict__ partials) {
  if (flags[F_DONE]) return;
  __shared__ double red[kVecThreads / 32];
  const double alpha = scal[S_RS0 + (it & 1)] / scal[S_PAP];
  double acc = 0.0;
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      x[base + (int64_t)i * kVecThreads] += alpha * p[base + (int64_t)i * kVecThreads];
      const double rv = r[base + (int64_t)i * kVecThreads] - alpha * Ap[base + (int64_t)i * kVecThreads];
      r[base + (int64_t)i * kVecThreads] = rv;
      acc += rv * rv;
    }
  const double t = cta_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// convergence test + p = r + (rsnew/rsold) p                        SparseLinearSolvers.hpp:220-231
__global__ void __launch_bounds__(kVecThreads)
cg_update_p_kernel(int64_t n, int it, double* __restrict__ scal, int32_t* __restrict__ flags,
                   const double* __restrict__ r, double* __restrict__ p, const PushDesc pd) {
  if (flags[F_DONE]) return;
  const double rsold = scal[S_RS0 + (it & 1)], rsnew = scal[S_RS0 + ((it + 1) & 1)];
  const bool converged = rsnew <= scal[S_TOL2];
  const double beta = rsnew / rsold;
  if (!converged) {
    const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
    // row-sharded: entries of the new p that a neighbour stages are stored into its copy of p as they are produced
    const bool push = pd.nsend && push_overlaps(pd, base - threadIdx.x, base - threadIdx.x + kVecThreads * kVecItems);
#pragma unroll
    for (int i = 0; i < kVecItems; i++)
      if (base + (int64_t)i * kVecThreads < n) {
        const double pv = r[base + (int64_t)i * kVecThreads] + beta * p[base + (int64_t)i * kVecThreads];
        p[base + (int64_t)i * kVecThreads] = pv;
        if (push) push_store(pd, base + (int64_t)i * kVecThreads, pv);
      }
    push_signal(pd);  // the grid's last CTA publishes the new halo epoch to the peers
  }
  // flags are only written by the grid's LAST CTA to finish, after every CTA has read them
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(&flags[F_COUNT]), 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    flags[F_COUNT] = 0;
    flags[F_TRIPS] = it + 1;
    scal[S_RS_FINAL] = rsnew;
    if (converged) { flags[F_CONVERGED] = 1; flags[F_DONE] = 1; }
    else flags[F_ITERATIONS] = it;   // :231 — assigned only at the end of a non-converged iteration
    __threadfence();
  }
}

// ---- BiCGStab ---------------------------------------------------------------------------------
__global__ void jacobi_diag_kernel(int64_t n, int64_t row0_global, const int32_t* __restrict__ row_ptr,
                                   const int32_t* __restrict__ col, const double* __restrict__ val,
                                   double* __restrict__ invdiag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0;
  for (int32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++)
    if (col[k] == i + row0_global) d = val[k];
  invdiag[i] = d != 0.0 ? 1.0 / d : 1.0;   // Eigen DiagonalPreconditioner
}

// r = b - r(=Ax); r0 = r; partial r.r
__global__ void __launch_bounds__(kVecThreads)
bicg_residual_kernel(int64_t n, const int32_t* __restrict__ flags, int only_on_restart, const double* __restrict__ b,
                     double* __restrict__ r, double* __restrict__ r0, const double* __restrict__ Ax,
                     double* __restrict__ partials) {
  if (flags[F_DONE]) return;
  if (only_on_restart && !flags[F_RESTART]) return;
  __shared__ double red[kVecThreads / 32];
  double acc = 0.0;
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      const double rv = b[base + (int64_t)i * kVecThreads] - Ax[base + (int64_t)i * kVecThreads];
      r[base + (int64_t)i * kVecThreads] = rv;
      r0[base + (int64_t)i * kVecThreads] = rv;
      acc += rv * rv;
    }
  const double t = cta_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// generic partial dot(s): partials[0][cta] = a.b ; partials[1][cta] = c.d (optional)
__global__ void __launch_bounds__(kVecThreads)
dot2_kernel(int64_t n, const int32_t* __restrict__ flags, const double* __restrict__ a, const double* __restrict__ b,
            const double* __restrict__ c, const double* __restrict__ d, double* __restrict__ partials, int stride) {
  if (flags && (flags[F_DONE] || flags[F_RESTART])) return;
  __shared__ double red[kVecThreads / 32];
  double s0 = 0.0, s1 = 0.0;
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      s0 += a[base + (int64_t)i * kVecThreads] * b[base + (int64_t)i * kVecThreads];
      if (c) s1 += c[base + (int64_t)i * kVecThreads] * d[base + (int64_t)i * kVecThreads];
    }
  const double t0 = cta_sum(s0, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t0;
  if (c) {
    const double t1 = cta_sum(s1, red);
    if (threadIdx.x == 0) partials[stride + blockIdx.x] = t1;
  }
}

// One thread: the scalar part of the loop head (Eigen BiCGSTAB.h: while-test, rho, restart test).
__global__ void bicg_head_kernel(double* scal, int32_t* flags) {
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const bool go = scal[B_RR] > scal[B_TOL2] && flags[F_I] < flags[F_MAXIT];
  if (!go) { flags[F_DONE] = 1; return; }
  scal[B_RHO_OLD] = scal[B_RHO];
  scal[B_RHO] = scal[B_R0R];
  const double eps2 = DBL_EPSILON * DBL_EPSILON;
  if (fabs(scal[B_RHO]) < eps2 * scal[B_R0SQ]) flags[F_RESTART] = 1;
}
// after a restart: rho = r0_sqnorm = r.r ; if (restarts++ == 0) i = 0
__global__ void bicg_restart_scalars_kernel(double* scal, int32_t* flags) {
  if (flags[F_DONE] || !flags[F_RESTART]) return;
  scal[B_RHO] = scal[B_TMP];
  scal[B_R0SQ] = scal[B_TMP];
  if (flags[F_RESTARTS]++ == 0) flags[F_I] = 0;
  flags[F_RESTART] = 0;
}

// beta = (rho/rho_old)(alpha/w); p = r + beta (p - w v); y = invdiag * p
__global__ void __launch_bounds__(kVecThreads)
bicg_p_kernel(int64_t n, const double* __restrict__ scal, const int32_t* __restrict__ flags,
              const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ invd,
              double* __restrict__ p, double* __restrict__ y, const PushDesc pd) {
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const double beta = (scal[B_RHO] / scal[B_RHO_OLD]) * (scal[B_ALPHA] / scal[B_W]);
  const double w = scal[B_W];
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
  const bool push = pd.nsend && push_overlaps(pd, base - threadIdx.x, base - threadIdx.x + kVecThreads * kVecItems);
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      const double pv = r[base + (int64_t)i * kVecThreads] + beta * (p[base + (int64_t)i * kVecThreads] - w * v[base + (int64_t)i * kVecThreads]);
      p[base + (int64_t)i * kVecThreads] = pv;
      const double yv = invd[base + (int64_t)i * kVecThreads] * pv;
      y[base + (int64_t)i * kVecThreads] = yv;
      if (push) push_store(pd, base + (int64_t)i * kVecThreads, yv);
    }
  push_signal(pd);
}

__global__ void bicg_alpha_kernel(double* scal, const int32_t* flags) {
  if (flags[F_DONE] || flags[F_RESTART]) return;
  scal[B_ALPHA] = scal[B_RHO] / scal[B_R0V];
}

// s = r - alpha v; z = invdiag * s
__global__ void __launch_bounds__(kVecThreads)
bicg_s_kernel(int64_t n, const double* __restrict__ scal, const int32_t* __restrict__ flags,
              const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ invd,
              double* __restrict__ s, double* __restrict__ z, const PushDesc pd) {
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const double alpha = scal[B_ALPHA];
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
  const bool push = pd.nsend && push_overlaps(pd, base - threadIdx.x, base - threadIdx.x + kVecThreads * kVecItems);
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      const double sv = r[base + (int64_t)i * kVecThreads] - alpha * v[base + (int64_t)i * kVecThreads];
      s[base + (int64_t)i * kVecThreads] = sv;
      const double zv = invd[base + (int64_t)i * kVecThreads] * sv;
      z[base + (int64_t)i * kVecThreads] = zv;
      if (push) push_store(pd, base + (int64_t)i * kVecThreads, zv);
    }
  push_signal(pd);
}

// w = t.s / t.t (0 if t.t == 0); x += alpha y + w z; r = s - w t; partials r.r and r0.r
__global__ void __launch_bounds__(kVecThreads)
bicg_xr_kernel(int64_t n, const double* __restrict__ scal, const int32_t* __restrict__ flags,
               const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ s,
               const double* __restrict__ t, const double* __restrict__ r0, double* __restrict__ x,
               double* __restrict__ r, double* __restrict__ partials, int stride) {
  if (flags[F_DONE] || flags[F_RESTART]) return;
  __shared__ double red[kVecThreads / 32];
  const double tt = scal[B_TT];
  const double w = tt > 0.0 ? scal[B_TS] / tt : 0.0;
  const double alpha = scal[B_ALPHA];
  double a0 = 0.0, a1 = 0.0;
  const int64_t base = (int64_t)blockIdx.x * (kVecThreads * kVecItems) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kVecItems; i++)
    if (base + (int64_t)i * kVecThreads < n) {
      x[base + (int64_t)i * kVecThreads] += alpha * y[base + (int64_t)i * kVecThreads] + w * z[base + (int64_t)i * kVecThreads];
      const double rv = s[base + (int64_t)i * kVecThreads] - w * t[base + (int64_t)i * kVecThreads];
      r[base + (int64_t)i * kVecThreads] = rv;
      a0 += rv * rv;
      a1 += r0[base + (int64_t)i * kVecThreads] * rv;
    }
  const double t0 = cta_sum(a0, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t0;
  const double t1 = cta_sum(a1, red);
  if (threadIdx.x == 0) partials[stride + blockIdx.x] = t1;
}

__global__ void bicg_tail_kernel(double* scal, int32_t* flags) {
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const double tt = scal[B_TT];
  scal[B_W] = tt > 0.0 ? scal[B_TS] / tt : 0.0;
  flags[F_I] += 1;
  flags[F_TRIPS] += 1;
}

// partial dots written by one fused SpMV: interior and halo-dependent launches are sized separately when sharded
int spmv_partials(cask_b200_ctx* ctx) {
  return dist_active(ctx) && !peer_ready(ctx) ? spmv_num_ctas(ctx, 1) + spmv_num_ctas(ctx, 2) : spmv_num_ctas(ctx, 0);
}

int vec_grid(int64_t n) { return (int)((n + (int64_t)kVecThreads * kVecItems - 1) / ((int64_t)kVecThreads * kVecItems)); }

int ensure_work(cask_b200_ctx* ctx, int nvec, int64_t len_full) {
  SolverWork& w = ctx->work;
  if (w.vec_len < len_full) {
    for (auto& v : w.d_vec) { cudaFree(v); v = nullptr; }
    w.vec_len = 0;
  }
  for (int i = 0; i < nvec; i++)
    if (!w.d_vec[i]) {
      CB_CUDA(cudaMalloc(&w.d_vec[i], sizeof(double) * std::max<int64_t>(len_full, 2)));
      CB_CUDA(cudaMemsetAsync(w.d_vec[i], 0, sizeof(double) * std::max<int64_t>(len_full, 2), ctx->stream));
    }
  w.vec_len = std::max(w.vec_len, len_full);
  if (!w.d_scalars) CB_CUDA(cudaMalloc(&w.d_scalars, sizeof(double) * 2 * S_COUNT));
  if (!w.d_counters) CB_CUDA(cudaMalloc(&w.d_counters, sizeof(int32_t) * (F_COUNT + 1)));
  if (!w.h_flags) CB_CUDA(cudaMallocHost(&w.h_flags, sizeof(int32_t) * (F_COUNT + 1) * 4));
  if (!w.h_scalars) CB_CUDA(cudaMallocHost(&w.h_scalars, sizeof(double) * S_COUNT));
  const int64_t np = 2 * (int64_t)std::max(std::max(spmv_partials(ctx), vec_grid(ctx->plan.n)), 1);
  static_assert(sizeof(int64_t) == 8, "");
  cudaFree(w.d_partials);
  w.d_partials = nullptr;
  CB_CUDA(cudaMalloc(&w.d_partials, sizeof(double) * np));
  return CASK_B200_OK;
}

// y = A x for a vector held in "full" layout (global length, own slice at row0_global): halo exchange
// overlapped with the interior slices when row-sharded.
// Peer-memory path (channel >= 0: d_full is vector `channel` of the symmetric arena and its producer has pushed
// this epoch's halo): ONE launch over all slices, interior first; the kernel's producer warp acquires the peers'
// epoch flags when it reaches the first halo-dependent slice.  flags != nullptr: the launch does nothing once the
// solver's DONE / RESTART flag is up (its producer skipped the push under the same condition).
int spmv_full(cask_b200_ctx* ctx, double* d_full, double* d_y, const SpmvFusion* f, int channel, const int32_t* flags) {
  cudaStream_t s = ctx->stream;
  HaloWait hw;
  if (channel >= 0 && peer_ready(ctx)) hw = peer_halo_wait(ctx, channel);
  if (flags) { hw.skip0 = flags + F_DONE; hw.skip1 = flags + F_RESTART; }
  if (!dist_active(ctx) || (channel >= 0 && peer_ready(ctx))) return launch_spmv(ctx, d_full, d_y, 0, s, f, &hw);
  CB_TRY(dist_exchange_begin(ctx, d_full, s));
  SpmvFusion fi, fb;
  const int n_int = spmv_num_ctas(ctx, 1);
  if (f) { fi = *f; fb = *f; fb.d_partials = f->d_partials + n_int; }
  CB_TRY(launch_spmv(ctx, d_full, d_y, 1, s, f ? &fi : nullptr));   // interior rows: no remote x
  CB_TRY(dist_exchange_wait(ctx, s));
  CB_TRY(launch_spmv(ctx, d_full, d_y, 2, s, f ? &fb : nullptr));   // rows that read the halo
  return CASK_B200_OK;
}

// sums `count` per-CTA partials (nq = 1 or 2 quantities, `stride` apart) into scal[slot0], scal[slot0 + 1];
// across ranks too when sharded
int reduce_dots(cask_b200_ctx* ctx, int count, int stride, int nq, int slot0, const int32_t* flags) {
  SolverWork& w = ctx->work;
  cudaStream_t s = ctx->stream;
  const bool dist = dist_active(ctx);
  if (peer_ready(ctx))
    return peer_allreduce_partials(ctx, w.d_partials, count, stride, nq, w.d_scalars, slot0, flags ? flags + F_DONE : nullptr,
                                   flags ? flags + F_RESTART : nullptr, s);
  double* dst = w.d_scalars + (dist ? S_COUNT : 0);
  reduce_partials_kernel<<<1, 1024, 0, s>>>(w.d_partials, count, stride, nq, dst, slot0, slot0 + 1, flags);
  ctx->launches++;
  if (dist) CB_TRY(dist_allreduce_sum(ctx, w.d_scalars + S_COUNT + slot0, w.d_scalars + slot0, nq, s));
  return CASK_B200_OK;
}

}  // namespace

void free_solver_work(cask_b200_ctx* ctx) {
  SolverWork& w = ctx->work;
  for (auto& v : w.d_vec) { cudaFree(v); v = nullptr; }
  cudaFree(w.d_scalars); cudaFree(w.d_partials); cudaFree(w.d_counters);
  if (w.h_flags) cudaFreeHost(w.h_flags);
  if (w.h_scalars) cudaFreeHost(w.h_scalars);
  w = SolverWork();
}

}  // namespace caskb200

using namespace caskb200;

extern "C" int cask_b200_cg_device(cask_b200_ctx* ctx, const double* d_rhs, double* d_x, int32_t maxiters,
                                   double tol, int32_t* iterations, int32_t* converged, double* rs_final,
                                   int32_t* loop_trips) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "cg: preprocess a matrix first");
  CB_TRY(ensure_device(ctx));
  Plan& pl = ctx->plan;
  if (pl.n_global != pl.m) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "cg: matrix must be square");
  const int64_t n = pl.n, off = pl.row0_global;
  cudaStream_t s = ctx->stream;
  CB_TRY(peer_ensure_arena(ctx, pl.m));  // collective; no-op unless row-sharded with the peer-memory path on
  CB_TRY(ensure_work(ctx, 3, pl.m));
  SolverWork& w = ctx->work;
  const bool peer = peer_ready(ctx);
  const int ch = peer ? 0 : -1;                       // p is vector 0 of the symmetric arena
  const PushDesc pd = peer_push_desc(ctx, 0);
  double* r = w.d_vec[0];
  double* p_full = peer ? peer_vector(ctx, 0) : w.d_vec[1];
  double* Ap = w.d_vec[2];
  double* p = p_full + off;
  double* scal = w.d_scalars;
  int32_t* flags = reinterpret_cast<int32_t*>(w.d_counters);
  const int vg = vec_grid(n);
  const bool dist = dist_active(ctx);

  double h_scal[S_COUNT] = {0};
  h_scal[S_TOL2] = tol * tol;
  int32_t h_flags[F_COUNT + 1] = {0};
  h_flags[F_ITERATIONS] = iterations ? *iterations : 0;
  CB_CUDA(cudaMemcpyAsync(scal, h_scal, sizeof(h_scal), cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, s));

  // r = A x (x staged through p's full-layout buffer), r = b - r, p = r, rsold = r.r   (:189-198)
  CB_CUDA(cudaMemcpyAsync(p, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
  CB_TRY(peer_push(ctx, 0, s));
  CB_TRY(spmv_full(ctx, p_full, r, nullptr, ch, nullptr));
  cg_init_kernel<<<vg, kVecThreads, 0, s>>>(n, d_rhs, r, p, w.d_partials);
  ctx->launches++;
  CB_TRY(reduce_dots(ctx, vg, 0, 1, S_RS0, nullptr));
  CB_TRY(peer_push(ctx, 0, s));   // p = r; after the all-reduce, so no peer is still reading the previous epoch

  // Enqueue batches of iterations; poll the device's done flag one batch behind so the host never
  // stalls the stream.  Kernels of iterations enqueued past convergence see F_DONE and do nothing.
  const int kBatch = 8;
  cudaEvent_t ev[2] = {ctx->ev_a, ctx->ev_b};
  int32_t* hf = w.h_flags;
  int enq = 0;
  for (int batch = 0;; batch++) {
    const int hi = std::min(maxiters, enq + kBatch);
    for (int it = enq; it < hi; it++) {
      SpmvFusion f;
      f.d_dot_with = p;
      f.d_partials = w.d_partials;
      CB_TRY(spmv_full(ctx, p_full, Ap, &f, ch, flags));                             // :206
      CB_TRY(reduce_dots(ctx, spmv_partials(ctx), 0, 1, S_PAP, flags));
      cg_update_xr_kernel<<<vg, kVecThreads, 0, s>>>(n, it, scal, flags, p, Ap, d_x, r, w.d_partials);   // :208-218
      CB_TRY(reduce_dots(ctx, vg, 0, 1, S_RS0 + ((it + 1) & 1), flags));
      cg_update_p_kernel<<<vg, kVecThreads, 0, s>>>(n, it, scal, flags, r, p, pd);     // :220-231 (+ halo push)
      ctx->launches += 2;
    }
    enq = hi;
    CB_CUDA(cudaMemcpyAsync(hf + (batch & 1) * (F_COUNT + 1), flags, sizeof(int32_t) * (F_COUNT + 1),
                            cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaEventRecord(ev[batch & 1], s));
    if (batch > 0) {
      CB_CUDA(cudaEventSynchronize(ev[(batch - 1) & 1]));
      if (hf[((batch - 1) & 1) * (F_COUNT + 1) + F_DONE]) break;
    }
    if (enq >= maxiters) break;
  }
  CB_CUDA(cudaMemcpyAsync(hf, flags, sizeof(int32_t) * (F_COUNT + 1), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(w.h_scalars, scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaGetLastError());
  if (iterations) *iterations = hf[F_ITERATIONS];
  if (converged) *converged = hf[F_CONVERGED];
  if (loop_trips) *loop_trips = hf[F_TRIPS];
  if (rs_final) *rs_final = w.h_scalars[S_RS_FINAL];
  return peer_check_error(ctx);
}

extern "C" int cask_b200_bicgstab_device(cask_b200_ctx* ctx, const double* d_b, double* d_x, int32_t* iters,
                                         double* tol_error) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "bicgstab: preprocess a matrix first");
  if (!iters || !tol_error) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "bicgstab: iters and tol_error are in/out");
  CB_TRY(ensure_device(ctx));
  Plan& pl = ctx->plan;
  if (pl.n_global != pl.m) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "bicgstab: matrix must be square");
  const int64_t n = pl.n, off = pl.row0_global;
  cudaStream_t st = ctx->stream;
  CB_TRY(peer_ensure_arena(ctx, pl.m));
  CB_TRY(ensure_work(ctx, 8, pl.m));
  SolverWork& w = ctx->work;
  const bool peer = peer_ready(ctx);
  const int ch_y = peer ? 0 : -1, ch_z = peer ? 1 : -1;   // y and z are vectors 0 and 1 of the symmetric arena
  const PushDesc pd_y = peer_push_desc(ctx, 0), pd_z = peer_push_desc(ctx, 1);
  double *r = w.d_vec[0], *r0 = w.d_vec[1], *v = w.d_vec[2], *p = w.d_vec[3];
  double *y_full = peer ? peer_vector(ctx, 0) : w.d_vec[4], *z_full = peer ? peer_vector(ctx, 1) : w.d_vec[5];
  double *s = w.d_vec[6], *t = w.d_vec[7];
  double *y = y_full + off, *z = z_full + off;
  // invdiag shares the tail of the partials allocation? keep it simple: own buffer
  double* invd = nullptr;
  CB_CUDA(cudaMalloc(&invd, sizeof(double) * std::max<int64_t>(n, 1)));
  double* scal = w.d_scalars;
  int32_t* flags = reinterpret_cast<int32_t*>(w.d_counters);
  const int vg = vec_grid(n);
  const int stride = std::max(std::max(spmv_partials(ctx), vg), 1);
  const bool dist = dist_active(ctx);
  const double tol = *tol_error > 0 ? *tol_error : DBL_EPSILON;
  const int64_t maxit64 = *iters > 0 ? *iters : 2 * pl.n_global;
  const int32_t maxit = (int32_t)std::min<int64_t>(maxit64, INT32_MAX);

  double h_scal[S_COUNT] = {0};
  h_scal[B_RHO] = 1; h_scal[B_ALPHA] = 1; h_scal[B_W] = 1;
  int32_t h_flags[F_COUNT + 1] = {0};
  h_flags[F_MAXIT] = maxit;
  CB_CUDA(cudaMemcpyAsync(scal, h_scal, sizeof(h_scal), cudaMemcpyHostToDevice, st));
  CB_CUDA(cudaMemcpyAsync(flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, st));
  jacobi_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, off, pl.d_row_ptr, pl.d_col, pl.d_val, invd);
  // Eigen's solve() starts from x = 0
  CB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, st));
  CB_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * n, st));
  CB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
  CB_CUDA(cudaMemsetAsync(y_full, 0, sizeof(double) * pl.m, st));
  // r = b - A x ; r0 = r ; r0_sq = r.r ; rhs_sq = b.b
  CB_TRY(peer_push(ctx, 0, st));
  CB_TRY(spmv_full(ctx, y_full, t, nullptr, ch_y, nullptr));
  bicg_residual_kernel<<<vg, kVecThreads, 0, st>>>(n, flags, 0, d_b, r, r0, t, w.d_partials);
  CB_TRY(reduce_dots(ctx, vg, 0, 1, B_R0SQ, nullptr));
  dot2_kernel<<<vg, kVecThreads, 0, st>>>(n, nullptr, d_b, d_b, nullptr, nullptr, w.d_partials, stride);
  CB_TRY(reduce_dots(ctx, vg, 0, 1, B_RHSSQ, nullptr));
  ctx->launches += 3;
  CB_CUDA(cudaMemcpyAsync(w.h_scalars, scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  const double rhs_sq = w.h_scalars[B_RHSSQ];
  if (rhs_sq == 0.0) {  // Eigen: x = 0, return
    cudaFree(invd);
    *iters = 0; *tol_error = 0.0;
    return CASK_B200_OK;
  }
  h_scal[B_R0SQ] = w.h_scalars[B_R0SQ];
  h_scal[B_RHSSQ] = rhs_sq;
  h_scal[B_TOL2] = tol * tol * rhs_sq;
  h_scal[B_RR] = w.h_scalars[B_R0SQ];
  h_scal[B_R0R] = w.h_scalars[B_R0SQ];  // r0.r with r0 == r
  CB_CUDA(cudaMemcpyAsync(scal, h_scal, sizeof(h_scal), cudaMemcpyHostToDevice, st));

  // body of one iteration after the loop head; every kernel no-ops while F_DONE or F_RESTART is up
  auto enqueue_body = [&]() -> int {
    bicg_p_kernel<<<vg, kVecThreads, 0, st>>>(n, scal, flags, r, v, invd, p, y, pd_y);
    SpmvFusion f1; f1.d_dot_with = r0; f1.d_partials = w.d_partials;
    CB_TRY(spmv_full(ctx, y_full, v, &f1, ch_y, flags));                     // v = A y, partials of r0.v
    CB_TRY(reduce_dots(ctx, spmv_partials(ctx), 0, 1, B_R0V, flags));
    bicg_alpha_kernel<<<1, 1, 0, st>>>(scal, flags);
    bicg_s_kernel<<<vg, kVecThreads, 0, st>>>(n, scal, flags, r, v, invd, s, z, pd_z);
    CB_TRY(spmv_full(ctx, z_full, t, nullptr, ch_z, flags));                  // t = A z
    dot2_kernel<<<vg, kVecThreads, 0, st>>>(n, flags, t, s, t, t, w.d_partials, stride);
    CB_TRY(reduce_dots(ctx, vg, stride, 2, B_TS, flags));
    bicg_xr_kernel<<<vg, kVecThreads, 0, st>>>(n, scal, flags, y, z, s, t, r0, d_x, r, w.d_partials, stride);
    CB_TRY(reduce_dots(ctx, vg, stride, 2, B_RR, flags));
    bicg_tail_kernel<<<1, 1, 0, st>>>(scal, flags);
    ctx->launches += 6;
    return CASK_B200_OK;
  };
  const int kBatch = 4;
  cudaEvent_t ev[2] = {ctx->ev_a, ctx->ev_b};
  int32_t* hf = w.h_flags;
  int64_t enq = 0;
  const int64_t enq_cap = (int64_t)maxit * 2 + 2;  // the first restart rewinds i once (BiCGSTAB.h)
  bool skip_check = true;
  for (int batch = 0;; batch++) {
    for (int b = 0; b < kBatch && enq < enq_cap; b++, enq++) {
      bicg_head_kernel<<<1, 1, 0, st>>>(scal, flags);
      ctx->launches++;
      CB_TRY(enqueue_body());
    }
    CB_CUDA(cudaMemcpyAsync(hf + (batch & 1) * (F_COUNT + 1), flags, sizeof(int32_t) * (F_COUNT + 1),
                            cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaEventRecord(ev[batch & 1], st));
    if (!skip_check) {
      CB_CUDA(cudaEventSynchronize(ev[(batch - 1) & 1]));
      const int32_t* q = hf + ((batch - 1) & 1) * (F_COUNT + 1);
      if (q[F_DONE]) break;
      if (q[F_RESTART]) {
        // rare: r became orthogonal to r0.  Everything enqueued since is idle; restart on the host's cue:
        // r = b - A x; r0 = r; rho = r0_sq = r.r; the interrupted iteration then continues.
        CB_CUDA(cudaMemcpyAsync(z, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        CB_TRY(peer_push(ctx, 1, st));
        CB_TRY(spmv_full(ctx, z_full, t, nullptr, ch_z, nullptr));
        bicg_residual_kernel<<<vg, kVecThreads, 0, st>>>(n, flags, 1, d_b, r, r0, t, w.d_partials);
        CB_TRY(reduce_dots(ctx, vg, 0, 1, B_TMP, nullptr));
        bicg_restart_scalars_kernel<<<1, 1, 0, st>>>(scal, flags);
        ctx->launches += 2;
        CB_TRY(enqueue_body());
        CB_CUDA(cudaStreamSynchronize(st));
        skip_check = true;   // the flag copy already in flight predates the restart
        continue;
      }
    }
    skip_check = false;
    if (enq >= enq_cap) break;
  }
  bicg_head_kernel<<<1, 1, 0, st>>>(scal, flags);  // evaluates the while-test one last time (sets DONE)
  CB_CUDA(cudaMemcpyAsync(hf, flags, sizeof(int32_t) * (F_COUNT + 1), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaMemcpyAsync(w.h_scalars, scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  CB_CUDA(cudaGetLastError());
  cudaFree(invd);
  *iters = hf[F_I];
  *tol_error = std::sqrt(w.h_scalars[B_RR] / rhs_sq);
  return peer_check_error(ctx);
}
