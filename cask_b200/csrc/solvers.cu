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
#include <cstdio>
#include <cstdlib>
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
constexpr int kVecTile = kVecThreads * kVecItems;
constexpr int kVecCtasPerSm = 2;  // vector kernels run as ONE wave of 2 x SMs CTAs, each owning a contiguous chunk

// Launch with programmatic stream serialization: the kernel may be scheduled while its predecessor drains; every
// kernel launched this way starts with pdl_enter() (griddepcontrol.wait) before it touches memory.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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

// Every vector kernel walks tiles of kVecTile elements; a thread owns kVecItems elements of a tile, kVecThreads apart
// (coalesced), loads all its operands into registers first and only then computes and stores, so kVecItems x (number
// of operand arrays) loads are in flight per thread whatever the compiler can prove about aliasing.
// Tile -> CTA map (c_vec_rr, CASK_B200_VEC_RR):
//   1 (default)  round robin: CTA c takes tiles c, c + grid, c + 2 grid ... - at any moment the resident wave reads ONE
//                window of grid x 16 KB per array, so consecutive requests reaching a DRAM channel fall into the same
//                few open rows (the order in which a plain copy, and the persistent SpMV kernel, sweep memory);
//   0            contiguous chunk per CTA (round 1): 2 x SMs x 6 arrays = 1 776 widely separated streams at once, which
//                measured 3.0 TB/s on the 72 n bytes of the fused CG update (ncu launch list, profiles/r2a_launches_summary.md).
// Either map is fixed for a given (n, grid), so the per-CTA partial sums - and the reductions - stay deterministic, and
// both phases of the fused kernels use the same map (a thread re-reads only what it wrote itself).
__constant__ int c_vec_rr = 1;
#define CB_TILE_LOOP(n)                                                                                             \
  const int64_t chunk_ = (((n) + gridDim.x - 1) / gridDim.x + kVecThreads - 1) / kVecThreads * kVecThreads;         \
  const int64_t lo_ = c_vec_rr ? (int64_t)blockIdx.x * kVecTile : (int64_t)blockIdx.x * chunk_;                     \
  const int64_t hi_ = c_vec_rr ? (int64_t)(n) : (lo_ + chunk_ < (n) ? lo_ + chunk_ : (n));                          \
  const int64_t step_ = c_vec_rr ? (int64_t)gridDim.x * kVecTile : (int64_t)kVecTile;                               \
  for (int64_t tile = lo_; tile < hi_; tile += step_)
#define CB_ITEMS for (int i = 0; i < kVecItems; i++)
#define CB_IDX(tile) ((tile) + threadIdx.x + (int64_t)i * kVecThreads)

// out[q] = sum_i partials[q*stride + i], i < count, in index order (deterministic). One CTA.  Used after kernels
// that cannot finish their reduction themselves (gather-CSR / two-part SpMV launches).
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
// r = b - Ax (Ax arrives in r), p = r, r.r            SparseLinearSolvers.hpp:189-198
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
cg_init_kernel(int64_t n, const double* __restrict__ b, double* __restrict__ r, double* __restrict__ p,
               const __grid_constant__ ReduceDesc rd) {
  __shared__ double red[kVecThreads / 32];
  double acc = 0.0;
  CB_TILE_LOOP(n) {
    double bv[kVecItems], rv[kVecItems];
#pragma unroll
    CB_ITEMS { const int64_t k = CB_IDX(tile); bv[i] = k < hi_ ? b[k] : 0.0; rv[i] = k < hi_ ? r[k] : 0.0; }
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      if (k < hi_) {
        const double v = bv[i] - rv[i];
        r[k] = v;
        p[k] = v;
        acc += v * v;
      }
    }
  }
  const double t = cta_sum(acc, red);
  grid_finish_reduce(rd, t, 0.0, blockIdx.x, gridDim.x);
}

// alpha = rsold / p.Ap; x += alpha p; r -= alpha Ap; r.r (summed, and all-reduced, by the grid's last CTA)
// SparseLinearSolvers.hpp:208-218
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
cg_update_xr_kernel(int64_t n, int it, const double* __restrict__ scal, const int32_t* __restrict__ flags,
                    const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ x,
                    double* __restrict__ r, const __grid_constant__ ReduceDesc rd, unsigned long long* trace) {
  trace_min(trace);
  pdl_enter();
  trace_min(trace ? trace + 1 : nullptr);
  if (flags[F_DONE]) return;
  __shared__ double red[kVecThreads / 32];
  const double alpha = scal[S_RS0 + (it & 1)] / scal[S_PAP];
  double acc = 0.0;
  CB_TILE_LOOP(n) {
    double pv[kVecItems], av[kVecItems], xv[kVecItems], rv[kVecItems];
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      const bool ok = k < hi_;
      pv[i] = ok ? p[k] : 0.0; av[i] = ok ? Ap[k] : 0.0; xv[i] = ok ? x[k] : 0.0; rv[i] = ok ? r[k] : 0.0;
    }
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      if (k < hi_) {
        x[k] = xv[i] + alpha * pv[i];
        const double v = rv[i] - alpha * av[i];
        r[k] = v;
        acc += v * v;
      }
    }
  }
  const double t = cta_sum(acc, red);
  grid_finish_reduce(rd, t, 0.0, blockIdx.x, gridDim.x);
  trace_max(trace ? trace + 2 : nullptr);
}

// convergence test + p = r + (rsnew/rsold) p                        SparseLinearSolvers.hpp:220-231
// Row-sharded on the peer path: entries of the new p that a neighbour stages are stored straight into its copy of p
// as they are produced, and the grid's last CTA publishes the new halo epoch.
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
cg_update_p_kernel(int64_t n, int it, double* __restrict__ scal, int32_t* __restrict__ flags,
                   const double* __restrict__ r, double* __restrict__ p, const PushDesc pd, unsigned long long* trace) {
  trace_min(trace);
  pdl_enter();
  trace_min(trace ? trace + 1 : nullptr);
  if (flags[F_DONE]) return;
  const double rsold = scal[S_RS0 + (it & 1)], rsnew = scal[S_RS0 + ((it + 1) & 1)];
  const bool converged = rsnew <= scal[S_TOL2];
  const double beta = rsnew / rsold;
  if (!converged) {
    // Row-sharded: the pushed ranges (the planes a neighbour stages) are updated FIRST and by ALL CTAs, 256-element
    // pieces each, so the NVLink stores leave from every SM at the start of the kernel instead of from the two or
    // three CTAs whose chunk happens to hold a plane; the chunk loop below then skips those rows.
    for (int sidx = 0; sidx < pd.nsend; sidx++) {
      const int64_t lo = pd.lo[sidx], hi = pd.hi[sidx];
      const int64_t piece = ((hi - lo + gridDim.x - 1) / gridDim.x + kVecThreads - 1) / kVecThreads * kVecThreads;
      const int64_t a = lo + (int64_t)blockIdx.x * piece, b = a + piece < hi ? a + piece : hi;
      for (int64_t k = a + threadIdx.x; k < b; k += kVecThreads) {
        bool seen = false;  // a row staged by two peers lies in two ranges: updated (and pushed to both) only once
        for (int e = 0; e < sidx; e++) seen |= (k >= pd.lo[e]) & (k < pd.hi[e]);
        if (seen) continue;
        const double v = r[k] + beta * p[k];
        p[k] = v;
        push_store(pd, k, v);
      }
    }
    CB_TILE_LOOP(n) {
      const bool halo = pd.nsend && push_overlaps(pd, tile, tile + kVecTile);
      double rv[kVecItems], pv[kVecItems];
#pragma unroll
      CB_ITEMS { const int64_t k = CB_IDX(tile); rv[i] = k < hi_ ? r[k] : 0.0; pv[i] = k < hi_ ? p[k] : 0.0; }
#pragma unroll
      CB_ITEMS {
        const int64_t k = CB_IDX(tile);
        if (k < hi_ && !(halo && push_contains(pd, k))) p[k] = rv[i] + beta * pv[i];
      }
    }
    push_signal(pd);
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
  trace_max(trace ? trace + 2 : nullptr);
}

// One kernel per iteration for everything after the SpMV (single rank, or row-sharded on the peer path):
//   phase 1   alpha = rsold / p.Ap; x += alpha p; r -= alpha Ap; r.r                        (:208-218)
//   barrier   the grid's last CTA sums the partials (+ all-reduce over peer memory) -> rsnew; all CTAs wait for it
//   phase 2   convergence test; p = r + (rsnew / rsold) p                                   (:220-231)
// Row-sharded: the rows of the new r that a neighbour stages are updated first, by all CTAs, and stored straight
// into the neighbour's copy of r; in phase 2 every rank recomputes the halo rows of p ITSELF from the halo rows of
// r it received and the old halo rows of p (same arithmetic as the owner, bit-identical).  The NVLink transfer is
// thereby off the critical path: it overlaps the all-reduce, and the following SpMV needs no halo wait at all.
// The grid (<= 2 CTAs per SM) is resident as a whole, which the barrier relies on.
// KEEP is a compile-time switch on purpose: with a run-time flag every load and store of the loops exists twice in the
// SASS under complementary predicates (with / without the L2 cache hint), and the predicated-off twins still take their
// turn in the load/store instruction queue - ncu showed "LG throttle" as the top stall and 2.9 TB/s on 72 n bytes
// (profiles/r2c_cg_iteration_ncu.md) where the same loop without twins streams 6.1 TB/s (profiles/r2d_stream_bench.txt).
template <bool KEEP>
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
cg_update_fused_kernel(int64_t n, int it, double* __restrict__ scal, int32_t* __restrict__ flags, double* p,
                       const double* __restrict__ Ap, double* __restrict__ x, double* r, const __grid_constant__ ReduceDesc rd,
                       const PushDesc pd, const HaloUpdate hu, const GatherDesc pap, unsigned long long* trace) {
  trace_min(trace);
  pdl_enter();
  trace_min(trace ? trace + 1 : nullptr);
  if (flags[F_DONE]) return;
  __shared__ double red[kVecThreads / 32];
  constexpr bool keep = KEEP;  // the vectors fit L2: evict-last on every access (peer.cuh: L2 residency control)
  const unsigned long long pol = l2_policy_evict_last();
  const unsigned int gen0 = *reinterpret_cast<volatile unsigned int*>(rd.gen);
  const double rsold = scal[S_RS0 + (it & 1)];
  // row-sharded: the SpMV's last CTA only PUBLISHED this rank's p.Ap to the peers; the W contributions are gathered here
  const double alpha = rsold / (pap.ctrl ? peer_gather_sum(pap, red) : scal[S_PAP]);
  double acc = 0.0;
  // ---- phase 1 ----
  for (int sidx = 0; sidx < pd.nsend; sidx++) {
    const int64_t lo = pd.lo[sidx], hi = pd.hi[sidx];
    const int64_t piece = ((hi - lo + gridDim.x - 1) / gridDim.x + kVecThreads - 1) / kVecThreads * kVecThreads;
    const int64_t a = lo + (int64_t)blockIdx.x * piece, b = a + piece < hi ? a + piece : hi;
    for (int64_t k = a + threadIdx.x; k < b; k += kVecThreads) {
      bool seen = false;  // a row staged by two peers lies in two ranges: updated (and pushed to both) only once
      for (int e = 0; e < sidx; e++) seen |= (k >= pd.lo[e]) & (k < pd.hi[e]);
      if (seen) continue;
      st_keep(x + k, ld_keep(x + k, pol, keep) + alpha * ld_keep(p + k, pol, keep), pol, keep);
      const double v = ld_keep(r + k, pol, keep) - alpha * ld_keep(Ap + k, pol, keep);
      st_keep(r + k, v, pol, keep);
      push_store(pd, k, v);
      acc += v * v;
    }
  }
  {
    CB_TILE_LOOP(n) {
      const bool halo = pd.nsend && push_overlaps(pd, tile, tile + kVecTile);
      double pv[kVecItems], av[kVecItems], xv[kVecItems], rv[kVecItems];
#pragma unroll
      CB_ITEMS {
        const int64_t k = CB_IDX(tile);
        const bool ok = k < hi_;
        pv[i] = ok ? ld_keep(p + k, pol, keep) : 0.0; av[i] = ok ? ld_keep(Ap + k, pol, keep) : 0.0;
        xv[i] = ok ? ld_keep(x + k, pol, keep) : 0.0; rv[i] = ok ? ld_keep(r + k, pol, keep) : 0.0;
      }
#pragma unroll
      CB_ITEMS {
        const int64_t k = CB_IDX(tile);
        if (k < hi_ && !(halo && push_contains(pd, k))) {
          st_keep(x + k, xv[i] + alpha * pv[i], pol, keep);
          const double v = rv[i] - alpha * av[i];
          st_keep(r + k, v, pol, keep);
          acc += v * v;
        }
      }
    }
  }
  push_signal(pd);  // the grid's last CTA publishes this epoch of r to the peers
  const double t = cta_sum(acc, red);
  grid_finish_reduce(rd, t, 0.0, blockIdx.x, gridDim.x, gen0);  // returns everywhere once scal[rs_new] is final
  // ---- phase 2 ----
  const double rsnew = *reinterpret_cast<volatile double*>(&scal[S_RS0 + ((it + 1) & 1)]);
  const bool converged = rsnew <= scal[S_TOL2];
  const double beta = rsnew / rsold;
  if (!converged) {
    for (int sidx = 0; sidx < pd.nsend; sidx++) {  // same thread <-> row map as in phase 1: each thread reads its own r
      const int64_t lo = pd.lo[sidx], hi = pd.hi[sidx];
      const int64_t piece = ((hi - lo + gridDim.x - 1) / gridDim.x + kVecThreads - 1) / kVecThreads * kVecThreads;
      const int64_t a = lo + (int64_t)blockIdx.x * piece, b = a + piece < hi ? a + piece : hi;
      for (int64_t k = a + threadIdx.x; k < b; k += kVecThreads) {
        bool seen = false;
        for (int e = 0; e < sidx; e++) seen |= (k >= pd.lo[e]) & (k < pd.hi[e]);
        if (!seen) st_keep(p + k, ld_keep(r + k, pol, keep) + beta * ld_keep(p + k, pol, keep), pol, keep);
      }
    }
    {
      CB_TILE_LOOP(n) {
        const bool halo = pd.nsend && push_overlaps(pd, tile, tile + kVecTile);
        double rv[kVecItems], pv[kVecItems];
#pragma unroll
        CB_ITEMS {
          const int64_t k = CB_IDX(tile);
          rv[i] = k < hi_ ? ld_keep(r + k, pol, keep) : 0.0; pv[i] = k < hi_ ? ld_keep(p + k, pol, keep) : 0.0;
        }
#pragma unroll
        CB_ITEMS {
          const int64_t k = CB_IDX(tile);
          if (k < hi_ && !(halo && push_contains(pd, k))) st_keep(p + k, rv[i] + beta * pv[i], pol, keep);
        }
      }
    }
    if (hu.nrecv) {
      // halo rows of p: the peers' r entries of this epoch (pushed in their phase 1) must have landed
      if (threadIdx.x == 0) {
        const unsigned long long want = *reinterpret_cast<volatile unsigned long long*>(&hu.ctrl->push_seq[hu.channel]);
        for (int q = 0; q < kMaxPeers; q++)
          if (hu.peer_mask & (1u << q)) peer_wait_ge(&hu.ctrl->halo_flag[hu.channel][q], want, &hu.ctrl->error);
      }
      __syncthreads();
      for (int sidx = 0; sidx < hu.nrecv; sidx++) {
        const int64_t lo = hu.lo[sidx], hi = hu.hi[sidx];
        const int64_t piece = ((hi - lo + gridDim.x - 1) / gridDim.x + kVecThreads - 1) / kVecThreads * kVecThreads;
        const int64_t a = lo + (int64_t)blockIdx.x * piece, b = a + piece < hi ? a + piece : hi;
        for (int64_t g = a + threadIdx.x; g < b; g += kVecThreads) {
          bool seen = false;
          for (int e = 0; e < sidx; e++) seen |= (g >= hu.lo[e]) & (g < hu.hi[e]);
          if (!seen) p[g] = __ldcg(r + g) + beta * p[g];  // written by a peer over NVLink: read past L1
        }
      }
    }
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
  trace_max(trace ? trace + 2 : nullptr);
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

// r = b - r(=Ax); r0 = r; r.r
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
bicg_residual_kernel(int64_t n, const int32_t* __restrict__ flags, int only_on_restart, const double* __restrict__ b,
                     double* __restrict__ r, double* __restrict__ r0, const double* __restrict__ Ax,
                     const __grid_constant__ ReduceDesc rd) {
  if (flags[F_DONE]) return;
  if (only_on_restart && !flags[F_RESTART]) return;
  __shared__ double red[kVecThreads / 32];
  double acc = 0.0;
  CB_TILE_LOOP(n) {
    double bv[kVecItems], av[kVecItems];
#pragma unroll
    CB_ITEMS { const int64_t k = CB_IDX(tile); bv[i] = k < hi_ ? b[k] : 0.0; av[i] = k < hi_ ? Ax[k] : 0.0; }
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      if (k < hi_) {
        const double v = bv[i] - av[i];
        r[k] = v;
        r0[k] = v;
        acc += v * v;
      }
    }
  }
  const double t = cta_sum(acc, red);
  grid_finish_reduce(rd, t, 0.0, blockIdx.x, gridDim.x);
}

// dot(s): out[0] = a.b ; out[1] = c.d (optional)
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
dot2_kernel(int64_t n, const int32_t* __restrict__ flags, const double* __restrict__ a, const double* __restrict__ b,
            const double* __restrict__ c, const double* __restrict__ d, const __grid_constant__ ReduceDesc rd) {
  pdl_enter();
  if (flags && (flags[F_DONE] || flags[F_RESTART])) return;
  __shared__ double red[kVecThreads / 32];
  double s0 = 0.0, s1 = 0.0;
  CB_TILE_LOOP(n) {
    double av[kVecItems], bv[kVecItems], cv[kVecItems], dv[kVecItems];
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      const bool ok = k < hi_;
      av[i] = ok ? a[k] : 0.0; bv[i] = ok ? b[k] : 0.0;
      cv[i] = ok && c ? c[k] : 0.0; dv[i] = ok && c ? d[k] : 0.0;
    }
#pragma unroll
    CB_ITEMS { s0 += av[i] * bv[i]; s1 += cv[i] * dv[i]; }
  }
  const double t0 = cta_sum(s0, red);
  const double t1 = cta_sum(s1, red);
  grid_finish_reduce(rd, t0, t1, blockIdx.x, gridDim.x);
}

// One thread: the scalar part of the loop head (Eigen BiCGSTAB.h: while-test, rho, restart test).
__global__ void bicg_head_kernel(double* scal, int32_t* flags) {
  pdl_enter();
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

// beta = (rho/rho_old)(alpha/w); p = r + beta (p - w v); y = invdiag * p   (+ halo push of y)
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
bicg_p_kernel(int64_t n, const double* __restrict__ scal, const int32_t* __restrict__ flags,
              const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ invd,
              double* __restrict__ p, double* __restrict__ y, const PushDesc pd) {
  pdl_enter();
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const double beta = (scal[B_RHO] / scal[B_RHO_OLD]) * (scal[B_ALPHA] / scal[B_W]);
  const double w = scal[B_W];
  CB_TILE_LOOP(n) {
    const bool push = pd.nsend && push_overlaps(pd, tile, tile + kVecTile);
    double rv[kVecItems], pv[kVecItems], vv[kVecItems], dv[kVecItems];
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      const bool ok = k < hi_;
      rv[i] = ok ? r[k] : 0.0; pv[i] = ok ? p[k] : 0.0; vv[i] = ok ? v[k] : 0.0; dv[i] = ok ? invd[k] : 0.0;
    }
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      if (k < hi_) {
        const double np = rv[i] + beta * (pv[i] - w * vv[i]);
        p[k] = np;
        const double yv = dv[i] * np;
        y[k] = yv;
        if (push) push_store(pd, k, yv);
      }
    }
  }
  push_signal(pd);
}

// alpha = rho / r0.v; s = r - alpha v; z = invdiag * s   (+ halo push of z)
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
bicg_s_kernel(int64_t n, double* __restrict__ scal, const int32_t* __restrict__ flags,
              const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ invd,
              double* __restrict__ s, double* __restrict__ z, const PushDesc pd) {
  pdl_enter();
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const double alpha = scal[B_RHO] / scal[B_R0V];
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[B_ALPHA] = alpha;  // read again by bicg_xr and the next bicg_p
  CB_TILE_LOOP(n) {
    const bool push = pd.nsend && push_overlaps(pd, tile, tile + kVecTile);
    double rv[kVecItems], vv[kVecItems], dv[kVecItems];
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      const bool ok = k < hi_;
      rv[i] = ok ? r[k] : 0.0; vv[i] = ok ? v[k] : 0.0; dv[i] = ok ? invd[k] : 0.0;
    }
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      if (k < hi_) {
        const double sv = rv[i] - alpha * vv[i];
        s[k] = sv;
        const double zv = dv[i] * sv;
        z[k] = zv;
        if (push) push_store(pd, k, zv);
      }
    }
  }
  push_signal(pd);
}

// w = t.s / t.t (0 if t.t == 0); x += alpha y + w z; r = s - w t; r.r and r0.r
__global__ void __launch_bounds__(kVecThreads, kVecCtasPerSm)
bicg_xr_kernel(int64_t n, const double* __restrict__ scal, const int32_t* __restrict__ flags,
               const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ s,
               const double* __restrict__ t, const double* __restrict__ r0, double* __restrict__ x,
               double* __restrict__ r, const __grid_constant__ ReduceDesc rd, double* __restrict__ scal_rw,
               int32_t* __restrict__ flags_rw, int fused_scalars) {
  pdl_enter();
  if (flags[F_DONE] || flags[F_RESTART]) return;
  __shared__ double red[kVecThreads / 32];
  const double tt = scal[B_TT];
  const double w = tt > 0.0 ? scal[B_TS] / tt : 0.0;
  const double alpha = scal[B_ALPHA];
  double a0 = 0.0, a1 = 0.0;
  CB_TILE_LOOP(n) {
    double yv[kVecItems], zv[kVecItems], sv[kVecItems], tv[kVecItems], qv[kVecItems], xv[kVecItems];
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      const bool ok = k < hi_;
      yv[i] = ok ? y[k] : 0.0; zv[i] = ok ? z[k] : 0.0; sv[i] = ok ? s[k] : 0.0;
      tv[i] = ok ? t[k] : 0.0; qv[i] = ok ? r0[k] : 0.0; xv[i] = ok ? x[k] : 0.0;
    }
#pragma unroll
    CB_ITEMS {
      const int64_t k = CB_IDX(tile);
      if (k < hi_) {
        x[k] = xv[i] + (alpha * yv[i] + w * zv[i]);
        const double nr = sv[i] - w * tv[i];
        r[k] = nr;
        a0 += nr * nr;
        a1 += qv[i] * nr;
      }
    }
  }
  const double t0 = cta_sum(a0, red);
  const double t1 = cta_sum(a1, red);
  const bool finished = grid_finish_reduce(rd, t0, t1, blockIdx.x, gridDim.x);
  // Single rank / peer path: the CTA that finished the reduction (every other CTA of the grid has read the scalars and
  // the flags by then) also runs the scalar end of this trip (bicg_tail_kernel) and the scalar head of the next one
  // (bicg_head_kernel: while-test, rho, restart test) - two single-thread launches less per iteration.  On the peer
  // path every rank holds bit-identical sums and takes the same decisions.
  if (fused_scalars && finished && threadIdx.x == 0) {
    scal_rw[B_W] = w;
    flags_rw[F_I] += 1;
    flags_rw[F_TRIPS] += 1;
    const bool go = scal_rw[B_RR] > scal_rw[B_TOL2] && flags_rw[F_I] < flags_rw[F_MAXIT];
    if (!go) {
      flags_rw[F_DONE] = 1;
    } else {
      scal_rw[B_RHO_OLD] = scal_rw[B_RHO];
      scal_rw[B_RHO] = scal_rw[B_R0R];
      const double eps2 = DBL_EPSILON * DBL_EPSILON;
      if (fabs(scal_rw[B_RHO]) < eps2 * scal_rw[B_R0SQ]) flags_rw[F_RESTART] = 1;
    }
    __threadfence();
  }
}

__global__ void bicg_tail_kernel(double* scal, int32_t* flags) {
  pdl_enter();
  if (flags[F_DONE] || flags[F_RESTART]) return;
  const double tt = scal[B_TT];
  scal[B_W] = tt > 0.0 ? scal[B_TS] / tt : 0.0;
  flags[F_I] += 1;
  flags[F_TRIPS] += 1;
}

#undef CB_TILE_LOOP
#undef CB_ITEMS
#undef CB_IDX

// partial dots written by one fused SpMV: interior and halo-dependent launches are sized separately when sharded
int spmv_partials(cask_b200_ctx* ctx) {
  return dist_active(ctx) && !peer_ready(ctx) ? spmv_num_ctas(ctx, 1) + spmv_num_ctas(ctx, 2) : spmv_num_ctas(ctx, 0);
}

int vec_grid(const cask_b200_ctx* ctx, int64_t n) {
  const int64_t units = (n + kVecThreads - 1) / kVecThreads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(units, (int64_t)ctx->sm_count * kVecCtasPerSm));
}

enum { T_SPMV = 0, T_VEC = 1, T_GEN = 2, T_COUNT = 4 };  // tickets of the in-kernel reductions
constexpr int kTraceFirst = 40, kTraceIters = 32;  // iterations covered by the CASK_B200_TRACE timeline

// The persistent SpMV kernel runs with the maximum shared-memory carve-out.  The vector kernels stream and gain
// nothing from L1, so they ask for the same carve-out: the SMs are not reconfigured between the kernels of an
// iteration and a dependent kernel can become resident while its predecessor drains.
int prefer_max_shared() {
  static thread_local int done_for_device = -1;
  int dev = -1;
  CB_CUDA(cudaGetDevice(&dev));
  if (done_for_device == dev) return CASK_B200_OK;
#define CB_CARVE(k) CB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))
  CB_CARVE(cg_init_kernel); CB_CARVE(cg_update_xr_kernel); CB_CARVE(cg_update_p_kernel); CB_CARVE(cg_update_fused_kernel<false>); CB_CARVE(cg_update_fused_kernel<true>);
  CB_CARVE(bicg_residual_kernel); CB_CARVE(dot2_kernel); CB_CARVE(bicg_head_kernel); CB_CARVE(bicg_p_kernel);
  CB_CARVE(bicg_s_kernel); CB_CARVE(bicg_xr_kernel); CB_CARVE(bicg_tail_kernel); CB_CARVE(bicg_restart_scalars_kernel);
  CB_CARVE(reduce_partials_kernel); CB_CARVE(jacobi_diag_kernel);
#undef CB_CARVE
  {
    const char* e = getenv("CASK_B200_VEC_RR");
    const int rr = e ? atoi(e) : 1;
    CB_CUDA(cudaMemcpyToSymbol(c_vec_rr, &rr, sizeof(rr)));
  }
  done_for_device = dev;
  return CASK_B200_OK;
}

int ensure_work(cask_b200_ctx* ctx, int nvec, int64_t len_full) {
  SolverWork& w = ctx->work;
  CB_TRY(prefer_max_shared());
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
  if (!w.d_tickets) {
    CB_CUDA(cudaMalloc(&w.d_tickets, sizeof(unsigned int) * T_COUNT));
    CB_CUDA(cudaMemsetAsync(w.d_tickets, 0, sizeof(unsigned int) * T_COUNT, ctx->stream));
  }
  if (!w.h_flags) CB_CUDA(cudaMallocHost(&w.h_flags, sizeof(int32_t) * (F_COUNT + 1) * 4));
  if (!w.h_scalars) CB_CUDA(cudaMallocHost(&w.h_scalars, sizeof(double) * S_COUNT));
  const int64_t np = 2 * (int64_t)std::max(std::max(spmv_partials(ctx), vec_grid(ctx, ctx->plan.n)), 1);
  if (w.partials_len < np) {
    cudaFree(w.d_partials);
    w.d_partials = nullptr;
    w.partials_len = 0;
    CB_CUDA(cudaMalloc(&w.d_partials, sizeof(double) * np));
    w.partials_len = np;
  }
  return CASK_B200_OK;
}

// Descriptor of an in-kernel reduction of nq dot products into scal[slot0 ..): single rank -> the sums; peer path ->
// the all-reduced sums; NCCL path -> the local sums go to the staging slots and finish_reduce() adds the all-reduce.
ReduceDesc make_reduce(cask_b200_ctx* ctx, int ticket, int stride, int nq, int slot0) {
  SolverWork& w = ctx->work;
  ReduceDesc rd;
  rd.partials = w.d_partials;
  rd.stride = stride;
  rd.nq = nq;
  rd.ticket = w.d_tickets + ticket;
  const bool nccl = dist_active(ctx) && !peer_ready(ctx);
  rd.out = w.d_scalars + (nccl ? S_COUNT : 0) + slot0;
  peer_fill_reduce(ctx, &rd);
  return rd;
}
int finish_reduce(cask_b200_ctx* ctx, int nq, int slot0) {
  if (!dist_active(ctx) || peer_ready(ctx)) return CASK_B200_OK;
  SolverWork& w = ctx->work;
  return dist_allreduce_sum(ctx, w.d_scalars + S_COUNT + slot0, w.d_scalars + slot0, nq, ctx->stream);
}

// sums `count` per-CTA partials written by a launch that could not reduce in-kernel; across ranks too when sharded
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

// y = A x for a vector held in "full" layout (global length, own slice at row0_global).
//  * single rank, or row-sharded on the peer path (channel >= 0: d_full is vector `channel` of the symmetric arena
//    and its producer has pushed this epoch's halo): ONE persistent launch over all slices, interior first; the
//    kernel's producer warp acquires the peers' epoch flags when it reaches the first halo-dependent slice, and with
//    a fused dot the grid's last CTA finishes the reduction (and the all-reduce) in place -> scal[dot_slot];
//  * row-sharded over NCCL: grouped send/recv on the communication stream overlapped with the interior slices,
//    halo-dependent slices after the event, then a reduction kernel + ncclAllReduce.
// flags != nullptr: the launches do nothing once the solver's DONE / RESTART flag is up.
int spmv_full(cask_b200_ctx* ctx, double* d_full, double* d_y, const double* d_dot_with, int dot_slot, int channel,
              const int32_t* flags, unsigned long long* trace = nullptr, bool publish_only = false, int keep = 0,
              int nq = 1, int stride = 0) {
  cudaStream_t s = ctx->stream;
  SolverWork& w = ctx->work;
  const bool peer = channel >= 0 && peer_ready(ctx);
  SpmvFusion f;
  f.d_dot_with = d_dot_with;
  f.d_partials = w.d_partials;
  f.pdl = true;
  f.trace = trace;
  f.keep_vectors = keep;
  HaloWait hw;
  if (peer) hw = peer_halo_wait(ctx, channel);
  if (flags) { hw.skip0 = flags + F_DONE; hw.skip1 = flags + F_RESTART; }
  if (!dist_active(ctx) || peer) {
    const bool in_kernel = d_dot_with && spmv_single_launch(ctx);
    if (nq == 2 && !in_kernel) return fail(CASK_B200_ERR_RUNTIME, "fused y.w + y.y needs the single-launch persistent SpMV");
    if (in_kernel) f.reduce = make_reduce(ctx, T_SPMV, nq == 2 ? stride : 0, nq, dot_slot);  // nq == 2: scal[dot_slot + 1] = y.y
    f.fuse_self_dot = nq == 2 ? 1 : 0;
    if (publish_only) {
      if (!(in_kernel && peer)) return fail(CASK_B200_ERR_RUNTIME, "publish-only reduction needs the in-kernel peer path");
      f.reduce.publish_only = 1;
    }
    CB_TRY(launch_spmv(ctx, d_full, d_y, 0, s, &f, &hw));
    if (d_dot_with && !in_kernel) CB_TRY(reduce_dots(ctx, spmv_partials(ctx), 0, 1, dot_slot, flags));
    return CASK_B200_OK;
  }
  f.pdl = false;
  CB_TRY(dist_exchange_begin(ctx, d_full, s));
  SpmvFusion fb = f;
  fb.d_partials = f.d_partials + spmv_num_ctas(ctx, 1);
  CB_TRY(launch_spmv(ctx, d_full, d_y, 1, s, &f));    // interior rows: no remote x
  CB_TRY(dist_exchange_wait(ctx, s));
  CB_TRY(launch_spmv(ctx, d_full, d_y, 2, s, &fb));   // rows that read the halo
  if (d_dot_with) CB_TRY(reduce_dots(ctx, spmv_partials(ctx), 0, 1, dot_slot, flags));
  return CASK_B200_OK;
}

}  // namespace

void free_solver_work(cask_b200_ctx* ctx) {
  SolverWork& w = ctx->work;
  for (auto& v : w.d_vec) { cudaFree(v); v = nullptr; }
  cudaFree(w.d_scalars); cudaFree(w.d_partials); cudaFree(w.d_counters); cudaFree(w.d_tickets); cudaFree(w.d_trace);
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
  // fused update kernel (single rank / peer path): r lives in the arena too (vector 1) - its boundary rows are what
  // travels every iteration, and each rank recomputes the halo rows of p from them
  bool fused = !dist_active(ctx) || peer;
  if (fused) {
    // cg_update_fused_kernel holds a whole-grid barrier: every CTA of the grid must be resident at once.  The grid is
    // sized as ONE wave of kVecCtasPerSm CTAs per SM; that the device really co-schedules that many is asked of the
    // runtime (registers, carve-out, MPS partitions all count) instead of assumed; the barrier's spin itself carries
    // a timeout (peer.cuh: grid_finish_reduce traps after kPeerTimeoutNs), so SMs taken by another stream of the
    // process surface as a CUDA error on the host instead of a hung GPU.
    int resident = 0;
    CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cg_update_fused_kernel<false>, kVecThreads, 0));
    int resident_keep = 0;
    CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_keep, cg_update_fused_kernel<true>, kVecThreads, 0));
    resident = std::min(resident, resident_keep);
    if ((int64_t)resident * ctx->sm_count < vec_grid(ctx, n)) {
      if (peer) return fail(CASK_B200_ERR_RUNTIME, "cg: the fused update kernel's grid cannot be co-resident on this device");
      fused = false;  // two kernels (update_xr, update_p) with the reduction between them: no barrier needed
    }
  }
  // x, r, p, Ap of this rank fit L2 (with room for the matrix stream): keep them there (peer.cuh: L2 residency control).
  // evict-last lines live in the persisting set-aside of L2, which is 0 by default.  The set-aside is a device-wide carve-out
  // of the cache that every kernel of the process then runs with, so it is claimed ONLY for a solve that will use it and
  // given back otherwise: round 1 claimed the maximum once per context, and the BiCGStab solve that followed a CG solve in
  // the same process lost a third of its rate to the smaller normal L2 (84 vs 127 iterations/s on C5, profiles/r2e_*).
  int max_persist = 0;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
  const bool fits = max_persist > 0 && 4 * sizeof(double) * (size_t)n <= (size_t)max_persist;
  const int keep = ctx->l2_keep < 0 ? (fits ? 1 : 0) : (ctx->l2_keep && max_persist > 0 ? 1 : 0);
  const int64_t want_persist = keep ? max_persist : 0;
  if (ctx->l2_persist_bytes != want_persist) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want_persist) != cudaSuccess) cudaGetLastError();
    if (!want_persist) cudaCtxResetPersistingL2Cache();  // lines kept by an earlier solve become ordinary lines again
    ctx->l2_persist_bytes = want_persist;
  }
  if (getenv("CASK_B200_TRACE")) fprintf(stderr, "cask_b200: L2 persisting set-aside %lld bytes, keep=%d\n", (long long)ctx->l2_persist_bytes, keep);
  const PushDesc pd_r = peer_push_desc(ctx, 1);
  const HaloUpdate hu = peer_halo_update(ctx, 1);
  double* r = peer ? peer_vector(ctx, 1) + off : w.d_vec[0];
  double* p_full = peer ? peer_vector(ctx, 0) : w.d_vec[1];
  double* Ap = w.d_vec[2];
  double* p = p_full + off;
  double* scal = w.d_scalars;
  int32_t* flags = reinterpret_cast<int32_t*>(w.d_counters);
  const int vg = vec_grid(ctx, n);

  double h_scal[S_COUNT] = {0};
  h_scal[S_TOL2] = tol * tol;
  int32_t h_flags[F_COUNT + 1] = {0};
  h_flags[F_ITERATIONS] = iterations ? *iterations : 0;
  CB_CUDA(cudaMemcpyAsync(scal, h_scal, sizeof(h_scal), cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(flags, h_flags, sizeof(h_flags), cudaMemcpyHostToDevice, s));

  // r = A x (x staged through p's full-layout buffer), r = b - r, p = r, rsold = r.r   (:189-198)
  CB_CUDA(cudaMemcpyAsync(p, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
  CB_TRY(peer_push(ctx, 0, s));
  CB_TRY(spmv_full(ctx, p_full, r, nullptr, 0, ch, nullptr));
  cg_init_kernel<<<vg, kVecThreads, 0, s>>>(n, d_rhs, r, p, make_reduce(ctx, T_VEC, 0, 1, S_RS0));
  ctx->launches++;
  CB_TRY(finish_reduce(ctx, 1, S_RS0));
  CB_TRY(peer_push(ctx, 0, s));   // p = r; after the all-reduce, so no peer is still reading the previous epoch

  // CASK_B200_TRACE=<file>: in-situ timeline of iterations [kTraceFirst, kTraceFirst + kTraceIters)
  const char* trace_path = getenv("CASK_B200_TRACE");
  std::vector<unsigned long long> h_trace;
  if (trace_path) {
    h_trace.assign((size_t)kTraceIters * 9, 0ull);
    for (size_t i = 0; i < h_trace.size(); i++) h_trace[i] = (i % 3) == 2 ? 0ull : ~0ull;
    if (!w.d_trace) CB_CUDA(cudaMalloc(&w.d_trace, sizeof(unsigned long long) * h_trace.size()));
    CB_CUDA(cudaMemcpyAsync(w.d_trace, h_trace.data(), sizeof(unsigned long long) * h_trace.size(), cudaMemcpyHostToDevice, s));
  } else if (w.d_trace) {
    cudaFree(w.d_trace);
    w.d_trace = nullptr;
  }

  // Enqueue batches of iterations; poll the device's done flag one batch behind so the host never
  // stalls the stream.  Kernels of iterations enqueued past convergence see F_DONE and do nothing.
  const int kBatch = 8;
  cudaEvent_t ev[2] = {ctx->ev_a, ctx->ev_b};
  int32_t* hf = w.h_flags;
  int enq = 0;
  for (int batch = 0;; batch++) {
    const int hi = std::min(maxiters, enq + kBatch);
    for (int it = enq; it < hi; it++) {
      // three launches per iteration, chained by programmatic dependent launch; dots are finished (and all-reduced
      // over peer memory when sharded) by the last CTA of the kernel that produces them
      unsigned long long* tr = w.d_trace && it >= kTraceFirst && it < kTraceFirst + kTraceIters
                                   ? w.d_trace + (size_t)(it - kTraceFirst) * 9 : nullptr;
      const bool publish = fused && peer && spmv_single_launch(ctx);  // p.Ap gathered by the consumer kernel
      CB_TRY(spmv_full(ctx, p_full, Ap, p, S_PAP, ch, flags, tr, publish, keep));                       // :206
      const int rs_new = S_RS0 + ((it + 1) & 1);
      if (fused) {
        ReduceDesc rd = make_reduce(ctx, T_VEC, 0, 1, rs_new);
        rd.gen = w.d_tickets + T_GEN;
        GatherDesc gd;
        if (publish) { gd.ctrl = rd.ctrl; gd.world = rd.world; }
        if (keep)
          CB_CUDA(launch_pdl(cg_update_fused_kernel<true>, vg, kVecThreads, s, n, it, scal, flags, p, Ap, d_x, r, rd, pd_r, hu, gd,   // :208-231
                             tr ? tr + 3 : nullptr));
        else
          CB_CUDA(launch_pdl(cg_update_fused_kernel<false>, vg, kVecThreads, s, n, it, scal, flags, p, Ap, d_x, r, rd, pd_r, hu, gd,
                             tr ? tr + 3 : nullptr));
        ctx->launches += 1;
      } else {  // NCCL between the two halves
        CB_CUDA(launch_pdl(cg_update_xr_kernel, vg, kVecThreads, s, n, it, scal, flags, p, Ap, d_x, r,   // :208-218
                           make_reduce(ctx, T_VEC, 0, 1, rs_new), tr ? tr + 3 : nullptr));
        CB_TRY(finish_reduce(ctx, 1, rs_new));
        CB_CUDA(launch_pdl(cg_update_p_kernel, vg, kVecThreads, s, n, it, scal, flags, r, p, pd,         // :220-231 (+ halo push)
                           tr ? tr + 6 : nullptr));
        ctx->launches += 2;
      }
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
  if (trace_path && w.d_trace) {
    CB_CUDA(cudaMemcpy(h_trace.data(), w.d_trace, sizeof(unsigned long long) * h_trace.size(), cudaMemcpyDeviceToHost));
    std::string path = std::string(trace_path) + "." + std::to_string((long long)pl.row0_global);
    if (FILE* fp = fopen(path.c_str(), "w")) {
      fprintf(fp, "iteration,kernel,entry_ns,start_ns,end_ns\n");
      static const char* names[3] = {"spmv_dot", "update_xr", "update_p"};  // fused path: update_xr = the fused kernel
      for (int i = 0; i < kTraceIters; i++)
        for (int k = 0; k < 3; k++) {
          const unsigned long long* q = h_trace.data() + (size_t)i * 9 + k * 3;
          if (q[2] == 0ull) continue;
          fprintf(fp, "%d,%s,%llu,%llu,%llu\n", kTraceFirst + i, names[k], q[0], q[1], q[2]);
        }
      fclose(fp);
    }
  }
  if (iterations) *iterations = hf[F_ITERATIONS];
  if (converged) *converged = hf[F_CONVERGED];
  if (loop_trips) *loop_trips = hf[F_TRIPS];
  if (rs_final) *rs_final = w.h_scalars[S_RS_FINAL];
  if (ctx->l2_persist_bytes > 0) {  // the solve is over (stream synchronised above): the rest of the process gets its L2 back
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) != cudaSuccess) cudaGetLastError();
    cudaCtxResetPersistingL2Cache();
    ctx->l2_persist_bytes = 0;
  }
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
  // inverse diagonal of the Jacobi preconditioner: own buffer, released on every exit path (cudaFree waits for the
  // kernels that still read it)
  struct DeviceBuffer {
    double* p = nullptr;
    ~DeviceBuffer() { cudaFree(p); }
  } invd_owner;
  CB_CUDA(cudaMalloc(&invd_owner.p, sizeof(double) * std::max<int64_t>(n, 1)));
  double* const invd = invd_owner.p;
  double* scal = w.d_scalars;
  int32_t* flags = reinterpret_cast<int32_t*>(w.d_counters);
  const int vg = vec_grid(ctx, n);
  const int stride = std::max(std::max(spmv_partials(ctx), vg), 1);
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
  CB_TRY(spmv_full(ctx, y_full, t, nullptr, 0, ch_y, nullptr));
  bicg_residual_kernel<<<vg, kVecThreads, 0, st>>>(n, flags, 0, d_b, r, r0, t, make_reduce(ctx, T_VEC, 0, 1, B_R0SQ));
  CB_TRY(finish_reduce(ctx, 1, B_R0SQ));
  dot2_kernel<<<vg, kVecThreads, 0, st>>>(n, nullptr, d_b, d_b, nullptr, nullptr, make_reduce(ctx, T_VEC, 0, 1, B_RHSSQ));
  CB_TRY(finish_reduce(ctx, 1, B_RHSSQ));
  ctx->launches += 3;
  CB_CUDA(cudaMemcpyAsync(w.h_scalars, scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  const double rhs_sq = w.h_scalars[B_RHSSQ];
  if (rhs_sq == 0.0) {  // Eigen: x = 0, return
    *iters = 0; *tol_error = 0.0;
    return CASK_B200_OK;
  }
  h_scal[B_R0SQ] = w.h_scalars[B_R0SQ];
  h_scal[B_RHSSQ] = rhs_sq;
  h_scal[B_TOL2] = tol * tol * rhs_sq;
  h_scal[B_RR] = w.h_scalars[B_R0SQ];
  h_scal[B_R0R] = w.h_scalars[B_R0SQ];  // r0.r with r0 == r
  CB_CUDA(cudaMemcpyAsync(scal, h_scal, sizeof(h_scal), cudaMemcpyHostToDevice, st));

  // Single rank or peer path with every slice in one persistent launch: t.s and t.t are finished inside the second SpMV
  // (no dot2 pass over t and s) and the scalar tail / head run in bicg_xr's finishing CTA - 5 launches per iteration
  // (p, SpMV + r0.v, s, SpMV + t.s + t.t, xr) instead of 8.  The NCCL path keeps the separate kernels: its all-reduces
  // sit between them.
  const bool fused = spmv_single_launch(ctx) && (!dist_active(ctx) || peer) && !getenv("CASK_B200_BICG_UNFUSED");
  // body of one iteration after the loop head; every kernel no-ops while F_DONE or F_RESTART is up
  auto enqueue_body = [&]() -> int {
    CB_CUDA(launch_pdl(bicg_p_kernel, vg, kVecThreads, st, n, scal, flags, r, v, invd, p, y, pd_y));
    CB_TRY(spmv_full(ctx, y_full, v, r0, B_R0V, ch_y, flags));               // v = A y, r0.v finished in-kernel
    CB_CUDA(launch_pdl(bicg_s_kernel, vg, kVecThreads, st, n, scal, flags, r, v, invd, s, z, pd_z));   // alpha, s, z
    if (fused) {
      CB_TRY(spmv_full(ctx, z_full, t, s, B_TS, ch_z, flags, nullptr, false, 0, 2, stride));   // t = A z, t.s and t.t in-kernel
      CB_CUDA(launch_pdl(bicg_xr_kernel, vg, kVecThreads, st, n, scal, flags, y, z, s, t, r0, d_x, r,
                         make_reduce(ctx, T_VEC, stride, 2, B_RR), scal, flags, 1));
      ctx->launches += 3;
      return CASK_B200_OK;
    }
    CB_TRY(spmv_full(ctx, z_full, t, nullptr, 0, ch_z, flags));              // t = A z
    CB_CUDA(launch_pdl(dot2_kernel, vg, kVecThreads, st, n, flags, t, s, t, t, make_reduce(ctx, T_VEC, stride, 2, B_TS)));
    CB_TRY(finish_reduce(ctx, 2, B_TS));
    CB_CUDA(launch_pdl(bicg_xr_kernel, vg, kVecThreads, st, n, scal, flags, y, z, s, t, r0, d_x, r,
                       make_reduce(ctx, T_VEC, stride, 2, B_RR), scal, flags, 0));
    CB_TRY(finish_reduce(ctx, 2, B_RR));
    CB_CUDA(launch_pdl(bicg_tail_kernel, 1, 1, st, scal, flags));
    ctx->launches += 5;
    return CASK_B200_OK;
  };
  const int kBatch = 4;
  cudaEvent_t ev[2] = {ctx->ev_a, ctx->ev_b};
  int32_t* hf = w.h_flags;
  int64_t enq = 0;
  const int64_t enq_cap = (int64_t)maxit * 2 + 2;  // the first restart rewinds i once (BiCGSTAB.h)
  bool skip_check = true;
  if (fused) {  // the head of the first trip; every later head runs inside bicg_xr
    CB_CUDA(launch_pdl(bicg_head_kernel, 1, 1, st, scal, flags));
    ctx->launches++;
  }
  for (int batch = 0;; batch++) {
    for (int b = 0; b < kBatch && enq < enq_cap; b++, enq++) {
      if (!fused) {
        CB_CUDA(launch_pdl(bicg_head_kernel, 1, 1, st, scal, flags));
        ctx->launches++;
      }
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
        CB_TRY(spmv_full(ctx, z_full, t, nullptr, 0, ch_z, nullptr));
        bicg_residual_kernel<<<vg, kVecThreads, 0, st>>>(n, flags, 1, d_b, r, r0, t, make_reduce(ctx, T_VEC, 0, 1, B_TMP));
        CB_TRY(finish_reduce(ctx, 1, B_TMP));
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
  if (!fused) bicg_head_kernel<<<1, 1, 0, st>>>(scal, flags);  // evaluates the while-test one last time (sets DONE)
  CB_CUDA(cudaMemcpyAsync(hf, flags, sizeof(int32_t) * (F_COUNT + 1), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaMemcpyAsync(w.h_scalars, scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaStreamSynchronize(st));
  CB_CUDA(cudaGetLastError());
  *iters = hf[F_I];
  *tol_error = std::sqrt(w.h_scalars[B_RR] / rhs_sq);
  return peer_check_error(ctx);
}
