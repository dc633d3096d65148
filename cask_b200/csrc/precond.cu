// Preconditioned CG: pcg<double, Precon> of the reference with the preconditioner left in
// (src/runtime/SparseLinearSolvers.hpp:162-239; IdentityPreconditioner :64-73, ILUPreconditioner :77-156), plus a
// Jacobi preconditioner (not in the reference).  Single rank.
//
// The ILU(0) factorisation and the two triangular solves are level-scheduled (precond_logic.inl); the CG recurrences
// around them are plain streaming kernels.  This loop is the straightforward one - one host synchronisation per
// iteration for the convergence test, dots as separate two-stage reductions - because with ILU the iteration is bound
// by the level-by-level solves, not by the vector passes; the identity-preconditioned loop in solvers.cu stays the
// tuned path for BASELINE's CG configuration.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ctx.cuh"
#include "devlogic.cuh"
#include "peer.cuh"

#include "precond_logic.inl"

namespace caskb200 {

struct PrecondState {
  precond::IluState ilu;
  bool ilu_ready = false;          // analysed + factored for the current matrix
  double* invd = nullptr;          // Jacobi
  bool jacobi_ready = false;
  const cask_b200_csr* pc_matrix = nullptr;  // optional: build the preconditioner from this matrix instead of A
  // the level sequence of one ILU application (one launch per dependency level: 1 786 for the 256^3 27-point system) as
  // ONE instantiated CUDA graph: captured on first use for a given (r, z, unit_lower), replayed by every later iteration
  // ... or, by default, ONE cooperative kernel that walks all levels of both solves with a grid barrier between them
  // (ilu_levels_kernel): level boundaries on the device, a monotonic barrier counter whose base the host tracks
  int64_t* d_level_ptr = nullptr;        // [ptr_l | ptr_u]
  unsigned long long* d_level_bar = nullptr;
  unsigned long long level_bar_base = 0;
  int64_t max_level_rows = 0;
  cudaGraphExec_t ilu_graph = nullptr;
  const double* graph_r = nullptr;
  double* graph_z = nullptr;
  int graph_unit = -1;
  int64_t graph_nodes = 0;
  // work vectors of the loop
  double* vec[4] = {nullptr, nullptr, nullptr, nullptr};  // r, p, Ap, z
  int64_t vec_len = 0;
  double* partials = nullptr;
  double* scal = nullptr;  // [0] p.Ap  [1] r.z
  int nparts = 0;
};

void free_precond(cask_b200_ctx* ctx) {
  PrecondState* st = ctx->precond;
  if (!st) return;
  precond::ilu_free(&st->ilu);
  if (st->ilu_graph) cudaGraphExecDestroy(st->ilu_graph);
  cudaFree(st->d_level_ptr);
  cudaFree(st->d_level_bar);
  cudaFree(st->invd);
  for (auto& v : st->vec) cudaFree(v);
  cudaFree(st->partials);
  cudaFree(st->scal);
  delete st;
  ctx->precond = nullptr;
}

namespace {

constexpr int kT = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kT / 32; w++) t += red[w];
  return t;  // valid in thread 0
}

// partials[c] = sum over CTA c's grid-stride elements of a[i] * b[i]
__global__ void __launch_bounds__(kT) pcg_dot_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                     double* __restrict__ partials) {
  __shared__ double red[kT / 32];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) acc += a[i] * b[i];
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// *out = partials[0] + ... + partials[count-1]; one CTA, fixed order -> deterministic
__global__ void __launch_bounds__(kT) pcg_finish_kernel(const double* __restrict__ partials, int count, double* __restrict__ out) {
  __shared__ double red[kT / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += kT) acc += partials[i];
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) *out = t;
}

// r = b - r   (r holds A x on entry)                                       :189-190
__global__ void __launch_bounds__(kT) pcg_residual_kernel(int64_t n, const double* __restrict__ b, double* __restrict__ r) {
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) r[i] = b[i] - r[i];
}

// alpha = rsold / p.Ap; x += alpha p; r -= alpha Ap                        :208-212
__global__ void __launch_bounds__(kT) pcg_xr_kernel(int64_t n, double rsold, const double* __restrict__ pap,
                                                    const double* __restrict__ p, const double* __restrict__ Ap,
                                                    double* __restrict__ x, double* __restrict__ r) {
  const double alpha = rsold / *pap;
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) {
    x[i] += alpha * p[i];
    r[i] -= alpha * Ap[i];
  }
}

// z = invd .* r
__global__ void __launch_bounds__(kT) pcg_scale_kernel(int64_t n, const double* __restrict__ invd, const double* __restrict__ r,
                                                       double* __restrict__ z) {
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) z[i] = invd[i] * r[i];
}

// p = z + beta p                                                            :229
__global__ void __launch_bounds__(kT) pcg_p_kernel(int64_t n, double beta, const double* __restrict__ z, double* __restrict__ p) {
  const int64_t stride = (int64_t)gridDim.x * kT;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += stride) p[i] = z[i] + beta * p[i];
}

// Both triangular solves of one ILU application in ONE launch: every level is a grid-stride loop over its rows (the same
// row functors as the per-level kernels: same arithmetic, same order, same bits), levels are separated by a grid barrier.
// A level of the 64^3 27-point twin holds ~600 rows and costs a per-level KERNEL about 7 us, most of it launch and drain;
// the barrier of a small co-resident grid costs 1-2 us.  Rows read the entries earlier levels wrote with L2 loads
// (dev::ld_l2): a stale L1 line must never serve them.  Launched cooperatively, so co-residency is guaranteed by the
// runtime; the spin traps after the peer timeout instead of hanging.
// (A thread-block cluster behind the hardware cluster barrier - barrier.cluster.arrive.release / wait.acquire - was tried for
// the narrow levels of the 64^3 twin, 2 CTAs x 512 threads: 7.7 us per level against 5.4 us for this barrier through L2,
// session r2w; not kept.)
__device__ __forceinline__ void level_barrier(unsigned long long* bar, unsigned long long target) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(bar, 1ull);
    const unsigned long long t0 = globaltimer_ns();
    unsigned int spins = 0;
    while (ld_volatile_u64(bar) < target)
      if ((++spins & 0x3ffu) == 0 && globaltimer_ns() - t0 > kPeerTimeoutNs) asm volatile("trap;");
  }
  __syncthreads();
}

// What a thread can fetch of its next row BEFORE the barrier that releases the level: everything but the entries of the
// solution vector - row number, extent, the first kPre column indices and factors, the right-hand side / pivot.  The
// loads then fly during the barrier wait and the dependent chain behind the barrier shrinks from
// order -> row_ptr -> (col, pc) -> y to the y loads alone.
constexpr int kPre = 16;
struct RowPre {
  int32_t i, k0, re, dp;   // row (-1: none), first entry of interest, end of the row, diagonal position
  int32_t j[kPre];
  double a[kPre];
  double rhs, d;           // lower solve: x_i and (non-unit) the pivot; upper solve: the pivot
};

template <bool kLower>
__device__ __forceinline__ void row_prefetch(RowPre& q, const int32_t* __restrict__ order, int64_t t, int64_t count,
                                             const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                             const int32_t* __restrict__ diag_pos, const double* __restrict__ pc,
                                             const double* __restrict__ x, int unit_lower) {
  q.i = -1;
  if (t >= count) return;
  const int32_t i = order[t];
  q.i = i;
  const int32_t rb = row_ptr[i];
  q.re = row_ptr[i + 1];
  q.dp = diag_pos[i];
  q.k0 = kLower ? rb : (q.dp >= 0 ? q.dp + 1 : rb);
#pragma unroll
  for (int u = 0; u < kPre; u++) {
    const bool in = q.k0 + u < q.re;
    q.j[u] = in ? col[q.k0 + u] : i;
    q.a[u] = in ? pc[q.k0 + u] : 0.0;
  }
  q.rhs = kLower ? x[i] : 0.0;
  q.d = (kLower && unit_lower) ? 1.0 : (q.dp >= 0 ? pc[q.dp] : 0.0);
}

// The row itself, after the barrier: the same subtractions in the same (ascending column) order as LowerRow / UpperRow.
template <bool kLower>
__device__ __forceinline__ void row_finish(const RowPre& q, const int32_t* __restrict__ col, const double* __restrict__ pc,
                                           const double* in_vec /* lower: y (own output); upper: z */, const double* y,
                                           double* out, int unit_lower, int32_t* flag) {
  if (q.i < 0) return;
  const int32_t i = q.i;
  double v[kPre];
#pragma unroll
  for (int u = 0; u < kPre; u++) v[u] = (kLower ? q.j[u] < i : q.j[u] > i) ? dev::ld_l2(in_vec + q.j[u]) : 0.0;
  double acc = kLower ? q.rhs : dev::ld_l2(y + i);
#pragma unroll
  for (int u = 0; u < kPre; u++)
    if (kLower ? q.j[u] < i : q.j[u] > i) acc = dev::sub_rn(acc, dev::mul_rn(q.a[u], v[u]));
  // rows with more than kPre entries of interest: the rest with plain dependent loads
  if (kLower ? q.j[kPre - 1] < i : q.k0 + kPre < q.re) {
    for (int32_t k = q.k0 + kPre; k < q.re; k++) {
      const int32_t j = col[k];
      if (kLower && j >= i) break;
      if (kLower || j > i) acc = dev::sub_rn(acc, dev::mul_rn(pc[k], dev::ld_l2(in_vec + j)));
    }
  }
  if (!(kLower && unit_lower) && q.d == 0.0) dev::atomic_or_i32(flag, 1);
  out[i] = acc / q.d;
}

__global__ void __launch_bounds__(kT)
ilu_levels_kernel(precond::LowerRow lo, precond::UpperRow up, const int64_t* __restrict__ ptr_l, int nlev_l,
                  const int64_t* __restrict__ ptr_u, int nlev_u, unsigned long long* bar, unsigned long long base) {
  const int64_t gsize = (int64_t)gridDim.x * blockDim.x, gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long target = base;
  const int32_t* order_l = lo.order;
  const int32_t* order_u = up.order;
  RowPre q;
  if (nlev_l > 0) row_prefetch<true>(q, order_l + ptr_l[0], gtid, ptr_l[1] - ptr_l[0], lo.row_ptr, lo.col, lo.diag_pos, lo.pc, lo.x, lo.unit_lower);
  for (int l = 0; l < nlev_l; l++) {
    const int64_t b = ptr_l[l], e = ptr_l[l + 1];
    row_finish<true>(q, lo.col, lo.pc, lo.y, lo.y, lo.y, lo.unit_lower, lo.flag);
    lo.order = order_l + b;
    for (int64_t t = gtid + gsize; t < e - b; t += gsize) lo(t);  // levels wider than the grid: the generic row functor
    if (l + 1 < nlev_l) row_prefetch<true>(q, order_l + e, gtid, ptr_l[l + 2] - e, lo.row_ptr, lo.col, lo.diag_pos, lo.pc, lo.x, lo.unit_lower);
    else if (nlev_u > 0) row_prefetch<false>(q, order_u + ptr_u[0], gtid, ptr_u[1] - ptr_u[0], up.row_ptr, up.col, up.diag_pos, up.pc, nullptr, 0);
    target += gridDim.x;
    level_barrier(bar, target);
  }
  if (nlev_l == 0 && nlev_u > 0)
    row_prefetch<false>(q, order_u + ptr_u[0], gtid, ptr_u[1] - ptr_u[0], up.row_ptr, up.col, up.diag_pos, up.pc, nullptr, 0);
  for (int l = 0; l < nlev_u; l++) {
    const int64_t b = ptr_u[l], e = ptr_u[l + 1];
    row_finish<false>(q, up.col, up.pc, up.z, up.y, up.z, 0, up.flag);
    up.order = order_u + b;
    for (int64_t t = gtid + gsize; t < e - b; t += gsize) up(t);
    if (l + 1 < nlev_u) {
      row_prefetch<false>(q, order_u + e, gtid, ptr_u[l + 2] - e, up.row_ptr, up.col, up.diag_pos, up.pc, nullptr, 0);
      target += gridDim.x;
      level_barrier(bar, target);
    }
  }
}

int vgrid(const cask_b200_ctx* ctx, int64_t n) {
  const int64_t ctas = (n + kT - 1) / kT;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)ctx->sm_count * 4));
}

PrecondState* state(cask_b200_ctx* ctx) {
  if (!ctx->precond) ctx->precond = new PrecondState();
  return ctx->precond;
}

struct MatrixView {
  int64_t n = 0, nnz = 0;
  const int32_t* rp = nullptr;
  const int32_t* ci = nullptr;
  const double* va = nullptr;
};

// the matrix the preconditioner is built from: the override if one is set, else the matrix given to preprocess
int precond_matrix(cask_b200_ctx* ctx, MatrixView* mv) {
  const Plan& pl = ctx->plan;
  if (dist_active(ctx)) return fail(CASK_B200_ERR_UNSUPPORTED, "pcg with a preconditioner is single-rank");
  if (pl.n != pl.m) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: matrix must be square");
  PrecondState* st = state(ctx);
  if (st->pc_matrix) {
    int64_t n = 0, m = 0, nnz = 0;
    CB_TRY(cask_b200_csr_get_info(st->pc_matrix, &n, &m, &nnz, nullptr));
    if (n != pl.n || m != pl.m) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: the preconditioner matrix has other dimensions than A");
    mv->n = n;
    mv->nnz = nnz;
    return cask_b200_csr_device_arrays(st->pc_matrix, &mv->rp, &mv->ci, &mv->va);
  }
  mv->n = pl.n;
  mv->nnz = pl.nnz;
  mv->rp = pl.d_row_ptr;
  mv->ci = pl.d_col;
  mv->va = pl.d_val;
  return CASK_B200_OK;
}

void drop_ilu_graph(PrecondState* st) {
  if (st->ilu_graph) cudaGraphExecDestroy(st->ilu_graph);
  st->ilu_graph = nullptr;
  st->graph_r = nullptr;
  st->graph_z = nullptr;
  st->graph_unit = -1;
}

int ensure_ilu(cask_b200_ctx* ctx) {
  PrecondState* st = state(ctx);
  if (st->ilu_ready) return CASK_B200_OK;
  drop_ilu_graph(st);
  MatrixView mv;
  CB_TRY(precond_matrix(ctx, &mv));
  dev::Exec ex = dev::exec_of(ctx);
  CB_TRY(precond::ilu_analyse(ex, mv.n, mv.nnz, mv.rp, mv.ci, &st->ilu));
  CB_TRY(precond::ilu_factor(ex, mv.va, &st->ilu));
  // level boundaries for the persistent solve kernel
  cudaFree(st->d_level_ptr);
  st->d_level_ptr = nullptr;
  std::vector<int64_t> lp(st->ilu.ptr_l);
  lp.insert(lp.end(), st->ilu.ptr_u.begin(), st->ilu.ptr_u.end());
  st->max_level_rows = 0;
  for (size_t l = 0; l + 1 < st->ilu.ptr_l.size(); l++) st->max_level_rows = std::max(st->max_level_rows, st->ilu.ptr_l[l + 1] - st->ilu.ptr_l[l]);
  for (size_t l = 0; l + 1 < st->ilu.ptr_u.size(); l++) st->max_level_rows = std::max(st->max_level_rows, st->ilu.ptr_u[l + 1] - st->ilu.ptr_u[l]);
  CB_CUDA(cudaMalloc(&st->d_level_ptr, sizeof(int64_t) * std::max<size_t>(lp.size(), 1)));
  if (!lp.empty()) CB_CUDA(cudaMemcpyAsync(st->d_level_ptr, lp.data(), sizeof(int64_t) * lp.size(), cudaMemcpyHostToDevice, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!st->d_level_bar) {
    CB_CUDA(cudaMalloc(&st->d_level_bar, sizeof(unsigned long long)));
    CB_CUDA(cudaMemset(st->d_level_bar, 0, sizeof(unsigned long long)));
    st->level_bar_base = 0;
  }
  st->ilu_ready = true;
  return CASK_B200_OK;
}

int ensure_jacobi(cask_b200_ctx* ctx) {
  PrecondState* st = state(ctx);
  if (st->jacobi_ready) return CASK_B200_OK;
  MatrixView mv;
  CB_TRY(precond_matrix(ctx, &mv));
  cudaFree(st->invd);
  st->invd = nullptr;
  CB_CUDA(cudaMalloc(&st->invd, sizeof(double) * (size_t)std::max<int64_t>(mv.n, 1)));
  dev::Exec ex = dev::exec_of(ctx);
  precond::InvDiag f{mv.rp, mv.ci, mv.va, st->invd};
  CB_TRY(dev::for_each(ex, mv.n, f));
  st->jacobi_ready = true;
  return CASK_B200_OK;
}

int ensure_vectors(cask_b200_ctx* ctx, int64_t n) {
  PrecondState* st = state(ctx);
  if (st->vec_len < n || !st->vec[0]) {
    for (auto& v : st->vec) { cudaFree(v); v = nullptr; }
    st->vec_len = 0;
    for (auto& v : st->vec) CB_CUDA(cudaMalloc(&v, sizeof(double) * (size_t)std::max<int64_t>(n, 2)));
    st->vec_len = n;
  }
  const int np = vgrid(ctx, n);
  if (st->nparts < np) {
    cudaFree(st->partials);
    st->partials = nullptr;
    st->nparts = 0;
    CB_CUDA(cudaMalloc(&st->partials, sizeof(double) * (size_t)np));
    st->nparts = np;
  }
  if (!st->scal) CB_CUDA(cudaMalloc(&st->scal, sizeof(double) * 4));
  return CASK_B200_OK;
}

int dot(cask_b200_ctx* ctx, int64_t n, const double* a, const double* b, double* d_out) {
  PrecondState* st = state(ctx);
  const int g = vgrid(ctx, n);
  pcg_dot_kernel<<<g, kT, 0, ctx->stream>>>(n, a, b, st->partials);
  pcg_finish_kernel<<<1, kT, 0, ctx->stream>>>(st->partials, g, d_out);
  ctx->launches += 2;
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

// z = M^-1 r
int apply(cask_b200_ctx* ctx, int32_t precon, int64_t n, const double* r, double* z) {
  PrecondState* st = state(ctx);
  if (precon == CASK_B200_PRECON_IDENTITY) {
    CB_CUDA(cudaMemcpyAsync(z, r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    return CASK_B200_OK;
  }
  if (precon == CASK_B200_PRECON_JACOBI) {
    pcg_scale_kernel<<<vgrid(ctx, n), kT, 0, ctx->stream>>>(n, st->invd, r, z);
    ctx->launches++;
    CB_CUDA(cudaGetLastError());
    return CASK_B200_OK;
  }
  dev::Exec ex = dev::exec_of(ctx);
  const int unit = precon == CASK_B200_PRECON_ILU_UNIT ? 1 : 0;
  const size_t levels = st->ilu.ptr_l.size() + st->ilu.ptr_u.size();
  cudaStream_t s = ctx->stream;
  if (ctx->ilu_persistent != 0 && levels > 16 && st->ilu.factored && st->d_level_ptr) {
    const int nl = (int)st->ilu.ptr_l.size() - 1, nu = (int)st->ilu.ptr_u.size() - 1;
    precond::LowerRow lo{st->ilu.row_ptr, st->ilu.col, st->ilu.diag_pos, st->ilu.order_l, st->ilu.pc, r, st->ilu.y, unit, st->ilu.flag};
    precond::UpperRow up{st->ilu.row_ptr, st->ilu.col, st->ilu.diag_pos, st->ilu.order_u, st->ilu.pc, st->ilu.y, z, st->ilu.flag};
    const int64_t* pl = st->d_level_ptr;
    const int64_t* pu = st->d_level_ptr + st->ilu.ptr_l.size();
    cudaLaunchConfig_t cfg = {};
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // one cooperative launch; the grid covers the widest level once (more CTAs only make the barrier dearer)
    int occ = 0;
    CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ilu_levels_kernel, kT, 0));
    const int64_t want = std::max<int64_t>(1, (st->max_level_rows + kT - 1) / kT);
    const int grid = (int)std::min<int64_t>(want, (int64_t)std::max(occ, 1) * ctx->sm_count);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kT);
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, ilu_levels_kernel, lo, up, pl, nl, pu, nu, st->d_level_bar, st->level_bar_base);
    if (e == cudaSuccess) st->level_bar_base += (unsigned long long)grid * (unsigned long long)(nl + std::max(nu - 1, 0));
    if (e == cudaSuccess) {
      ctx->launches++;
      return CASK_B200_OK;
    }
    cudaGetLastError();  // cooperative launch not available here: the graph / per-level paths below
  }
  // stream capture needs a real stream (not the legacy default one a caller may have handed over with set_stream(NULL))
  const bool graphable = ctx->ilu_graph != 0 && levels > 16 && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
  if (!graphable) return precond::ilu_apply(ex, &st->ilu, unit, r, z, nullptr);
  if (!st->ilu_graph || st->graph_r != r || st->graph_z != z || st->graph_unit != unit) {
    drop_ilu_graph(st);
    int64_t counted = 0;
    dev::Exec cap = ex;
    cap.launches = &counted;
    CB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = precond::ilu_apply(cap, &st->ilu, unit, r, z, nullptr);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &g);  // always ends the capture, also after a failed launch
    if (rc != CASK_B200_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    CB_CUDA(e);
    const cudaError_t ei = cudaGraphInstantiate(&st->ilu_graph, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) {
      st->ilu_graph = nullptr;
      return fail(CASK_B200_ERR_CUDA, std::string("cudaGraphInstantiate (ILU levels): ") + cudaGetErrorString(ei));
    }
    st->graph_r = r; st->graph_z = z; st->graph_unit = unit; st->graph_nodes = counted;
  }
  CB_CUDA(cudaGraphLaunch(st->ilu_graph, s));
  ctx->launches += st->graph_nodes;
  return CASK_B200_OK;
}

}  // namespace
}  // namespace caskb200

using namespace caskb200;

extern "C" {

int cask_b200_precond_set_matrix(cask_b200_ctx* ctx, const cask_b200_csr* csr) {
  if (!ctx) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "precond_set_matrix: null context");
  PrecondState* st = state(ctx);
  st->pc_matrix = csr;
  drop_ilu_graph(st);
  st->ilu_ready = false;
  st->jacobi_ready = false;
  return CASK_B200_OK;
}

int cask_b200_ilu_factor(cask_b200_ctx* ctx, double* pc, int32_t* levels_lower, int32_t* levels_upper) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ilu_factor: preprocess a matrix first");
  CB_TRY(ensure_device(ctx));
  CB_TRY(ensure_ilu(ctx));
  PrecondState* st = state(ctx);
  if (levels_lower) *levels_lower = (int32_t)st->ilu.ptr_l.size() - 1;
  if (levels_upper) *levels_upper = (int32_t)st->ilu.ptr_u.size() - 1;
  if (pc && st->ilu.nnz) {
    dev::Exec ex = dev::exec_of(ctx);
    CB_TRY(dev::download(ex, pc, st->ilu.pc, sizeof(double) * (size_t)st->ilu.nnz));
  }
  return CASK_B200_OK;
}

int cask_b200_ilu_apply(cask_b200_ctx* ctx, int32_t unit_lower, const double* x, double* z, int32_t* zero_pivot) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ilu_apply: preprocess a matrix first");
  if (!x || !z) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ilu_apply: null vector");
  CB_TRY(ensure_device(ctx));
  CB_TRY(ensure_ilu(ctx));
  PrecondState* st = state(ctx);
  const int64_t n = st->ilu.n;
  CB_TRY(ensure_vectors(ctx, n));
  dev::Exec ex = dev::exec_of(ctx);
  CB_TRY(dev::upload(ex, st->vec[0], x, sizeof(double) * (size_t)n));
  CB_TRY(dev::zero(ex, st->ilu.flag, sizeof(int32_t)));
  int32_t zp = 0;
  CB_TRY(precond::ilu_apply(ex, &st->ilu, unit_lower ? 1 : 0, st->vec[0], st->vec[3], &zp));
  CB_TRY(dev::download(ex, z, st->vec[3], sizeof(double) * (size_t)n));
  if (zero_pivot) *zero_pivot = zp;
  return CASK_B200_OK;
}

int cask_b200_pcg_device(cask_b200_ctx* ctx, const double* d_rhs, double* d_x, int32_t maxiters, double tol, int32_t precon,
                         int32_t* iterations, int32_t* converged, double* rs_final) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: preprocess a matrix first");
  if (!d_rhs || !d_x) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: null vector");
  if (precon < CASK_B200_PRECON_IDENTITY || precon > CASK_B200_PRECON_ILU_UNIT)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: unknown preconditioner");
  CB_TRY(ensure_device(ctx));
  const Plan& pl = ctx->plan;
  if (dist_active(ctx)) return fail(CASK_B200_ERR_UNSUPPORTED, "pcg with a preconditioner is single-rank");
  if (pl.n != pl.m) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: matrix must be square");
  const int64_t n = pl.n;
  if (precon == CASK_B200_PRECON_ILU || precon == CASK_B200_PRECON_ILU_UNIT) CB_TRY(ensure_ilu(ctx));   // :171  Precon precon{a}
  if (precon == CASK_B200_PRECON_JACOBI) CB_TRY(ensure_jacobi(ctx));
  CB_TRY(ensure_vectors(ctx, n));
  PrecondState* st = state(ctx);
  double *r = st->vec[0], *p = st->vec[1], *Ap = st->vec[2], *z = st->vec[3];
  cudaStream_t s = ctx->stream;
  const int vg = vgrid(ctx, n);
  const size_t nb = sizeof(double) * (size_t)n;
  if (st->ilu.flag) CB_CUDA(cudaMemsetAsync(st->ilu.flag, 0, sizeof(int32_t), s));

  // r = b - A x (x staged through p: the SpMV wants a 16-byte aligned operand)      :189-190
  CB_CUDA(cudaMemcpyAsync(p, d_x, nb, cudaMemcpyDeviceToDevice, s));
  CB_TRY(launch_spmv(ctx, p, r, 0, s, nullptr));
  pcg_residual_kernel<<<vg, kT, 0, s>>>(n, d_rhs, r);
  ctx->launches++;
  CB_TRY(apply(ctx, precon, n, r, z));                                                // :193
  CB_CUDA(cudaMemcpyAsync(p, z, nb, cudaMemcpyDeviceToDevice, s));                    // :195
  CB_TRY(dot(ctx, n, r, z, st->scal + 1));                                            // :198
  double rsold = 0.0, rsnew = 0.0;
  CB_CUDA(cudaMemcpyAsync(&rsold, st->scal + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  rsnew = rsold;
  int conv = 0;
  for (int32_t it = 0; it < maxiters; it++) {                                         // :200
    CB_TRY(launch_spmv(ctx, p, Ap, 0, s, nullptr));                                   // :206
    CB_TRY(dot(ctx, n, p, Ap, st->scal + 0));                                         // :208
    pcg_xr_kernel<<<vg, kT, 0, s>>>(n, rsold, st->scal + 0, p, Ap, d_x, r);           // :208-212
    ctx->launches++;
    CB_TRY(apply(ctx, precon, n, r, z));                                              // :215
    CB_TRY(dot(ctx, n, r, z, st->scal + 1));                                          // :218
    CB_CUDA(cudaMemcpyAsync(&rsnew, st->scal + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (rsnew <= tol * tol) { conv = 1; break; }                                      // :220-226
    pcg_p_kernel<<<vg, kT, 0, s>>>(n, rsnew / rsold, z, p);                           // :229
    ctx->launches++;
    rsold = rsnew;                                                                    // :230
    if (iterations) *iterations = it;                                                 // :231
  }
  CB_CUDA(cudaGetLastError());
  if (converged) *converged = conv;
  if (rs_final) *rs_final = rsnew;
  if (precon == CASK_B200_PRECON_ILU || precon == CASK_B200_PRECON_ILU_UNIT) {
    int32_t zp = 0;
    CB_CUDA(cudaMemcpyAsync(&zp, st->ilu.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (zp) return fail(CASK_B200_ERR_RUNTIME, "pcg: the ILU solve met a zero or missing pivot");
  }
  return CASK_B200_OK;
}

int cask_b200_pcg(cask_b200_ctx* ctx, const double* rhs, double* x, int32_t maxiters, double tol, int32_t precon,
                  int32_t* iterations, int32_t* converged, double* rs_final) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: preprocess a matrix first");
  if (!rhs || !x) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pcg: null vector");
  CB_TRY(ensure_device(ctx));
  const int64_t n = ctx->plan.n;
  double *d_b = nullptr, *d_x = nullptr;
  CB_CUDA(cudaMalloc(&d_b, sizeof(double) * (size_t)std::max<int64_t>(n, 2)));
  if (cudaMalloc(&d_x, sizeof(double) * (size_t)std::max<int64_t>(n, 2)) != cudaSuccess) {
    cudaFree(d_b);
    return fail(CASK_B200_ERR_CUDA, "pcg: out of device memory");
  }
  int rc = CASK_B200_OK;
  if (cudaMemcpyAsync(d_b, rhs, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    rc = fail(CASK_B200_ERR_CUDA, "pcg: upload failed");
  if (rc == CASK_B200_OK) rc = cask_b200_pcg_device(ctx, d_b, d_x, maxiters, tol, precon, iterations, converged, rs_final);
  if (rc == CASK_B200_OK || rc == CASK_B200_ERR_RUNTIME) {
    // the reference leaves x where the loop stopped, whatever the outcome
    if (cudaMemcpyAsync(x, d_x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      rc = fail(CASK_B200_ERR_CUDA, "pcg: download failed");
  }
  cudaFree(d_b);
  cudaFree(d_x);
  return rc;
}

}  // extern "C"
