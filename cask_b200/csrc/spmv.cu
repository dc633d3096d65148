// fp64 SpMV kernels for sm_100a.  Replaces the device side of
// cask::spmv::Spmv::spmv (src/runtime/Spmv.cpp:185-328; dataflow kernels src/spmv/src/SpmvKernel.java).
//
//  spmv_ell_staged_kernel  one CTA per staged-ELL slice.  The slice's x windows ("runs") are staged
//                          into shared memory by TMA bulk copies (cp.async.bulk -> SASS UBLKCP)
//                          completing on an mbarrier — the B200 analogue of CASK's on-chip vector
//                          cache (SpmvCacheKernel, SpmvKernel.java:131-151).  Each thread owns four
//                          rows, streams their values with 128-bit loads and 16-bit cache indices
//                          with 64-bit loads, double-buffered in registers, and sums every row in
//                          ascending column order with separate multiply and add — bit-identical to
//                          the reference's CsrMatrix::dot.
//  spmv_csr_vec_kernel     one CTA per gather-CSR slice, VEC lanes per row (2..32 from the row-length
//                          histogram), x gathered through the read-only path, warp-shuffle reduction.
//
// Both can fuse the per-CTA partial of dot(y, w) (CG's p.Ap) into their epilogue.
#include "ctx.cuh"

namespace caskb200 {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// sum over the CTA, result valid in thread 0
__device__ __forceinline__ double cta_sum_d(double v, double* red) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
  return t;
}

constexpr int RPT = kEllRowsPerThread;  // 4 rows per thread
constexpr int KU = 4;                   // ELL columns per register buffer

struct EllChunk {
  double2 v[KU][RPT / 2];
  uint2 ix[KU];
};

__device__ __forceinline__ void ell_load(EllChunk& c, const double* __restrict__ vb, const uint16_t* __restrict__ ib,
                                         int k0, int L) {
#pragma unroll
  for (int u = 0; u < KU; u++) {
    if (k0 + u < L) {
      const size_t off = (size_t)(k0 + u) * (kEllThreads * RPT);
      const double2* vp = reinterpret_cast<const double2*>(vb + off);
      c.v[u][0] = __ldcs(vp);          // streamed once: evict-first
      c.v[u][1] = __ldcs(vp + 1);
      c.ix[u] = __ldcs(reinterpret_cast<const uint2*>(ib + off));
    } else {
      c.v[u][0] = make_double2(0.0, 0.0);
      c.v[u][1] = make_double2(0.0, 0.0);
      c.ix[u] = make_uint2(0u, 0u);     // zero slot
    }
  }
}

__device__ __forceinline__ void ell_accumulate(const EllChunk& c, const double* __restrict__ xs, double (&acc)[RPT]) {
#pragma unroll
  for (int u = 0; u < KU; u++) {
    // separate multiply and add, ascending column order: the summation order of DokMatrix::dot
    // (src/runtime/SparseMatrix.hpp:255-264)
    acc[0] = __dadd_rn(acc[0], __dmul_rn(c.v[u][0].x, xs[c.ix[u].x & 0xffffu]));
    acc[1] = __dadd_rn(acc[1], __dmul_rn(c.v[u][0].y, xs[c.ix[u].x >> 16]));
    acc[2] = __dadd_rn(acc[2], __dmul_rn(c.v[u][1].x, xs[c.ix[u].y & 0xffffu]));
    acc[3] = __dadd_rn(acc[3], __dmul_rn(c.v[u][1].y, xs[c.ix[u].y >> 16]));
  }
}

template <bool kDot>
__global__ void __launch_bounds__(kEllThreads, 2)
spmv_ell_staged_kernel(const SliceDesc* __restrict__ slices, const int32_t* __restrict__ list,
                       const Run* __restrict__ runs, const double* __restrict__ ell_vals,
                       const uint16_t* __restrict__ ell_idx, const double* __restrict__ x,
                       double* __restrict__ y, const double* __restrict__ dot_with,
                       double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(smem_raw);
  double* xs = reinterpret_cast<double*>(smem_raw + 16);  // 16-byte aligned: bulk-copy destination
  __shared__ double red[kEllThreads / 32];

  const int tid = threadIdx.x;
  const SliceDesc sd = slices[list[blockIdx.x]];
  const uint32_t bar = smem_u32(bar_ptr);
  const int L = sd.width;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    xs[0] = 0.0;
    xs[1] = 0.0;
  }
  __syncthreads();

  if (tid < 32) {
    // warp 0 stages the x windows: one bulk copy per run, all counted on one mbarrier phase
    const Run* rr = runs + sd.run_off;
    uint32_t bytes = 0;
    for (int i = tid; i < sd.nruns; i += 32) bytes += (uint32_t)(rr[i].len & ~1) * 8u;
#pragma unroll
    for (int d = 16; d; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
    if (tid == 0) mbar_expect_tx(bar, bytes);
    __syncwarp();
    for (int i = tid; i < sd.nruns; i += 32) {
      const Run r = rr[i];
      const uint32_t b = (uint32_t)(r.len & ~1) * 8u;
      if (b) bulk_g2s(smem_u32(xs + r.local_base), x + r.col0, b, bar);
      if (r.len & 1) xs[r.local_base + r.len - 1] = x[r.col0 + r.len - 1];  // odd tail at column m-1
    }
  }

  // stream the slice while the x windows are in flight
  const double* vb = ell_vals + sd.val_off + (size_t)tid * RPT;
  const uint16_t* ib = ell_idx + sd.val_off + (size_t)tid * RPT;
  double acc[RPT] = {0.0, 0.0, 0.0, 0.0};
  EllChunk a, b;
  ell_load(a, vb, ib, 0, L);
  ell_load(b, vb, ib, KU, L);
  __syncthreads();      // zero slots and odd tails are visible
  mbar_wait(bar, 0);    // bulk copies have landed
  for (int k0 = 0; k0 < L; k0 += 2 * KU) {
    ell_accumulate(a, xs, acc);
    ell_load(a, vb, ib, k0 + 2 * KU, L);
    ell_accumulate(b, xs, acc);
    ell_load(b, vb, ib, k0 + 3 * KU, L);
  }

  double dot = 0.0;
#pragma unroll
  for (int j = 0; j < RPT; j++) {
    const int row = j * kEllThreads + tid;
    if (row < sd.nrows) {
      y[sd.row0 + row] = acc[j];
      if (kDot) dot += acc[j] * dot_with[sd.row0 + row];
    }
  }
  if (kDot) {
    const double t = cta_sum_d(dot, red);
    if (tid == 0) partials[blockIdx.x] = t;
  }
}

template <int VEC, bool kDot>
__global__ void __launch_bounds__(256)
spmv_csr_vec_kernel(const SliceDesc* __restrict__ slices, const int32_t* __restrict__ list,
                    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                    const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                    const double* __restrict__ dot_with, double* __restrict__ partials) {
  __shared__ double red[8];
  const SliceDesc sd = slices[list[blockIdx.x]];
  const int lane = threadIdx.x % VEC, sub = threadIdx.x / VEC;
  constexpr int kRowsPerPass = 256 / VEC;
  double dot = 0.0;
  for (int rb = 0; rb < sd.nrows; rb += kRowsPerPass) {
    const int r = rb + sub;
    double acc = 0.0;
    if (r < sd.nrows) {
      const int32_t kb = row_ptr[sd.row0 + r], ke = row_ptr[sd.row0 + r + 1];
      for (int32_t k = kb + lane; k < ke; k += VEC) acc += __ldcs(val + k) * __ldg(x + __ldcs(col + k));
    }
#pragma unroll
    for (int d = VEC / 2; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0 && r < sd.nrows) {
      y[sd.row0 + r] = acc;
      if (kDot) dot += acc * dot_with[sd.row0 + r];
    }
  }
  if (kDot) {
    const double t = cta_sum_d(dot, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
  }
}

template <bool kDot>
void launch_csr(int vec, int grid, cudaStream_t s, const SliceDesc* sl, const int32_t* list, const Plan& p,
                const double* x, double* y, const double* w, double* partials) {
#define CB_CSR(V)                                                                                         \
  spmv_csr_vec_kernel<V, kDot><<<grid, 256, 0, s>>>(sl, list, p.d_row_ptr, p.d_col, p.d_val, x, y, w, partials)
  switch (vec) {
    case 2: CB_CSR(2); break;
    case 4: CB_CSR(4); break;
    case 8: CB_CSR(8); break;
    case 16: CB_CSR(16); break;
    default: CB_CSR(32); break;
  }
#undef CB_CSR
}

}  // namespace

int spmv_num_ctas(cask_b200_ctx* ctx, int part) {
  const Plan& p = ctx->plan;
  if (part == 1) return p.n_ell_interior + p.n_csr_interior;
  if (part == 2) return (p.n_ell - p.n_ell_interior) + (p.n_csr - p.n_csr_interior);
  return p.n_ell + p.n_csr;
}

// part: 0 = every slice, 1 = slices that read only this rank's own x (interior), 2 = the rest.
// With fusion, partials[0 .. spmv_num_ctas(part)) receives one partial dot per CTA.
int launch_spmv(cask_b200_ctx* ctx, const double* d_x, double* d_y, int part, cudaStream_t s,
                const SpmvFusion* fusion) {
  const Plan& p = ctx->plan;
  if ((reinterpret_cast<uintptr_t>(d_x) & 15u) != 0)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "x must be 16-byte aligned (TMA bulk copies)");
  int ell_lo = 0, ell_hi = p.n_ell, csr_lo = 0, csr_hi = p.n_csr;
  if (part == 1) { ell_hi = p.n_ell_interior; csr_hi = p.n_csr_interior; }
  if (part == 2) { ell_lo = p.n_ell_interior; csr_lo = p.n_csr_interior; }
  const bool dot = fusion && fusion->d_dot_with;
  double* partials = dot ? fusion->d_partials : nullptr;
  const double* w = dot ? fusion->d_dot_with : nullptr;
  if (ell_hi > ell_lo) {
    const size_t smem = 16 + sizeof(double) * (size_t)p.max_xcache;
    static thread_local int attr_smem[2] = {0, 0};
    if ((int)smem > attr_smem[dot ? 1 : 0]) {
      if (dot) CB_CUDA(cudaFuncSetAttribute(spmv_ell_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else CB_CUDA(cudaFuncSetAttribute(spmv_ell_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem[dot ? 1 : 0] = (int)smem;
    }
    if (dot)
      spmv_ell_staged_kernel<true><<<ell_hi - ell_lo, kEllThreads, smem, s>>>(
          p.d_slices, p.d_list_ell + ell_lo, p.d_runs, p.d_ell_vals, p.d_ell_idx, d_x, d_y, w, partials);
    else
      spmv_ell_staged_kernel<false><<<ell_hi - ell_lo, kEllThreads, smem, s>>>(
          p.d_slices, p.d_list_ell + ell_lo, p.d_runs, p.d_ell_vals, p.d_ell_idx, d_x, d_y, nullptr, nullptr);
    ctx->launches++;
  }
  if (csr_hi > csr_lo) {
    double* pp = partials ? partials + (ell_hi - ell_lo) : nullptr;
    if (dot) launch_csr<true>(p.csr_vec, csr_hi - csr_lo, s, p.d_slices, p.d_list_csr + csr_lo, p, d_x, d_y, w, pp);
    else launch_csr<false>(p.csr_vec, csr_hi - csr_lo, s, p.d_slices, p.d_list_csr + csr_lo, p, d_x, d_y, nullptr, nullptr);
    ctx->launches++;
  }
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

}  // namespace caskb200
