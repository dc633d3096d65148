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
//  spmv_csr_items_kernel   gather-CSR slices cut into work items of ~8K nonzeros: lanes per row (2..32) from
//                          the item's mean row length, x gathered through the read-only path, warp-shuffle
//                          reduction; long rows get a CTA each and rows beyond 16K nonzeros are split into
//                          segments that a fix-up kernel sums in order (R-MAT hubs).
//
// Both can fuse the per-CTA partial of dot(y, w) (CG's p.Ap) into their epilogue.
#include <algorithm>
#include <cstdlib>

#include "ctx.cuh"
#include "peer.cuh"

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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// same, with an L2 eviction-priority hint: the matrix is streamed once per SpMV and must not push the vectors
// (x, y and the solver's r, p) out of the 126 MB L2
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
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


// ---- persistent, warp-specialised staged-ELL kernel ------------------------------------------------
// One producer warp moves EVERYTHING with TMA bulk copies: the x windows of slice i+1 into the second x
// buffer and the ELL column chunks (KU columns = KU*1024 values + KU*1024 16-bit indices) into a ring of
// `stages` shared-memory buffers.  Eight consumer warps touch shared memory only.  Every buffer has a
// full/empty mbarrier pair, so global loads stay in flight across chunk AND slice boundaries and the
// per-slice chain (descriptor -> runs -> x -> compute) of the one-CTA-per-slice kernel disappears.
constexpr int kConsumerThreads = kEllThreads;           // 256: thread t owns rows t, t+256, t+512, t+768
constexpr int kPersistThreads = kConsumerThreads + 32;  // + the producer warp
constexpr int kMaxStages = 8;
constexpr int kPersistHeaderBytes = 256;

struct PersistHeader {
  uint64_t full_x[2], empty_x[2];
  uint64_t full_c[kMaxStages], empty_c[kMaxStages];
  int32_t meta[2][4];  // per x buffer: row0, nrows, width
};
static_assert(sizeof(PersistHeader) <= kPersistHeaderBytes, "header must fit its reservation");

// kCoded (option value_dict, plan.cu: build_value_dict):
//   1  the value stream is replaced by 8-bit codes into a table of the slice's distinct values.  The table (dict_len
//      doubles) travels with the x windows into the tail of the x buffer, a ring stage holds KU*1024 codes + KU*1024
//      16-bit indices (3 bytes per stored nonzero instead of 10), and a consumer reads its four codes with one 32-bit
//      load and looks the doubles up in shared memory;
//   2  the code names a (value, x-cache displacement) pair: position = displacement + row inside the slice.  The index
//      stream is gone too: a ring stage holds KU*1024 codes, one byte per stored nonzero; the slice's table of dict_len
//      16-byte records {value, displacement in bytes} rides with the x windows.  Code 0 = padding = (+0.0, -1024 rows):
//      every x cache is preceded by 1024 zeros, so a padding entry reads a zero like any other entry reads its x - the
//      inner loop has no special case (value lookup, displacement lookup, one add, x load, multiply, add).
// In both the multiplied doubles, the x entries, the order of the operations and therefore y are those of the uncoded
// kernel, bit for bit.
struct CodedArgs {
  const uint8_t* codes = nullptr;   // same indexing as ell_vals
  const double* dict = nullptr;     // value codes: valuedict::kStride doubles per slice id
  const void* pairs = nullptr;      // pair codes: valuedict::kPairStride 16-byte records per slice id
  int32_t dict_len = 0;             // table entries staged per slice (multiple of 2; of 8 with pair codes)
  int32_t pad_ = 0;
};
constexpr int kDictStride = 256;    // == valuedict::kStride (plan.cu)
constexpr int kPairStride = 256;    // == valuedict::kPairStride

// doubles appended to an x buffer for the slice's table: values (8 B per entry) or pair records (16 B per entry);
// pair codes also put kSliceRows zeros in FRONT of the x cache (what a padding entry reads)
__host__ __device__ constexpr int dict_value_doubles(int dict_len) { return (dict_len + 15) & ~15; }
__host__ __device__ constexpr int dict_pair_doubles(int dict_len) { return (2 * dict_len + 15) & ~15; }

template <int KU, bool kDot, int kCoded>
__global__ void __launch_bounds__(kPersistThreads, 1)
spmv_ell_persistent_kernel(const SliceDesc* __restrict__ slices, const int32_t* __restrict__ list, int count,
                           const Run* __restrict__ runs, const double* __restrict__ ell_vals,
                           const uint16_t* __restrict__ ell_idx, const double* __restrict__ x, double* __restrict__ y,
                           const double* __restrict__ dot_with, double* __restrict__ partials, int xbuf_doubles,
                           int stages, const HaloWait hw, const __grid_constant__ ReduceDesc rd, int keep_i, unsigned long long* trace,
                           const CodedArgs ca, int self_dot) {
  trace_min(trace);
  const bool keep = keep_i != 0;  // vectors fit L2: x windows, y and dot_with are accessed with evict-last
  const unsigned long long keep_policy = l2_policy_evict_last();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PersistHeader* hdr = reinterpret_cast<PersistHeader*>(smem_raw);
  double* xbuf = reinterpret_cast<double*>(smem_raw + kPersistHeaderBytes);
  // ring: [values | indices] per stage; value codes: [indices | codes]; pair codes: [codes] (every region a multiple
  // of 2 KB)
  double* cvals = xbuf + 2 * (size_t)xbuf_doubles;
  uint16_t* cidx = kCoded ? reinterpret_cast<uint16_t*>(cvals)
                          : reinterpret_cast<uint16_t*>(cvals + (size_t)stages * KU * kSliceRows);
  uint8_t* ccode = kCoded == 2 ? reinterpret_cast<uint8_t*>(cvals)
                               : reinterpret_cast<uint8_t*>(cidx + (size_t)stages * KU * kSliceRows);  // coded formats only
  // the table(s) of a slice sit at the tail of its x buffer: values, then (pair codes) the displacements
  // an x buffer: [pair codes: kSliceRows zeros] [x cache: zero slots, staged windows] [table]; positions, and the
  // table offset dict_base, count from the x cache
  constexpr int zpre = kCoded == 2 ? kSliceRows : 0;
  const int dict_base = kCoded == 2 ? xbuf_doubles - zpre - dict_pair_doubles(ca.dict_len)
                        : kCoded == 1 ? xbuf_doubles - dict_value_doubles(ca.dict_len) : 0;
  __shared__ double red[kPersistThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int b = 0; b < 2; b++) {
      mbar_init(smem_u32(&hdr->full_x[b]), 1);
      mbar_init(smem_u32(&hdr->empty_x[b]), kConsumerThreads / 32);
      xbuf[(size_t)b * xbuf_doubles + zpre] = 0.0;      // zero slots: target of every padding entry,
      xbuf[(size_t)b * xbuf_doubles + zpre + 1] = 0.0;  // never overwritten (runs start at local index 2)
    }
    for (int s = 0; s < stages; s++) {
      mbar_init(smem_u32(&hdr->full_c[s]), 1);
      mbar_init(smem_u32(&hdr->empty_c[s]), kConsumerThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (kCoded == 2)  // the zeros a padding entry reads: position = row - kSliceRows, never written again
    for (int i = tid; i < 2 * kSliceRows; i += kPersistThreads)
      xbuf[(size_t)(i / kSliceRows) * xbuf_doubles + (i % kSliceRows)] = 0.0;
  __syncthreads();

  const bool is_producer = warp == kConsumerThreads / 32;
  const uint64_t stream_policy = l2_evict_first_policy();
  auto issue_chunk = [&](const SliceDesc* sdp, int c, int chunk_no) {  // lane 0 of the producer warp
    const int L = sdp->width;
    const int s = chunk_no % stages;
    const int cols = min(KU, L - c * KU);
    const uint32_t fc = smem_u32(&hdr->full_c[s]);
    if constexpr (kCoded == 2) {
      mbar_expect_tx(fc, (uint32_t)cols * kSliceRows);
      const size_t src = (size_t)sdp->val_off + (size_t)c * KU * kSliceRows;
      bulk_g2s_hint(smem_u32(ccode + (size_t)s * KU * kSliceRows), ca.codes + src, (uint32_t)cols * kSliceRows, fc, stream_policy);
    } else if constexpr (kCoded == 1) {
      mbar_expect_tx(fc, (uint32_t)cols * kSliceRows * 3u);
      const size_t src = (size_t)sdp->val_off + (size_t)c * KU * kSliceRows;
      bulk_g2s_hint(smem_u32(ccode + (size_t)s * KU * kSliceRows), ca.codes + src, (uint32_t)cols * kSliceRows, fc, stream_policy);
      bulk_g2s_hint(smem_u32(cidx + (size_t)s * KU * kSliceRows), ell_idx + src, (uint32_t)cols * kSliceRows * 2u, fc, stream_policy);
    } else {
      mbar_expect_tx(fc, (uint32_t)cols * kSliceRows * 10u);
      const size_t src = (size_t)sdp->val_off + (size_t)c * KU * kSliceRows;
      bulk_g2s_hint(smem_u32(cvals + (size_t)s * KU * kSliceRows), ell_vals + src, (uint32_t)cols * kSliceRows * 8u, fc, stream_policy);
      bulk_g2s_hint(smem_u32(cidx + (size_t)s * KU * kSliceRows), ell_idx + src, (uint32_t)cols * kSliceRows * 2u, fc, stream_policy);
    }
  };
  // Programmatic dependent launch: this CTA may be resident while the kernel that produces x is still running.
  // The matrix is static, so the producer warp fills the ring with the first chunks of its first slice BEFORE it
  // waits for the predecessor; only then are x, the solver flags and the peers' halo epochs looked at.
  int prefetched = 0;
  if (is_producer && blockIdx.x < count) {
    const SliceDesc* sdp = slices + list[blockIdx.x];
    prefetched = min((sdp->width + KU - 1) / KU, stages);
    if (lane == 0)
      for (int c = 0; c < prefetched; c++) issue_chunk(sdp, c, c);
  }
  pdl_wait();
  pdl_trigger();
  trace_min(trace ? trace + 1 : nullptr);
  // solver loops enqueue iterations ahead of the convergence test: once the device-side flag is up nothing runs
  if ((hw.skip0 && *hw.skip0) || (hw.skip1 && *hw.skip1)) {
    if (is_producer)
      for (int c = 0; c < prefetched; c++) mbar_wait(smem_u32(&hdr->full_c[c]), 0);  // drain the copies in flight
    return;
  }

  // plain sharded SpMV: the boundary rows of x travel from inside this kernel (peer.cuh: acked_push), by the CTA that
  // has the least to do; everybody else is already streaming interior slices
  if (hw.ack && blockIdx.x == gridDim.x - 1) acked_push(hw, x);
  double dot = 0.0, dot2 = 0.0;  // y . dot_with and (self_dot: BiCGStab's t.t beside t.s) y . y
  if (is_producer) {
    // ===== producer warp =====
    int chunk_no = 0;
    int it = 0;
    bool halo_landed = hw.ctrl == nullptr;
    for (int item = blockIdx.x; item < count; item += gridDim.x, it++) {
      if (!halo_landed && item >= hw.first_item) {
        // first slice that stages x entries owned by a peer: wait until the peers' pushes of this epoch have landed
        if (lane == 0) halo_wait(hw);
        __syncwarp();
        halo_landed = true;
      }
      const SliceDesc* sdp = slices + list[item];
      const int L = sdp->width, nruns = sdp->nruns;
      const int xb = it & 1;
      mbar_wait(smem_u32(&hdr->empty_x[xb]), ((it >> 1) & 1) ^ 1);
      double* xs = xbuf + (size_t)xb * xbuf_doubles + zpre;
      const uint32_t fx = smem_u32(&hdr->full_x[xb]);
      const Run* rr = nruns <= kInlineRuns ? sdp->inl : runs + sdp->run_off;
      uint32_t bytes = 0;
      for (int i = lane; i < nruns; i += 32) {
        const Run r = rr[i];
        const uint32_t b = (uint32_t)(r.len & ~1) * 8u;
        bytes += b;
        if (b) {
          if (keep) bulk_g2s_hint(smem_u32(xs + r.local_base), x + r.col0, b, fx, keep_policy);
          else bulk_g2s(smem_u32(xs + r.local_base), x + r.col0, b, fx);
        }
        if (r.len & 1) xs[r.local_base + r.len - 1] = x[r.col0 + r.len - 1];  // odd tail at column m-1
      }
      if constexpr (kCoded == 1) if (lane == 31) {  // the slice's value table rides on the same barrier phase as its x windows
        const uint32_t b = (uint32_t)ca.dict_len * 8u;
        bytes += b;
        bulk_g2s(smem_u32(xs + dict_base), ca.dict + (size_t)list[item] * kDictStride, b, fx);
      }
      if constexpr (kCoded == 2) if (lane == 31) {  // ... or its table of {value, displacement} records
        const uint32_t b = (uint32_t)ca.dict_len * 16u;
        bytes += b;
        bulk_g2s(smem_u32(xs + dict_base), static_cast<const char*>(ca.pairs) + (size_t)list[item] * kPairStride * 16u, b, fx);
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
      if (lane == 0) {
        hdr->meta[xb][0] = sdp->row0;
        hdr->meta[xb][1] = sdp->nrows;
        hdr->meta[xb][2] = L;
      }
      __syncwarp();
      if (lane == 0) mbar_expect_tx(fx, bytes);  // arrive (release): meta and odd tails become visible with it
      const int nchunks = (L + KU - 1) / KU;
      for (int c = 0; c < nchunks; c++, chunk_no++) {
        if (chunk_no < prefetched) continue;  // already in flight (first slice)
        const int s = chunk_no % stages;
        mbar_wait(smem_u32(&hdr->empty_c[s]), ((chunk_no / stages) & 1) ^ 1);
        if (lane == 0) issue_chunk(sdp, c, chunk_no);
      }
    }
  } else {
    // ===== consumer warps =====
    int chunk_no = 0;
    int it = 0;
    for (int item = blockIdx.x; item < count; item += gridDim.x, it++) {
      const int xb = it & 1;
      mbar_wait(smem_u32(&hdr->full_x[xb]), (it >> 1) & 1);
      const int row0 = hdr->meta[xb][0], nrows = hdr->meta[xb][1], L = hdr->meta[xb][2];
      const double* xs = xbuf + (size_t)xb * xbuf_doubles + zpre;
      double acc[RPT] = {0.0, 0.0, 0.0, 0.0};
      // fused dot: this thread's four entries of dot_with are requested NOW and used after the slice's last column, so
      // the HBM latency hides behind the column loop instead of stalling the consumers (and, through the full ring, the
      // producer) at the end of every slice - on a 7-point stencil a slice is only ~3 us of streaming
      double wv[RPT] = {0.0, 0.0, 0.0, 0.0};
      if (kDot) {
#pragma unroll
        for (int j = 0; j < RPT; j++) {
          const int row = j * kConsumerThreads + tid;
          if (row < nrows) wv[j] = ld_keep(dot_with + row0 + row, keep_policy, keep);
        }
      }
      const int nchunks = (L + KU - 1) / KU;
      for (int c = 0; c < nchunks; c++, chunk_no++) {
        const int s = chunk_no % stages;
        mbar_wait(smem_u32(&hdr->full_c[s]), (chunk_no / stages) & 1);
        const int cols = min(KU, L - c * KU);
        const double* vb = cvals + (size_t)s * KU * kSliceRows + (size_t)tid * RPT;
        const uint16_t* ib = cidx + (size_t)s * KU * kSliceRows + (size_t)tid * RPT;
        const uint8_t* cb = ccode + (size_t)s * KU * kSliceRows + (size_t)tid * RPT;
        const double* dict = xs + dict_base;
        const char* tab = reinterpret_cast<const char*>(xs + dict_base);  // pair codes: 16-byte records
        const char* xrow = reinterpret_cast<const char*>(xs + tid);       // pair codes: x cache at this thread's row 0
#pragma unroll
        for (int u = 0; u < KU; u++) {
          if (u < cols) {
            if constexpr (kCoded == 2) {
              // byte j of thread t's word = code of entry (k, row j*256 + t); record = {value, (position - row) * 8};
              // a padding entry (code 0) reads one of the zeros in front of the x cache
              const uint32_t cw = *reinterpret_cast<const uint32_t*>(cb + (size_t)u * kSliceRows);
              const char* e0 = tab + ((cw & 0xffu) << 4);
              const char* e1 = tab + ((cw >> 4) & 0xff0u);
              const char* e2 = tab + ((cw >> 12) & 0xff0u);
              const char* e3 = tab + ((cw >> 20) & 0xff0u);
              constexpr int kRowStep = kConsumerThreads * (int)sizeof(double);  // rows j and j+1 of a thread
              const double x0 = *reinterpret_cast<const double*>(xrow + *reinterpret_cast<const int*>(e0 + 8));
              const double x1 = *reinterpret_cast<const double*>(xrow + kRowStep + *reinterpret_cast<const int*>(e1 + 8));
              const double x2 = *reinterpret_cast<const double*>(xrow + 2 * kRowStep + *reinterpret_cast<const int*>(e2 + 8));
              const double x3 = *reinterpret_cast<const double*>(xrow + 3 * kRowStep + *reinterpret_cast<const int*>(e3 + 8));
              // ascending column order, separate multiply and add (DokMatrix::dot, SparseMatrix.hpp:255-264)
              acc[0] = __dadd_rn(acc[0], __dmul_rn(*reinterpret_cast<const double*>(e0), x0));
              acc[1] = __dadd_rn(acc[1], __dmul_rn(*reinterpret_cast<const double*>(e1), x1));
              acc[2] = __dadd_rn(acc[2], __dmul_rn(*reinterpret_cast<const double*>(e2), x2));
              acc[3] = __dadd_rn(acc[3], __dmul_rn(*reinterpret_cast<const double*>(e3), x3));
            } else {
              double2 v0, v1;
              if constexpr (kCoded == 1) {
                // entry (k, row j*256 + t) is byte j of thread t's 32-bit word (little endian), as in the value layout
                const uint32_t cw = *reinterpret_cast<const uint32_t*>(cb + (size_t)u * kSliceRows);
                v0 = make_double2(dict[cw & 0xffu], dict[(cw >> 8) & 0xffu]);
                v1 = make_double2(dict[(cw >> 16) & 0xffu], dict[cw >> 24]);
              } else {
                v0 = *reinterpret_cast<const double2*>(vb + (size_t)u * kSliceRows);
                v1 = *reinterpret_cast<const double2*>(vb + (size_t)u * kSliceRows + 2);
              }
              const uint2 ix = *reinterpret_cast<const uint2*>(ib + (size_t)u * kSliceRows);
              // ascending column order, separate multiply and add (DokMatrix::dot, SparseMatrix.hpp:255-264)
              acc[0] = __dadd_rn(acc[0], __dmul_rn(v0.x, xs[ix.x & 0xffffu]));
              acc[1] = __dadd_rn(acc[1], __dmul_rn(v0.y, xs[ix.x >> 16]));
              acc[2] = __dadd_rn(acc[2], __dmul_rn(v1.x, xs[ix.y & 0xffffu]));
              acc[3] = __dadd_rn(acc[3], __dmul_rn(v1.y, xs[ix.y >> 16]));
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&hdr->empty_c[s]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&hdr->empty_x[xb]));
#pragma unroll
      for (int j = 0; j < RPT; j++) {
        const int row = j * kConsumerThreads + tid;
        if (row < nrows) {
          st_keep(y + row0 + row, acc[j], keep_policy, keep);
          if (kDot) {
            dot += acc[j] * wv[j];
            if (self_dot) dot2 += acc[j] * acc[j];
          }
        }
      }
    }
  }
  if (kDot) {
    const double t = cta_sum_d(dot, red);  // deterministic: fixed slice->CTA map, fixed order inside the CTA
    double t2 = 0.0;
    if (self_dot) {
      __syncthreads();  // red[] is reused
      t2 = cta_sum_d(dot2, red);
    }
    if (rd.partials) grid_finish_reduce(rd, t, t2, blockIdx.x, gridDim.x);  // last CTA: sum (+ all-reduce) in place
    else if (tid == 0) partials[blockIdx.x] = t;
  }
  trace_max(trace ? trace + 2 : nullptr);
}

// ---- gather-CSR path: irregular rows, x gathered through the read-only path / L2 -----------------------
// One CTA per CsrItem.  Multi-row items: `vec` lanes per row (chosen per item from its mean row length),
// warp-shuffle reduction.  Single-row items (long rows and segments of split rows): the whole CTA strides
// over the row with four independent accumulators per thread, then a CTA reduction.
// kStream ("CSR-stream"): the multi-row branch first turns the item's contiguous nonzero range into products -
// values and column indices read with perfectly coalesced streaming loads, eight independent x gathers in flight per
// thread whatever the row structure - parks them in shared memory, and only then reduces row by row out of shared
// memory.  The gather, which bounds power-law matrices, no longer waits on the row-length distribution.
template <bool kDot, bool kStream>
__global__ void __launch_bounds__(256)
spmv_csr_items_kernel(const CsrItem* __restrict__ items, const int32_t* __restrict__ row_ptr,
                      const int32_t* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                      double* __restrict__ y, const double* __restrict__ dot_with, double* __restrict__ partials,
                      double* __restrict__ scratch) {
  __shared__ double red[8];
  extern __shared__ __align__(16) double prod[];  // kStream: products of the item's nonzeros
  const CsrItem it = items[blockIdx.x];
  double dot = 0.0;
  if (it.vec == 0) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int32_t k = it.k_lo + threadIdx.x;
    for (; k + 768 < it.k_hi; k += 1024) {
      a0 += __ldcs(val + k) * __ldg(x + __ldcs(col + k));
      a1 += __ldcs(val + k + 256) * __ldg(x + __ldcs(col + k + 256));
      a2 += __ldcs(val + k + 512) * __ldg(x + __ldcs(col + k + 512));
      a3 += __ldcs(val + k + 768) * __ldg(x + __ldcs(col + k + 768));
    }
    for (; k < it.k_hi; k += 256) a0 += __ldcs(val + k) * __ldg(x + __ldcs(col + k));
    const double t = cta_sum_d((a0 + a1) + (a2 + a3), red);
    if (threadIdx.x == 0) {
      if (it.scratch >= 0) scratch[it.scratch] = t;
      else {
        y[it.row0] = t;
        if (kDot) dot = t * dot_with[it.row0];
      }
      if (kDot) partials[blockIdx.x] = dot;
    }
    return;
  }
  const int vec = it.vec;
  const int lane = threadIdx.x & (vec - 1), sub = threadIdx.x / vec;
  const int rows_per_pass = 256 / vec;
  if (kStream) {
    const int32_t k_lo = it.k_lo, cnt = it.k_hi - it.k_lo;
    int32_t* rps = reinterpret_cast<int32_t*>(prod + cnt + (cnt & 1));  // the item's row pointers, rebased
    for (int r = threadIdx.x; r <= it.nrows; r += 256) rps[r] = row_ptr[it.row0 + r] - k_lo;
    for (int32_t j0 = 0; j0 < cnt; j0 += 256 * 8) {
      int32_t c[8];
      double v[8], xv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int32_t j = j0 + u * 256 + (int32_t)threadIdx.x;
        c[u] = j < cnt ? __ldcs(col + k_lo + j) : 0;
        v[u] = j < cnt ? __ldcs(val + k_lo + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int32_t j = j0 + u * 256 + (int32_t)threadIdx.x;
        if (j < cnt) prod[j] = v[u] * xv[u];
      }
    }
    __syncthreads();
    if (vec <= 8) {
      // short rows (and the many empty ones of a power-law matrix): a thread per row, no global loads left
      for (int r = threadIdx.x; r < it.nrows; r += 256) {
        double acc = 0.0;
        for (int32_t k = rps[r]; k < rps[r + 1]; k++) acc += prod[k];
        y[it.row0 + r] = acc;
        if (kDot) dot += acc * dot_with[it.row0 + r];
      }
    } else {
      for (int rb = 0; rb < it.nrows; rb += rows_per_pass) {
        const int r = rb + sub;
        double acc = 0.0;
        if (r < it.nrows)
          for (int32_t k = rps[r] + lane; k < rps[r + 1]; k += vec) acc += prod[k];
        for (int d = vec >> 1; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0 && r < it.nrows) {
          y[it.row0 + r] = acc;
          if (kDot) dot += acc * dot_with[it.row0 + r];
        }
      }
    }
  } else {
    for (int rb = 0; rb < it.nrows; rb += rows_per_pass) {
      const int r = rb + sub;
      double acc = 0.0;
      if (r < it.nrows) {
        const int32_t kb = row_ptr[it.row0 + r], ke = row_ptr[it.row0 + r + 1];
        for (int32_t k = kb + lane; k < ke; k += vec) acc += __ldcs(val + k) * __ldg(x + __ldcs(col + k));
      }
      for (int d = vec >> 1; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if (lane == 0 && r < it.nrows) {
        y[it.row0 + r] = acc;
        if (kDot) dot += acc * dot_with[it.row0 + r];
      }
    }
  }
  if (kDot) {
    const double t = cta_sum_d(dot, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
  }
}

// split rows: partial sums of the segments meet here, in segment order
template <bool kDot>
__global__ void __launch_bounds__(256)
spmv_csr_fixup_kernel(const SplitRow* __restrict__ rows, int count, const double* __restrict__ scratch,
                      double* __restrict__ y, const double* __restrict__ dot_with, double* __restrict__ partial) {
  __shared__ double red[8];
  double dot = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const SplitRow r = rows[i];
    double acc = 0.0;
    for (int g = 0; g < r.nseg; g++) acc += scratch[r.first + g];
    y[r.row] = acc;
    if (kDot) dot += acc * dot_with[r.row];
  }
  if (kDot) {
    const double t = cta_sum_d(dot, red);
    if (threadIdx.x == 0) *partial = t;
  }
}

// ---- gather path, merge-path tiles (default for large irregular matrices: R-MAT, BASELINE configs[2]) -----------
// What bounded the row-group kernel above on R-MAT scale 25 (ncu, profiles/r2a_spmv_csr_rmat_ncu.md): 65 long-scoreboard
// stall cycles per issued instruction at 16 % issue utilisation, DRAM at 45 % - a latency chain row_ptr -> col -> x ->
// shuffle per row with one or two gathers in flight per lane - and the matrix stream fetched 1.4x (8.7 GB of DRAM reads for
// 6.0 GB of values + indices) because vec lanes per row touch every sector of a row twice.  Here a CTA owns one TILE of
// the merge path of (row ends, nonzeros) (plan.cu: build_merge_tiles), i.e. at most 256 x merge_items nonzeros AND as many
// rows whatever the degree distribution, and works in three decoupled phases:
//   A  the tile's nonzeros are read perfectly coalesced (thread t takes nonzeros t, t+256, ...: every sector of the matrix
//      stream is requested exactly once, evict-first), eight column indices and values per thread in flight, THEN the
//      eight x gathers (all independent), then the products go to shared memory;
//   B  every thread walks its merge_items steps of the merge path over shared memory only (products and row ends),
//      emitting finished rows and keeping the partial sum of the row it stops in;
//   C  a segmented scan over the threads' partial sums (keys = row, monotonic) hands each thread the part of its first
//      row that earlier threads summed; rows that finish inside the tile are stored, the tile's last partial row goes to
//      carry[tile], and spmv_csr_merge_fixup_kernel adds the carries of the tiles a long row spans, in tile order.
// Everything is deterministic (fixed tile -> CTA and item -> thread maps, ordered carries); the summation order inside a
// row differs from CsrMatrix::dot, as for every gather kernel (parity bar: 1e-12 relative to sum |a_ij x_j|).
constexpr int kMergeXPastL1 = 1;     // x gathers with ld.global.cg
constexpr int kMergeStreamOnly = 2;  // diagnostic: stop after phase A (what the matrix stream + gathers cost alone)
constexpr int kMergeNoGather = 4;    // diagnostic: x = 1 instead of the gather (what everything but the gathers costs)

// x in the plan's hub-clustered column numbering (plan.cu: build_col_reorder): xp[i] = x[perm[i]] for the referenced
// columns.  perm is ascending inside every run of equal reference counts, so beyond the hubs the reads are a handful of
// interleaved forward sweeps over x; two entries per thread, 16-byte stores.
__global__ void __launch_bounds__(256) permute_x_kernel(const int32_t* __restrict__ perm, const double* __restrict__ x,
                                                        double* __restrict__ xp, int32_t used) {
  const int32_t stride = (int32_t)(gridDim.x * blockDim.x) * 2;
  for (int32_t i = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2; i < used; i += stride) {
    if (i + 1 < used) {
      const int2 q = __ldcs(reinterpret_cast<const int2*>(perm + i));
      *reinterpret_cast<double2*>(xp + i) = make_double2(__ldg(x + q.x), __ldg(x + q.y));
    } else {
      xp[i] = __ldg(x + perm[i]);
    }
  }
}

template <bool kDot, int ITEMS, int CTAS = (ITEMS > 11 ? 4 : 5)>
__global__ void __launch_bounds__(kMergeThreads, CTAS)
spmv_csr_merge_kernel(const MergeTile* __restrict__ tiles, int32_t n_rows, const int32_t* __restrict__ row_ptr,
                      const int32_t* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                      double* __restrict__ y, const double* __restrict__ dot_with, double* __restrict__ partials,
                      double* __restrict__ carry, int flags) {
  constexpr int TILE = kMergeThreads * ITEMS;       // merge items (row ends + nonzeros) per tile
  constexpr int BATCH = ITEMS < 8 ? ITEMS : 8;      // nonzeros per thread in flight in phase A
  extern __shared__ __align__(16) unsigned char merge_smem[];
  double* prod = reinterpret_cast<double*>(merge_smem);           // [TILE] products of the tile's nonzeros
  int32_t* rend = reinterpret_cast<int32_t*>(prod + TILE);        // [TILE + 1] row ends relative to k0
  __shared__ double red[kMergeThreads / 32];
  __shared__ double tail_val[kMergeThreads / 32], pre_val[kMergeThreads / 32];
  __shared__ int32_t tail_key[kMergeThreads / 32], head_key[kMergeThreads / 32], pre_key[kMergeThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MergeTile t = tiles[blockIdx.x];
  const int nnzT = t.k1 - t.k0, nrowsT = t.r1 - t.r0;

  // ---- A: row ends, then products ----
  for (int r = tid; r <= nrowsT; r += kMergeThreads)
    rend[r] = t.r0 + r < n_rows ? row_ptr[t.r0 + r + 1] - t.k0 : 0x7fffffff;  // entry nrowsT: the row the tile stops in
  // the sweep starts on the 32-element boundary below k0: every warp-wide load then covers whole 128-byte lines of the
  // index stream (256 bytes of the value stream) and no sector of the matrix is requested by two warps
  const int lead = t.k0 & 31;
  const int32_t* cp = col + (t.k0 - lead);
  const double* vp = val + (t.k0 - lead);
  for (int j0 = 0; j0 < nnzT + lead; j0 += kMergeThreads * BATCH) {
    int32_t c[BATCH];
    double v[BATCH], xv[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
      const int j = j0 + u * kMergeThreads + tid;
      const bool ok = j >= lead && j < nnzT + lead;
      c[u] = ok ? __ldcs(cp + j) : 0;
      v[u] = ok ? __ldcs(vp + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < BATCH; u++)  // A/B: gathers cached in L2 only; diagnostic: no gather at all
      xv[u] = (flags & kMergeNoGather) ? 1.0 : (flags & kMergeXPastL1) ? __ldcg(x + c[u]) : __ldg(x + c[u]);
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
      const int j = j0 + u * kMergeThreads + tid;
      if (j >= lead && j < nnzT + lead) prod[j - lead] = v[u] * xv[u];
    }
  }
  __syncthreads();
  if (flags & kMergeStreamOnly) {  // diagnostic (CASK_B200_MERGE_DIAG): phase A alone, y is NOT computed
    if (tid == 0) carry[blockIdx.x] = prod[0];
    return;
  }

  // ---- B: this thread's stretch of the merge path ----
  const int total = nnzT + nrowsT;
  const int diag = min(tid * ITEMS, total), diag_end = min(diag + ITEMS, total);
  int lo = max(0, diag - nnzT), hi = min(diag, nrowsT);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rend[mid] <= diag - mid - 1) lo = mid + 1; else hi = mid;
  }
  int r = lo, k = diag - lo;
  const int r_start = r;
  double acc = 0.0, first_val = 0.0, dot = 0.0;
  bool has_first = false;
  // the tile's row 0 continues a row begun in an earlier tile iff the tile does not start on that row's first nonzero
  const bool row0_continues = t.k0 > row_ptr[t.r0];
  int re = rend[r];
  for (int step = diag; step < diag_end; step++) {
    if (k < re) {
      acc += prod[k];
      k++;
    } else {
      if (!has_first) {
        has_first = true;
        first_val = acc;
      } else {
        y[t.r0 + r] = acc;
        if (kDot) dot += acc * dot_with[t.r0 + r];
      }
      acc = 0.0;
      r++;
      re = rend[r];
    }
  }

  // ---- C: segmented inclusive scan of (row r, partial acc) over the threads; rows are non-decreasing in tid ----
  int key = r;
  double sv = acc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int ok = __shfl_up_sync(0xffffffffu, key, d);
    const double ov = __shfl_up_sync(0xffffffffu, sv, d);
    if (lane >= d && ok == key) sv += ov;
  }
  if (lane == 31) { tail_key[warp] = key; tail_val[warp] = sv; }
  if (lane == 0) head_key[warp] = key;
  __syncthreads();
  if (tid == 0) {
    // pre_*[w]: the partial sum (and its row) that the threads of warps < w hand to warp w
    pre_key[0] = -1; pre_val[0] = 0.0;
    for (int w = 1; w < kMergeThreads / 32; w++) {
      double pv = tail_val[w - 1];
      if (head_key[w - 1] == tail_key[w - 1] && pre_key[w - 1] == tail_key[w - 1]) pv += pre_val[w - 1];  // a row spanning whole warps
      pre_key[w] = tail_key[w - 1];
      pre_val[w] = pv;
    }
  }
  __syncthreads();
  if (key == pre_key[warp]) sv += pre_val[warp];
  // what the previous thread hands over: its inclusive sum if it stopped in the row this thread started in
  int prev_key = __shfl_up_sync(0xffffffffu, key, 1);
  double prev_val = __shfl_up_sync(0xffffffffu, sv, 1);
  if (lane == 0) { prev_key = pre_key[warp]; prev_val = pre_val[warp]; }
  if (has_first) {
    const double total_first = prev_key == r_start ? prev_val + first_val : first_val;
    y[t.r0 + r_start] = total_first;
    // a row continued from an earlier tile is finished (and enters the dot product) in the fix-up kernel
    if (kDot && !(r_start == 0 && row0_continues)) dot += total_first * dot_with[t.r0 + r_start];
  }
  if (tid == kMergeThreads - 1) carry[blockIdx.x] = key == nrowsT ? sv : 0.0;
  if (kDot) {
    const double s = cta_sum_d(dot, red);
    if (tid == 0) partials[blockIdx.x] = s;
  }
}

// Rows that span tiles: tile i ends inside row R = r1 iff k1 > row_ptr[R]; the chain of such tiles with the same R is
// summed in tile order by the thread of its first tile and added in front of what the finishing tile stored.
template <bool kDot>
__global__ void __launch_bounds__(256)
spmv_csr_merge_fixup_kernel(const MergeTile* __restrict__ tiles, int count, const int32_t* __restrict__ row_ptr,
                            const double* __restrict__ carry, double* __restrict__ y, const double* __restrict__ dot_with,
                            double* __restrict__ partials) {
  __shared__ double red[8];
  double dot = 0.0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const MergeTile t = tiles[i];
    const bool open = t.k1 > row_ptr[t.r1];
    bool head = open;
    if (open && i > 0) {
      const MergeTile q = tiles[i - 1];
      head = !(q.r1 == t.r1 && q.k1 > row_ptr[q.r1] && q.k1 == t.k0);
    }
    if (head) {
      double sum = carry[i];
      for (int j = i + 1; j < count; j++) {
        const MergeTile u = tiles[j];
        if (u.r1 != t.r1 || u.k0 != tiles[j - 1].k1) break;  // tile j finished the row (or belongs to another run)
        sum += carry[j];
      }
      const double v = sum + y[t.r1];
      y[t.r1] = v;
      if (kDot) dot = v * dot_with[t.r1];
    }
  }
  if (kDot) {
    const double s = cta_sum_d(dot, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
  }
}

}  // namespace

bool spmv_single_launch(const cask_b200_ctx* ctx) {
  const Plan& p = ctx->plan;
  return p.n_csr == 0 && p.n_ell > 0 && ctx->ell_kernel == 1 && p.persist_ku != 0 && (!dist_active(ctx) || peer_ready(ctx));
}

static int ell_grid(const cask_b200_ctx* ctx, int n_slices) {
  const Plan& p = ctx->plan;
  if (ctx->ell_kernel == 1 && p.persist_ku) return std::min(n_slices, ctx->sm_count * p.persist_ctas_per_sm);
  return n_slices;
}

// partial dots written by the gather-CSR part of a launch over list positions [lo, hi): one per item, plus one
// for the fix-up kernel when split rows are involved
static int csr_partials(const Plan& p, int lo, int hi) {
  if (hi <= lo || p.h_item_begin.size() <= (size_t)hi) return 0;
  if (p.csr_merge) {  // one partial per tile + one per CTA of the fix-up kernel
    const int tiles = p.h_item_begin[hi] - p.h_item_begin[lo];
    return tiles + (tiles + 255) / 256;
  }
  return (p.h_item_begin[hi] - p.h_item_begin[lo]) + (p.h_split_begin[hi] > p.h_split_begin[lo] ? 1 : 0);
}

// number of per-CTA partial dots a fused launch over `part` writes
int spmv_num_ctas(cask_b200_ctx* ctx, int part) {
  const Plan& p = ctx->plan;
  if (part == 1) return ell_grid(ctx, p.n_ell_interior) + csr_partials(p, 0, p.n_csr_interior);
  if (part == 2) return ell_grid(ctx, p.n_ell - p.n_ell_interior) + csr_partials(p, p.n_csr_interior, p.n_csr);
  return ell_grid(ctx, p.n_ell) + csr_partials(p, 0, p.n_csr);
}

// Picks KU (ELL columns per ring stage), the ring depth and the CTAs per SM of the persistent kernel from
// the shared memory the plan's largest x cache leaves: two CTAs per SM with >= 3 stages of 2 columns if
// that fits, else one CTA per SM with up to 4 stages of 4 columns.
int configure_persistent(cask_b200_ctx* ctx) {
  Plan& p = ctx->plan;
  p.persist_ku = 0;
  if (p.n_ell == 0) return CASK_B200_OK;
  int dev_smem = 0;
  CB_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  // doubles per x buffer (keeps every buffer 128-B aligned); the coded format appends the slice's value table
  const size_t xbuf = (((size_t)p.max_xcache + 15) & ~(size_t)15) + (p.coded == 1 ? (size_t)dict_value_doubles(p.dict_len) : 0) +
                      (p.coded == 2 ? (size_t)kSliceRows + (size_t)dict_pair_doubles(p.dict_len) : 0);
  const size_t fixed = kPersistHeaderBytes + 2 * xbuf * sizeof(double);
  const size_t per_sm = 228 * 1024, sys_reserve = 1024;  // 1 KB per resident CTA belongs to the driver
  const size_t entry_bytes = p.coded == 2 ? 1 : p.coded == 1 ? 3 : 10;
  auto stage_bytes = [&](int ku) { return (size_t)ku * kSliceRows * entry_bytes; };
  // Static shared memory of the instantiations this plan can launch (reduction scratch of the fused-dot variants):
  // it counts against the per-CTA limit together with the dynamic part, so it is asked of the runtime, not guessed.
  size_t static_smem[5] = {0, 0, 0, 0, 0};  // indexed by KU
  {
    cudaFuncAttributes fa;
#define CB_STATIC(KU_, DOT, CODED)                                                              \
    CB_CUDA(cudaFuncGetAttributes(&fa, spmv_ell_persistent_kernel<KU_, DOT, CODED>));            \
    static_smem[KU_] = std::max(static_smem[KU_], (size_t)fa.sharedSizeBytes)
    if (p.coded == 2) { CB_STATIC(2, false, 2); CB_STATIC(2, true, 2); CB_STATIC(4, false, 2); CB_STATIC(4, true, 2); }
    else if (p.coded == 1) { CB_STATIC(2, false, 1); CB_STATIC(2, true, 1); CB_STATIC(4, false, 1); CB_STATIC(4, true, 1); }
    else { CB_STATIC(2, false, 0); CB_STATIC(2, true, 0); CB_STATIC(4, false, 0); CB_STATIC(4, true, 0); }
#undef CB_STATIC
  }
  int ku = 0, stages = 0, ctas = 0;
  auto fits = [&](int k, int st, int c) {
    const size_t cta = fixed + st * stage_bytes(k) + static_smem[k];
    return c * (cta + sys_reserve) <= per_sm && cta <= (size_t)dev_smem;
  };
  if (p.coded) {
    // a stage is 3 KB (pair codes: 1 KB) per ELL column: deep rings of 4-column stages fit beside the x buffers; more
    // CTAs per SM keep more slices (x windows) in flight, which is what the shorter per-slice time needs
    const int k = ctx->persist_ku == 2 ? 2 : 4;
    const int c_hi = ctx->persist_ctas >= 1 && ctx->persist_ctas <= 3 ? ctx->persist_ctas : 2;
    for (int c = c_hi; c >= 1 && !ku; c--)
      for (int st = p.coded == 2 ? kMaxStages : 6; st >= 2 && !ku; st--)
        if (fits(k, st, c)) { ku = k; stages = st; ctas = c; }
  } else {
    if ((ctx->persist_ku == 0 || ctx->persist_ku == 2) && fits(2, 3, 2)) { ku = 2; stages = 3; ctas = 2; while (stages < 4 && fits(2, stages + 1, 2)) stages++; }
    if (!ku || ctx->persist_ku == 4) {
      ku = 0;
      for (int st = 4; st >= 2 && !ku; st--)
        if (fits(4, st, 1)) { ku = 4; stages = st; ctas = 1; }
    }
    if (!ku && fits(2, 2, 1)) { ku = 2; stages = 2; ctas = 1; while (stages < kMaxStages && fits(2, stages + 1, 1)) stages++; }
  }
  if (!ku) return CASK_B200_OK;  // x cache too large for the ring: the one-CTA-per-slice kernel is used
  p.persist_ku = ku;
  p.persist_stages = stages;
  p.persist_ctas_per_sm = ctas;
  p.persist_xbuf = (int32_t)xbuf;
  p.persist_smem = fixed + stages * stage_bytes(ku);
  return CASK_B200_OK;
}

// part: 0 = every slice, 1 = slices that read only this rank's own x (interior), 2 = the rest.
// With fusion, partials[0 .. spmv_num_ctas(part)) receives one partial dot per CTA.
int launch_spmv(cask_b200_ctx* ctx, const double* d_x, double* d_y, int part, cudaStream_t s,
                const SpmvFusion* fusion, const HaloWait* wait) {
  const Plan& p = ctx->plan;
  int ell_lo = 0, ell_hi = p.n_ell, csr_lo = 0, csr_hi = p.n_csr;
  if (part == 1) { ell_hi = p.n_ell_interior; csr_hi = p.n_csr_interior; }
  if (part == 2) { ell_lo = p.n_ell_interior; csr_lo = p.n_csr_interior; }
  return launch_spmv_range(ctx, d_x, d_y, ell_lo, ell_hi, csr_lo, csr_hi, s, fusion, wait);
}

// entries [ell_lo, ell_hi) of the staged-ELL list and [csr_lo, csr_hi) of the gather-CSR list
int launch_spmv_range(cask_b200_ctx* ctx, const double* d_x, double* d_y, int ell_lo, int ell_hi, int csr_lo, int csr_hi,
                      cudaStream_t s, const SpmvFusion* fusion, const HaloWait* wait) {
  const Plan& p = ctx->plan;
  const HaloWait hw = wait ? *wait : HaloWait();
  if (hw.ctrl && !(ctx->ell_kernel == 1 && p.persist_ku && csr_hi == csr_lo))
    return fail(CASK_B200_ERR_RUNTIME, "peer halo wait needs the persistent staged-ELL kernel");
  if ((reinterpret_cast<uintptr_t>(d_x) & 15u) != 0)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "x must be 16-byte aligned (TMA bulk copies)");
  const bool dot = fusion && fusion->d_dot_with;
  double* partials = dot ? fusion->d_partials : nullptr;
  const double* w = dot ? fusion->d_dot_with : nullptr;
  if (ell_hi > ell_lo && ctx->ell_kernel == 1 && p.persist_ku) {
    const int grid = ell_grid(ctx, ell_hi - ell_lo);
#define CB_PERSIST(KU, DOT, CODED)                                                                                   \
  do {                                                                                                               \
    CB_CUDA(cudaFuncSetAttribute(spmv_ell_persistent_kernel<KU, DOT, CODED>,                                         \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.persist_smem));                 \
    CB_CUDA(cudaLaunchKernelEx(&cfg, spmv_ell_persistent_kernel<KU, DOT, CODED>, (const SliceDesc*)p.d_slices,       \
                               (const int32_t*)(p.d_list_ell + ell_lo), ell_hi - ell_lo, (const Run*)p.d_runs,       \
                               (const double*)p.d_ell_vals, (const uint16_t*)p.d_ell_idx, d_x, d_y, w, partials,     \
                               (int)p.persist_xbuf, (int)p.persist_stages, hw, rd,                                   \
                               fusion ? fusion->keep_vectors : 0,                                                    \
                               fusion ? fusion->trace : (unsigned long long*)nullptr, ca, self_dot));                \
  } while (0)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPersistThreads);
    cfg.dynamicSmemBytes = p.persist_smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = fusion && fusion->pdl ? 1 : 0;
    const ReduceDesc rd = dot && fusion->reduce.partials && ell_lo == 0 && ell_hi == p.n_ell && csr_hi == csr_lo
                              ? fusion->reduce : ReduceDesc();
    const int self_dot = dot && fusion->fuse_self_dot && rd.partials && rd.nq == 2 ? 1 : 0;
    if (dot && fusion->fuse_self_dot && !self_dot)
      return fail(CASK_B200_ERR_RUNTIME, "fused y.y needs the in-kernel reduction of a single persistent launch");
    CodedArgs ca;
    if (p.coded) { ca.codes = p.d_ell_codes; ca.dict = p.d_ell_dict; ca.pairs = p.d_ell_pairs; ca.dict_len = p.dict_len; }
    if (p.coded == 2) {
      if (p.persist_ku == 2) { if (dot) CB_PERSIST(2, true, 2); else CB_PERSIST(2, false, 2); }
      else { if (dot) CB_PERSIST(4, true, 2); else CB_PERSIST(4, false, 2); }
    } else if (p.coded == 1) {
      if (p.persist_ku == 2) { if (dot) CB_PERSIST(2, true, 1); else CB_PERSIST(2, false, 1); }
      else { if (dot) CB_PERSIST(4, true, 1); else CB_PERSIST(4, false, 1); }
    } else {
      if (p.persist_ku == 2) { if (dot) CB_PERSIST(2, true, 0); else CB_PERSIST(2, false, 0); }
      else { if (dot) CB_PERSIST(4, true, 0); else CB_PERSIST(4, false, 0); }
    }
#undef CB_PERSIST
    ctx->launches++;
    if (partials) partials += grid;
  } else if (ell_hi > ell_lo) {
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
    if (partials) partials += ell_hi - ell_lo;
  }
  if (csr_hi > csr_lo && p.csr_merge) {
    const int i_lo = p.h_item_begin[csr_lo], i_hi = p.h_item_begin[csr_hi];
    if (i_lo < 0 || i_hi < 0) return fail(CASK_B200_ERR_RUNTIME, "gather range splits a merge-path run");
    const int tiles = i_hi - i_lo;
    if (tiles > 0) {
      const int tile_items = kMergeThreads * p.merge_items;
      const size_t smem = sizeof(double) * tile_items + sizeof(int32_t) * (tile_items + 4);
      const int fix = (tiles + 255) / 256;
      // x gathers cached in L2 only: an L1 line per scattered 8-byte gather buys nothing (4 % sector hit rate) and costs
      // the miss tracking the next gathers need; measured 3.25 vs 3.33 ms on R-MAT scale 25 (profiles/r2c_rmat_sweep.md)
      // with hub-clustered columns (plan.cu: build_col_reorder) the hubs share lines and L1 is what serves them
      static const int xcg_env = getenv("CASK_B200_MERGE_XCG") ? atoi(getenv("CASK_B200_MERGE_XCG")) : -1;
      static const int diag = getenv("CASK_B200_MERGE_DIAG") ? atoi(getenv("CASK_B200_MERGE_DIAG")) & 6 : 0;
      const int32_t* cols = p.d_col;
      const double* xg = d_x;
      if (p.d_col_perm) {
        if (!p.xperm_external) {  // sharded sparse exchange: dist_exchange_begin / wait has filled d_xperm
          permute_x_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(p.d_perm, d_x, p.d_xperm, p.cols_used);
          ctx->launches++;
        }
        cols = p.d_col_perm;
        xg = p.d_xperm;
      }
      const int flags = ((xcg_env >= 0 ? xcg_env : (p.hub_ordered ? 0 : 1)) ? kMergeXPastL1 : 0) | diag;
#define CB_MERGE(DOT, ITEMS, CTAS)                                                                                        \
  do {                                                                                                                    \
    CB_CUDA(cudaFuncSetAttribute(spmv_csr_merge_kernel<DOT, ITEMS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    spmv_csr_merge_kernel<DOT, ITEMS, CTAS><<<tiles, kMergeThreads, smem, s>>>(p.d_merge_tiles + i_lo, (int32_t)p.n, p.d_row_ptr,   \
                                                                               cols, p.d_val, xg, d_y, w, partials,                \
                                                                               p.d_merge_carry + i_lo, flags);                     \
  } while (0)
      // resident CTAs per SM the kernel is compiled for (register budget): 5 x 46 registers by default; profiling sessions
      // try more warps in flight (CASK_B200_MERGE_CTAS = 6: 40 registers, 7: 36, 8 with 5 items: 32)
      static const int env_ctas = getenv("CASK_B200_MERGE_CTAS") ? atoi(getenv("CASK_B200_MERGE_CTAS")) : -1;
      const int want_ctas = env_ctas >= 0 ? env_ctas : p.merge_ctas;
#define CB_MERGE_ITEMS(DOT)                                                      \
  switch (p.merge_items) {                                                       \
    case 5:                                                                      \
      if (want_ctas == 8) CB_MERGE(DOT, 5, 8);                                   \
      else if (want_ctas == 7) CB_MERGE(DOT, 5, 7);                              \
      else if (want_ctas == 6) CB_MERGE(DOT, 5, 6);                              \
      else CB_MERGE(DOT, 5, 5);                                                  \
      break;                                                                     \
    case 7:                                                                      \
      if (want_ctas == 7) CB_MERGE(DOT, 7, 7);                                   \
      else if (want_ctas == 6) CB_MERGE(DOT, 7, 6);                              \
      else CB_MERGE(DOT, 7, 5);                                                  \
      break;                                                                     \
    case 11: CB_MERGE(DOT, 11, 5); break;                                        \
    case 17: CB_MERGE(DOT, 17, 4); break;                                        \
    default: return fail(CASK_B200_ERR_RUNTIME, "merge_items must be 5, 7, 11 or 17"); \
  }
      if (dot) {
        CB_MERGE_ITEMS(true);
        spmv_csr_merge_fixup_kernel<true><<<fix, 256, 0, s>>>(p.d_merge_tiles + i_lo, tiles, p.d_row_ptr, p.d_merge_carry + i_lo, d_y, w,
                                                              partials + tiles);
      } else {
        CB_MERGE_ITEMS(false);
        spmv_csr_merge_fixup_kernel<false><<<fix, 256, 0, s>>>(p.d_merge_tiles + i_lo, tiles, p.d_row_ptr, p.d_merge_carry + i_lo, d_y,
                                                               nullptr, nullptr);
      }
#undef CB_MERGE_ITEMS
#undef CB_MERGE
      ctx->launches += 2;
    }
  } else if (csr_hi > csr_lo) {
    const int i_lo = p.h_item_begin[csr_lo], i_hi = p.h_item_begin[csr_hi];
    const int s_lo = p.h_split_begin[csr_lo], s_hi = p.h_split_begin[csr_hi];
    if (i_hi > i_lo) {
      if (p.csr_stream) {
        // products of one item in shared memory: item target + one row short of the long-row threshold
        const size_t smem = sizeof(double) * (size_t)(p.csr_item_nnz + kCsrLongRow + 2) + sizeof(int32_t) * (kSliceRows + 2);
        if (dot) {
          CB_CUDA(cudaFuncSetAttribute(spmv_csr_items_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          CB_CUDA(cudaFuncSetAttribute(spmv_csr_items_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          spmv_csr_items_kernel<true, true><<<i_hi - i_lo, 256, smem, s>>>(p.d_csr_items + i_lo, p.d_row_ptr, p.d_col, p.d_val,
                                                                             d_x, d_y, w, partials, p.d_csr_scratch);
        } else {
          CB_CUDA(cudaFuncSetAttribute(spmv_csr_items_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          CB_CUDA(cudaFuncSetAttribute(spmv_csr_items_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          spmv_csr_items_kernel<false, true><<<i_hi - i_lo, 256, smem, s>>>(p.d_csr_items + i_lo, p.d_row_ptr, p.d_col, p.d_val,
                                                                              d_x, d_y, nullptr, nullptr, p.d_csr_scratch);
        }
      } else if (dot) {
        spmv_csr_items_kernel<true, false><<<i_hi - i_lo, 256, 0, s>>>(p.d_csr_items + i_lo, p.d_row_ptr, p.d_col, p.d_val,
                                                                        d_x, d_y, w, partials, p.d_csr_scratch);
      } else {
        spmv_csr_items_kernel<false, false><<<i_hi - i_lo, 256, 0, s>>>(p.d_csr_items + i_lo, p.d_row_ptr, p.d_col, p.d_val,
                                                                         d_x, d_y, nullptr, nullptr, p.d_csr_scratch);
      }
      ctx->launches++;
    }
    if (s_hi > s_lo) {
      if (dot) spmv_csr_fixup_kernel<true><<<1, 256, 0, s>>>(p.d_split_rows + s_lo, s_hi - s_lo, p.d_csr_scratch, d_y, w,
                                                             partials + (i_hi - i_lo));
      else spmv_csr_fixup_kernel<false><<<1, 256, 0, s>>>(p.d_split_rows + s_lo, s_hi - s_lo, p.d_csr_scratch, d_y, nullptr, nullptr);
      ctx->launches++;
    }
  }
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

}  // namespace caskb200
