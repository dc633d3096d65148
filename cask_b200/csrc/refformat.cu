// GPU partitioner that emits the REFERENCE partition format, byte for byte:
//   cask::spmv::Spmv::preprocess        src/runtime/Spmv.cpp:329-365   (row stripes)
//   cask::CsrMatrix::sliceColumns       src/runtime/SparseMatrix.hpp:459-482 (column blocks,
//                                       cumulative END offsets, block-local column index)
//   SkipEmptyRowsSpmv::encodeEmptyRows  src/runtime/Spmv.hpp:213-238   (run-length empty rows)
//   Spmv::do_blocking                   src/runtime/Spmv.cpp:42-107    (concatenation, padding to
//                                       input_width, cycle model)
// and the kernel that computes y from those arrays the way the dataflow engine consumes them
// (src/spmv/src/ParallelCsrReadControl.java:160-196, SpmvKernel.java:61-78,286-297).
//
// The format is O(rows x blocks) by construction (one END offset per row per column block), so
// this path exists for parity and for the legacy run/write/read plugin, not for BASELINE sizes;
// the throughput path is plan.cu / spmv.cu.
#include <algorithm>
#include <climits>
#include <cstring>

#include "ctx.cuh"
#include "scan.cuh"

namespace caskb200 {

namespace {

constexpr uint32_t kEmptyFlag = 0x80000000u;

// counts[b*ns + i] = nonzeros of stripe row i that fall in column block b.  One thread per row,
// so no atomics are needed and unsorted rows are handled too.
__global__ void count_block_entries_kernel(const int32_t* __restrict__ row_ptr,
                                           const int32_t* __restrict__ col, int32_t row0, int32_t ns,
                                           int32_t bs, int32_t* __restrict__ counts) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  for (int32_t k = row_ptr[row0 + i]; k < row_ptr[row0 + i + 1]; k++)
    counts[(int64_t)(col[k] / bs) * ns + i] += 1;
}

struct BlockSummary {  // per column block
  int32_t nnz;         // entries in the block
  int32_t cycles;      // countComputeCycles of the un-encoded END offsets (Spmv.cpp:25-40)
  int32_t enc_len;     // length of the (possibly run-length encoded) colptr of this block
  int32_t pad_;
};

// One CTA per column block: END offsets (inclusive scan of the counts), the cycle model in closed
// form and the encoded length.  Row i with length L starting at lane cursor c0 = END[i-1] mod w
// costs max(1, ceil((c0 + L) / w)) cycles: the do/while of Spmv.cpp:31-37 restated.
__global__ void __launch_bounds__(kScanThreads)
scan_block_rows_kernel(const int32_t* __restrict__ counts, int32_t* __restrict__ endoff, int32_t ns,
                       int32_t w, int32_t nblocks, int arch, BlockSummary* __restrict__ summary) {
  __shared__ int32_t smem[kScanThreads / 32];
  __shared__ int32_t red[3][kScanThreads / 32];
  const int32_t b = blockIdx.x;
  const int32_t* c = counts + (int64_t)b * ns;
  int32_t* e = endoff + (int64_t)b * ns;
  int32_t carry = 0, cycles = 0, nonempty = 0, runs = 0;
  for (int32_t base = 0; base < ns; base += kScanThreads) {
    const int32_t i = base + threadIdx.x;
    const int32_t len = i < ns ? c[i] : 0;
    int32_t excl, total;
    int32_t incl = cta_scan(len, OpAddI32(), 0, smem, &excl, &total);
    if (i < ns) {
      e[i] = carry + incl;
      const int32_t c0 = (carry + excl) % w;
      const int32_t cyc = (c0 + len + w - 1) / w;
      cycles += cyc > 1 ? cyc : 1;
      if (len) nonempty++;
      else if (i == 0 || c[i - 1] != 0) runs++;  // an empty run starts here
    }
    carry += total;
  }
  // CTA reductions of the three per-thread counters
  int32_t vals[3] = {cycles, nonempty, runs};
#pragma unroll
  for (int q = 0; q < 3; q++) {
    int32_t v = vals[q];
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t[3] = {0, 0, 0};
    for (int q = 0; q < 3; q++)
      for (int wv = 0; wv < kScanThreads / 32; wv++) t[q] += red[q][wv];
    const bool encode = arch == CASK_B200_ARCH_SKIPEMPTY && b != 0 && b != nblocks - 1;  // Spmv.hpp:244
    summary[b].nnz = carry;
    summary[b].cycles = t[0];
    summary[b].enc_len = encode ? t[1] + t[2] : ns;
    summary[b].pad_ = 0;
  }
}

// One CTA per column block: writes the block's colptr segment.  Encoded blocks: a non-empty row i
// goes to (#non-empty rows before i) + (#empty runs before i); a run is written from its LAST row e
// at (#non-empty rows before e) + (#runs up to e) - 1 with length e - (last non-empty row before e).
__global__ void __launch_bounds__(kScanThreads)
emit_colptr_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ endoff, int32_t ns,
                   int32_t nblocks, int arch, const int64_t* __restrict__ colptr_base,
                   int32_t* __restrict__ colptr) {
  __shared__ int64_t smem[kScanThreads / 32];
  const int32_t b = blockIdx.x;
  const int32_t* c = counts + (int64_t)b * ns;
  const int32_t* e = endoff + (int64_t)b * ns;
  int32_t* out = colptr + colptr_base[b];
  const bool encode = arch == CASK_B200_ARCH_SKIPEMPTY && b != 0 && b != nblocks - 1;
  if (!encode) {
    for (int32_t i = threadIdx.x; i < ns; i += kScanThreads) out[i] = e[i];
    return;
  }
  int64_t carry_cnt = 0;   // (non-empty rows << 32) | run starts, so one scan carries both
  int64_t carry_last = -1; // index of the last non-empty row seen
  for (int32_t base = 0; base < ns; base += kScanThreads) {
    const int32_t i = base + threadIdx.x;
    const bool in = i < ns;
    const bool empty = in && c[i] == 0;
    const bool nonempty = in && !empty;
    const bool run_start = empty && (i == 0 || c[i - 1] != 0);
    const bool run_end = empty && (i == ns - 1 || c[i + 1] != 0);
    int64_t excl, total;
    int64_t packed = ((int64_t)(nonempty ? 1 : 0) << 32) | (run_start ? 1 : 0);
    int64_t incl = carry_cnt + cta_scan(packed, OpAddI64(), (int64_t)0, smem, &excl, &total);
    int64_t lexcl, ltotal;
    int64_t last = cta_scan((int64_t)(nonempty ? i : -1), OpMaxI64(), (int64_t)-1, smem, &lexcl, &ltotal);
    last = last > carry_last ? last : carry_last;
    const int32_t ne_incl = (int32_t)(incl >> 32), rs_incl = (int32_t)(incl & 0xffffffff);
    if (nonempty) out[ne_incl - 1 + rs_incl] = e[i];
    if (run_end) out[ne_incl + rs_incl - 1] = (int32_t)((uint32_t)(i - (int32_t)last) | kEmptyFlag);
    carry_cnt += total;
    carry_last = ltotal > carry_last ? ltotal : carry_last;
  }
}

// One thread per stripe row: scatter (value, block-local index) records to their place in the pair
// stream.  counts[] doubles as the per-(block,row) countdown, which keeps the original order of the
// row's entries inside a block (SparseMatrix.hpp:473-479) for sorted and unsorted rows alike.
__global__ void fill_pairs_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                  const double* __restrict__ val, int32_t row0, int32_t ns, int32_t bs,
                                  int32_t* __restrict__ counts, const int32_t* __restrict__ endoff,
                                  const int64_t* __restrict__ pair_base, uint32_t* __restrict__ pairs) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  for (int32_t k = row_ptr[row0 + i]; k < row_ptr[row0 + i + 1]; k++) {
    const int32_t cidx = col[k];
    const int32_t b = cidx / bs;
    const int64_t cell = (int64_t)b * ns + i;
    const int32_t remaining = counts[cell];
    counts[cell] = remaining - 1;
    const int64_t pos = pair_base[b] + endoff[cell] - remaining;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(val[k]);
    uint32_t* rec = pairs + pos * 3;  // 12-byte packed record: value (8) then index (4), Spmv.hpp:13-20
    rec[0] = (uint32_t)(bits & 0xffffffffu);
    rec[1] = (uint32_t)(bits >> 32);
    rec[2] = (uint32_t)(cidx - b * bs);
  }
}

__global__ void zero_pair_values_kernel(uint32_t* pairs, int64_t npairs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npairs) { pairs[i * 3] = 0; pairs[i * 3 + 1] = 0; }
}

int32_t round_up_i32(int32_t v, int32_t to) { return v % to == 0 ? v : (v / to + 1) * to; }

int build_one(cask_b200_ctx* ctx, int32_t row0, int32_t ns, RefPartition* out) {
  const Plan& pl = ctx->plan;
  const cask_b200_design& d = ctx->design;
  const int32_t bs = d.cache_size, w = d.input_width;
  const int32_t m = (int32_t)pl.m;
  const int32_t nblocks = m / bs + (m % bs == 0 ? 0 : 1);
  const int64_t cells = (int64_t)nblocks * ns;
  if (cells > (int64_t)INT32_MAX)
    return fail(CASK_B200_ERR_UNSUPPORTED,
                "reference partition format needs rows x blocks = " + std::to_string(cells) +
                    " row pointers (> INT32_MAX; the reference's own int counters overflow here, "
                    "Spmv.cpp:68) - use a larger cache_size for the export");
  cudaStream_t s = ctx->stream;
  int32_t *d_counts = nullptr, *d_endoff = nullptr;
  BlockSummary* d_sum = nullptr;
  int64_t *d_bases = nullptr;
  const int64_t alloc_cells = std::max<int64_t>(cells, 1);
  CB_CUDA(cudaMalloc(&d_counts, sizeof(int32_t) * alloc_cells));
  CB_CUDA(cudaMalloc(&d_endoff, sizeof(int32_t) * alloc_cells));
  CB_CUDA(cudaMalloc(&d_sum, sizeof(BlockSummary) * std::max(nblocks, 1)));
  CB_CUDA(cudaMalloc(&d_bases, sizeof(int64_t) * 2 * std::max(nblocks, 1)));
  CB_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * alloc_cells, s));
  std::vector<BlockSummary> sum(nblocks);
  if (ns > 0) {
    count_block_entries_kernel<<<(ns + 255) / 256, 256, 0, s>>>(pl.d_row_ptr, pl.d_col, row0, ns, bs, d_counts);
    ctx->launches++;
    scan_block_rows_kernel<<<nblocks, kScanThreads, 0, s>>>(d_counts, d_endoff, ns, w, nblocks, d.arch, d_sum);
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(sum.data(), d_sum, sizeof(BlockSummary) * nblocks, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
  } else {
    for (auto& b : sum) b = BlockSummary{0, 0, 0, 0};
  }
  // O(nblocks) bookkeeping: bases of every block in the two streams and the Partition scalars.
  std::vector<int64_t> bases(2 * (size_t)std::max(nblocks, 1));
  int64_t ncolptr = 0, npairs = 0;
  int32_t cycles = 0, reduction = ns * nblocks, empty = 0;  // Spmv.cpp:66-69
  for (int32_t b = 0; b < nblocks; b++) {
    bases[b] = ncolptr;
    bases[nblocks + b] = npairs;
    const int32_t diff = ns - sum[b].enc_len;  // Spmv.cpp:76
    empty += diff;
    reduction -= diff;
    cycles += sum[b].cycles - diff;            // Spmv.cpp:79
    ncolptr += sum[b].enc_len;
    npairs += round_up_i32(sum[b].nnz, w);     // Spmv.cpp:81-82 via Utils.hpp:61-68
  }
  const int32_t out_len = ns ? round_up_i32(ns, 384 / 8) : 0;  // Spmv.cpp:91-92 (384-byte bursts)
  const int32_t v_len = round_up_i32(m, bs);                    // Spmv.cpp:93
  cask_b200_partition_info& info = out->info;
  info.nBlocks = nblocks;
  info.n = ns;
  info.paddingCycles = out_len - ns;
  info.totalCycles = cycles + v_len;
  info.vector_load_cycles = v_len / nblocks;
  info.outSize = out_len * 8;
  info.reductionCycles = reduction;
  info.emptyCycles = empty;
  info.m_colptr_unpaddedLength = (int32_t)ncolptr;
  info.m_indptr_values_unpaddedLength = (int32_t)npairs;
  info.len_colptr = ncolptr;
  info.len_pairs = npairs;
  out->row0 = row0;
  CB_CUDA(cudaMalloc(&out->d_colptr, sizeof(int32_t) * std::max<int64_t>(ncolptr, 1)));
  CB_CUDA(cudaMalloc(&out->d_pairs, 12 * std::max<int64_t>(npairs, 1)));
  CB_CUDA(cudaMemsetAsync(out->d_pairs, 0, 12 * std::max<int64_t>(npairs, 1), s));  // padding = (0.0, 0)
  if (ns > 0) {
    CB_CUDA(cudaMemcpyAsync(d_bases, bases.data(), sizeof(int64_t) * 2 * nblocks, cudaMemcpyHostToDevice, s));
    emit_colptr_kernel<<<nblocks, kScanThreads, 0, s>>>(d_counts, d_endoff, ns, nblocks, d.arch, d_bases,
                                                        out->d_colptr);
    ctx->launches++;
    fill_pairs_kernel<<<(ns + 255) / 256, 256, 0, s>>>(pl.d_row_ptr, pl.d_col, pl.d_val, row0, ns, bs, d_counts,
                                                       d_endoff, d_bases + nblocks, (uint32_t*)out->d_pairs);
    ctx->launches++;
  }
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaGetLastError());
  cudaFree(d_counts); cudaFree(d_endoff); cudaFree(d_sum); cudaFree(d_bases);
  return CASK_B200_OK;
}

}  // namespace

void free_ref_partitions(cask_b200_ctx* ctx) {
  for (auto& p : ctx->ref_parts) {
    cudaFree(p.d_colptr);
    cudaFree(p.d_pairs);
  }
  ctx->ref_parts.clear();
  ctx->ref_built = false;
}

// Row striping of Spmv::preprocess (Spmv.cpp:334-364), including the rows < pipes corner where every
// pipe receives the whole matrix and pipes 1.. get their values zeroed (Spmv.cpp:337-351).
int build_ref_partitions(cask_b200_ctx* ctx) {
  if (ctx->ref_built) return CASK_B200_OK;
  free_ref_partitions(ctx);
  const Plan& pl = ctx->plan;
  const cask_b200_design& d = ctx->design;
  if (pl.n_global != pl.n)
    return fail(CASK_B200_ERR_UNSUPPORTED, "reference-format export is defined for an unsharded matrix");
  if (pl.m <= 0) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "matrix has no columns");
  const int32_t n = (int32_t)pl.n, P = d.num_pipes;
  ctx->ref_parts.resize(P);
  const int32_t rpp = n / P;
  if (rpp == 0) {
    CB_TRY(build_one(ctx, 0, n, &ctx->ref_parts[0]));
    const RefPartition& p0 = ctx->ref_parts[0];
    for (int32_t p = 1; p < P; p++) {
      RefPartition& q = ctx->ref_parts[p];
      q.info = p0.info;
      q.row0 = 0;
      q.values_zeroed = true;
      const int64_t nc = std::max<int64_t>(p0.info.len_colptr, 1), np = std::max<int64_t>(p0.info.len_pairs, 1);
      CB_CUDA(cudaMalloc(&q.d_colptr, sizeof(int32_t) * nc));
      CB_CUDA(cudaMalloc(&q.d_pairs, 12 * np));
      CB_CUDA(cudaMemcpyAsync(q.d_colptr, p0.d_colptr, sizeof(int32_t) * nc, cudaMemcpyDeviceToDevice, ctx->stream));
      CB_CUDA(cudaMemcpyAsync(q.d_pairs, p0.d_pairs, 12 * np, cudaMemcpyDeviceToDevice, ctx->stream));
      if (p0.info.len_pairs > 0) {
        zero_pair_values_kernel<<<(unsigned)((p0.info.len_pairs + 255) / 256), 256, 0, ctx->stream>>>(
            (uint32_t*)q.d_pairs, p0.info.len_pairs);
        ctx->launches++;
      }
    }
    CB_CUDA(cudaStreamSynchronize(ctx->stream));
  } else {
    int32_t start = 0;
    for (int32_t p = 0; p < P; p++) {
      const int32_t rows = p == P - 1 ? n - start : rpp;
      CB_TRY(build_one(ctx, start, rows, &ctx->ref_parts[p]));
      start += rows;
    }
  }
  ctx->ref_built = true;
  return CASK_B200_OK;
}

// ---- y from the reference-format arrays -----------------------------------------------------
namespace {

// weight of a colptr entry in rows: a run-length marker stands for k rows, anything else for one
__global__ void entry_rows_kernel(const int32_t* __restrict__ colptr, int64_t len, int64_t* __restrict__ rows) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  const uint32_t e = (uint32_t)colptr[j];
  rows[j] = (e & kEmptyFlag) ? (int64_t)(e & 0x7fffffffu) : 1;
}

// key = (block << 32) | END offset for real entries, (block << 32) for markers: an inclusive max-scan
// of the keys gives every entry the END offset of the previous real entry of ITS block.
__global__ void entry_keys_kernel(const int32_t* __restrict__ colptr, const int64_t* __restrict__ rows_incl,
                                  int64_t len, int32_t ns, int64_t* __restrict__ keys) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  const uint32_t e = (uint32_t)colptr[j];
  const bool marker = e & kEmptyFlag;
  const int64_t w = marker ? (int64_t)(e & 0x7fffffffu) : 1;
  const int64_t before = rows_incl[j] - w;
  keys[j] = ((before / ns) << 32) | (marker ? 0 : (int64_t)e);
}

// nnz of block b = largest END offset inside it = low word of the scanned key at the block's last entry
__global__ void block_nnz_kernel(const int64_t* __restrict__ rows_incl, const int64_t* __restrict__ keys_incl,
                                 int64_t len, int32_t ns, int32_t w, int64_t* __restrict__ block_pairs) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  if (rows_incl[j] % ns == 0 && (j + 1 == len || rows_incl[j + 1] != rows_incl[j])) {
    const int64_t b = rows_incl[j] / ns - 1;
    const int64_t nnz = keys_incl[j] & 0xffffffff;
    block_pairs[b] = (nnz + w - 1) / w * w;
  }
}

__global__ void refformat_rows_kernel(const int32_t* __restrict__ colptr, const int64_t* __restrict__ rows_incl,
                                      const int64_t* __restrict__ keys_incl, const int64_t* __restrict__ pair_base_incl,
                                      const uint32_t* __restrict__ pairs, int64_t len, int32_t ns, int32_t bs,
                                      int32_t m, const double* __restrict__ x, double* __restrict__ y) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  const uint32_t e = (uint32_t)colptr[j];
  if (e & kEmptyFlag) return;
  const int64_t before = rows_incl[j] - 1;
  const int64_t b = before / ns;
  const int32_t row = (int32_t)(before % ns);
  int64_t start = 0;
  if (j > 0 && (keys_incl[j - 1] >> 32) == b) start = keys_incl[j - 1] & 0xffffffff;
  const int64_t base = (b ? pair_base_incl[b - 1] : 0) + start;
  const int32_t cnt = (int32_t)((int64_t)e - start);
  double acc = 0.0;
  for (int32_t k = 0; k < cnt; k++) {
    const uint32_t* rec = pairs + (base + k) * 3;
    const double v = __longlong_as_double((long long)(((unsigned long long)rec[1] << 32) | rec[0]));
    const int64_t colg = b * bs + (int32_t)rec[2];
    const double xv = colg < m ? x[colg] : 0.0;  // Spmv.cpp:211-213: x is zero-padded to the cache size
    acc = __dadd_rn(acc, __dmul_rn(v, xv));
  }
  if (cnt) atomicAdd(&y[row], acc);  // blocks of one row accumulate (SpmvKernel.java:286-297)
}

}  // namespace

// One stripe: y[0..ns) += A_stripe x, from raw device arrays in the reference layout.  `m` bounds the
// readable part of x (columns beyond it read as the zero padding of Spmv.cpp:211-213).
int refformat_stripe(cudaStream_t s, int64_t* launches, const int32_t* d_colptr, int64_t len, const uint8_t* d_pairs,
                     int32_t ns, int32_t nb, int32_t cache, int32_t w, int64_t m, const double* d_x, double* d_y) {
  if (len == 0 || ns == 0) return CASK_B200_OK;
  int64_t *rows = nullptr, *keys = nullptr, *bp = nullptr;
  CB_CUDA(cudaMalloc(&rows, sizeof(int64_t) * len));
  CB_CUDA(cudaMalloc(&keys, sizeof(int64_t) * len));
  CB_CUDA(cudaMalloc(&bp, sizeof(int64_t) * std::max(nb, 1)));
  const unsigned grid = (unsigned)((len + 255) / 256);
  entry_rows_kernel<<<grid, 256, 0, s>>>(d_colptr, len, rows);
  CB_CUDA((device_inclusive_scan<int64_t, OpAddI64>(rows, rows, len, OpAddI64(), 0, s, launches)));
  entry_keys_kernel<<<grid, 256, 0, s>>>(d_colptr, rows, len, ns, keys);
  CB_CUDA((device_inclusive_scan<int64_t, OpMaxI64>(keys, keys, len, OpMaxI64(), (int64_t)-1, s, launches)));
  CB_CUDA(cudaMemsetAsync(bp, 0, sizeof(int64_t) * std::max(nb, 1), s));
  block_nnz_kernel<<<grid, 256, 0, s>>>(rows, keys, len, ns, w, bp);
  CB_CUDA((device_inclusive_scan<int64_t, OpAddI64>(bp, bp, nb, OpAddI64(), 0, s, launches)));
  refformat_rows_kernel<<<grid, 256, 0, s>>>(d_colptr, rows, keys, bp, (const uint32_t*)d_pairs, len, ns, cache,
                                             (int32_t)m, d_x, d_y);
  if (launches) *launches += 4;
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaGetLastError());
  cudaFree(rows); cudaFree(keys); cudaFree(bp);
  return CASK_B200_OK;
}

int spmv_refformat_device(cask_b200_ctx* ctx, const double* d_x, double* d_y) {
  CB_TRY(build_ref_partitions(ctx));
  const Plan& pl = ctx->plan;
  const cask_b200_design& d = ctx->design;
  cudaStream_t s = ctx->stream;
  const int32_t n = (int32_t)pl.n;
  CB_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double) * std::max<int64_t>(n, 1), s));
  // Spmv.cpp:303-326: results of the pipes are concatenated, then cut to n rows; with fewer rows than
  // pipes only pipe 0 (the one with real values) survives the cut.
  const int32_t live = (n / d.num_pipes == 0) ? 1 : d.num_pipes;
  for (int32_t p = 0; p < live; p++) {
    const RefPartition& q = ctx->ref_parts[p];
    CB_TRY(refformat_stripe(s, &ctx->launches, q.d_colptr, q.info.len_colptr, q.d_pairs, q.info.n, q.info.nBlocks,
                            d.cache_size, d.input_width, pl.m, d_x, d_y + q.row0));
  }
  return CASK_B200_OK;
}

}  // namespace caskb200
