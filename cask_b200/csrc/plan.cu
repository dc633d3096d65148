// GPU-side partitioner for the THROUGHPUT format.
//
// Same partitioning decisions as the reference — row stripes of Spmv::preprocess
// (src/runtime/Spmv.cpp:334-364) and an on-chip x cache of `cache_size` doubles per unit of work
// (the column blocks of CsrMatrix::sliceColumns, src/runtime/SparseMatrix.hpp:459-482) — but stored
// for a B200 instead of a dataflow pipe:
//
//   * every stripe is cut into slices of kSliceRows rows;
//   * for each slice the set of 128-byte x granules its rows touch is found with a shared-memory
//     bitmap; consecutive granules merge into "runs", each staged by ONE TMA bulk copy; if the
//     staged doubles fit the cache budget the slice becomes a STAGED-ELL slice: values stored
//     column-major (thread-per-row, 128-bit coalesced loads) and column indices rewritten to 16-bit
//     positions inside the slice's shared-memory x cache;
//   * slices whose rows are too irregular (ELL fill below ell_min_fill) or whose columns are too
//     scattered for the cache keep plain CSR and are executed vector-per-row with x gathered
//     through L2; the lanes-per-row width is chosen from the row-length histogram.
#include <algorithm>
#include <cstring>

#include "ctx.cuh"
#include "devlogic.cuh"

#include "valuedict_logic.inl"

namespace caskb200 {

namespace {

struct SliceCount {     // output of the analysis pass, one per slice
  int32_t width;        // max row length
  int32_t nnz;
  int32_t nruns;        // -1: span too wide for the bitmap
  int32_t xlen;         // staged doubles incl. zero slots (valid when nruns >= 0)
  int32_t col_hi;       // one past the largest referenced column, rounded up to a granule
  int32_t pad_[3];
};

__device__ __forceinline__ int32_t warp_max(int32_t v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ int32_t warp_min(int32_t v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ int32_t warp_sum(int32_t v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Shared state of one slice while its x-cache layout is derived.
struct SliceSmem {
  uint32_t bitmap[kBitmapWords];   // bit g set: granule gmin+g is referenced
  uint16_t prefix[kBitmapWords];   // number of set bits in words before this one (<= 1536 granules)
  int32_t red[3][8];
  int32_t gmin, gmax, width, nruns, ngran, ok;
};

// Finds gmin/gmax/width, fills bitmap + prefix.  Returns false (for the whole CTA) if the granule
// span does not fit the bitmap or more granules are referenced than max_gran.
__device__ bool analyse_slice(SliceSmem& sm, const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                              int32_t row0, int32_t nrows, int32_t max_gran) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t k0 = row_ptr[row0], k1 = row_ptr[row0 + nrows];
  int32_t lmax = 0, cmin = INT32_MAX, cmax = -1;
  for (int32_t r = tid; r < nrows; r += blockDim.x)
    lmax = max(lmax, row_ptr[row0 + r + 1] - row_ptr[row0 + r]);
  for (int32_t k = k0 + tid; k < k1; k += blockDim.x) {
    const int32_t c = col[k];
    cmin = min(cmin, c);
    cmax = max(cmax, c);
  }
  lmax = warp_max(lmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
  if (lane == 0) { sm.red[0][warp] = lmax; sm.red[1][warp] = cmin; sm.red[2][warp] = cmax; }
  __syncthreads();
  if (tid == 0) {
    int32_t a = 0, b = INT32_MAX, c = -1;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a = max(a, sm.red[0][w]); b = min(b, sm.red[1][w]); c = max(c, sm.red[2][w]); }
    sm.width = a;
    sm.gmin = b == INT32_MAX ? 0 : (b >> kGranuleShift);
    sm.gmax = c < 0 ? -1 : (c >> kGranuleShift);
    sm.ok = (sm.gmax - sm.gmin + 1) <= kBitmapWords * 32;
  }
  __syncthreads();
  if (!sm.ok) return false;
  const int32_t gmin = sm.gmin;
  const int32_t words = sm.gmax < gmin ? 0 : ((sm.gmax - gmin) >> 5) + 1;
  for (int32_t w = tid; w < words; w += blockDim.x) sm.bitmap[w] = 0;
  __syncthreads();
  for (int32_t k = k0 + tid; k < k1; k += blockDim.x) {
    const int32_t g = (col[k] >> kGranuleShift) - gmin;
    atomicOr(&sm.bitmap[g >> 5], 1u << (g & 31));
  }
  __syncthreads();
  // prefix popcounts + run count; words <= 8192, done by warp 0 in chunks of 32 words
  if (warp == 0) {
    int32_t carry = 0, runs = 0;
    uint32_t prev_msb = 0;
    for (int32_t base = 0; base < words; base += 32) {
      const int32_t w = base + lane;
      const uint32_t bits = w < words ? sm.bitmap[w] : 0u;
      int32_t pc = __popc(bits), incl = pc;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      uint32_t below = __shfl_up_sync(0xffffffffu, bits, 1);
      const uint32_t carry_in = lane ? (below >> 31) : prev_msb;
      runs += __popc(bits & ~((bits << 1) | carry_in));  // bits that start a run
      const int32_t excl = carry + incl - pc;
      if (w < words) sm.prefix[w] = (uint16_t)min(excl, 65535);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      prev_msb = __shfl_sync(0xffffffffu, bits, 31) >> 31;
    }
    runs = warp_sum(runs);
    if (lane == 0) { sm.ngran = carry; sm.nruns = runs; sm.ok = carry <= max_gran && runs <= kMaxRuns; }
  }
  __syncthreads();
  return sm.ok;
}

__global__ void __launch_bounds__(256)
plan_count_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const int32_t* __restrict__ slice_row0,
                  const int32_t* __restrict__ slice_nrows, int32_t max_gran, SliceCount* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char raw[];
  SliceSmem& sm = *reinterpret_cast<SliceSmem*>(raw);
  const int32_t s = blockIdx.x;
  const int32_t row0 = slice_row0[s], nrows = slice_nrows[s];
  const bool ok = analyse_slice(sm, row_ptr, col, row0, nrows, max_gran);
  if (threadIdx.x == 0) {
    SliceCount c;
    c.width = sm.width;
    c.nnz = row_ptr[row0 + nrows] - row_ptr[row0];
    c.nruns = ok ? sm.nruns : -1;
    c.xlen = ok ? kZeroSlots + sm.ngran * kGranule : 0;
    c.col_hi = (sm.gmax + 1) << kGranuleShift;
    c.pad_[0] = c.pad_[1] = c.pad_[2] = 0;
    out[s] = c;
  }
}

// Second pass for staged-ELL slices: rebuild the bitmap, write the runs, rewrite the slice into the
// ELL arrays.  Layout inside a slice (T = kEllThreads, RPT = kEllRowsPerThread):
//   entry (k, row) with row = j*T + t  lives at  val_off + (k*T + t)*RPT + j
// so thread t reads RPT consecutive values with 128-bit loads and, for a fixed j, the lanes of a
// warp own consecutive rows (coalesced y, conflict-free shared-memory x reads on banded matrices).
__global__ void __launch_bounds__(256)
plan_fill_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                 SliceDesc* __restrict__ slices, const int32_t* __restrict__ list, int32_t max_gran, int32_t m,
                 Run* __restrict__ runs, double* __restrict__ ell_vals, uint16_t* __restrict__ ell_idx) {
  extern __shared__ __align__(16) unsigned char raw[];
  SliceSmem& sm = *reinterpret_cast<SliceSmem*>(raw);
  const SliceDesc sd = slices[list[blockIdx.x]];
  analyse_slice(sm, row_ptr, col, sd.row0, sd.nrows, max_gran);
  const int tid = threadIdx.x;
  const int32_t gmin = sm.gmin;
  const int32_t words = sm.gmax < gmin ? 0 : ((sm.gmax - gmin) >> 5) + 1;
  // runs: a bit starts a run if the bit below is clear; its rank among run starts = popcount of
  // earlier starts (small: recount serially per word owner), its length = distance to the next clear bit.
  // Deterministic order: thread 0 of warp 0 walks the words (<= 8192) — cheap next to the nnz pass.
  if (tid == 0) {
    int32_t r = 0;
    int32_t open_start = -1;
    for (int32_t w = 0; w < words; w++) {
      uint32_t bits = sm.bitmap[w];
      if (bits == 0u) {
        if (open_start >= 0) {
          Run q; q.col0 = (gmin + open_start) << kGranuleShift; q.len = ((w << 5) - open_start) << kGranuleShift;
          q.local_base = 0; q.pad_ = 0; runs[sd.run_off + r++] = q; open_start = -1;
        }
        continue;
      }
      if (bits == 0xffffffffu) { if (open_start < 0) open_start = w << 5; continue; }
      for (int b = 0; b < 32; b++) {
        const bool set = (bits >> b) & 1u;
        if (set && open_start < 0) open_start = (w << 5) + b;
        if (!set && open_start >= 0) {
          Run q; q.col0 = (gmin + open_start) << kGranuleShift; q.len = (((w << 5) + b) - open_start) << kGranuleShift;
          q.local_base = 0; q.pad_ = 0; runs[sd.run_off + r++] = q; open_start = -1;
        }
      }
    }
    if (open_start >= 0) {
      Run q; q.col0 = (gmin + open_start) << kGranuleShift; q.len = ((words << 5) - open_start) << kGranuleShift;
      q.local_base = 0; q.pad_ = 0; runs[sd.run_off + r++] = q;
    }
    // local bases and clipping at m
    int32_t base = kZeroSlots;
    for (int32_t i = 0; i < r; i++) {
      Run q = runs[sd.run_off + i];
      q.local_base = base;
      base += q.len;
      if (q.col0 + q.len > m) q.len = m - q.col0;
      runs[sd.run_off + i] = q;
      if (r <= kInlineRuns) slices[list[blockIdx.x]].inl[i] = q;
    }
  }
  // ELL fill: one thread per row walks its entries; padding entries point at the zero slot.
  const int32_t T = kEllThreads, RPT = kEllRowsPerThread;
  for (int32_t row = tid; row < T * RPT; row += blockDim.x) {
    const int32_t t = row % T, j = row / T;
    int32_t k = 0, kb = 0, ke = 0;
    if (row < sd.nrows) { kb = row_ptr[sd.row0 + row]; ke = row_ptr[sd.row0 + row + 1]; }
    for (; k < sd.width; k++) {
      const int64_t pos = sd.val_off + ((int64_t)k * T + t) * RPT + j;
      if (kb + k < ke) {
        const int32_t c = col[kb + k];
        const int32_t g = (c >> kGranuleShift) - gmin;
        const int32_t rank = sm.prefix[g >> 5] + __popc(sm.bitmap[g >> 5] & ((1u << (g & 31)) - 1u));
        ell_vals[pos] = val[kb + k];
        ell_idx[pos] = (uint16_t)(kZeroSlots + rank * kGranule + (c & (kGranule - 1)));
      } else {
        ell_vals[pos] = 0.0;
        ell_idx[pos] = 0;
      }
    }
  }
}

__global__ void row_length_histogram_kernel(const int32_t* __restrict__ row_ptr, int64_t n,
                                            unsigned long long* __restrict__ hist, int32_t* __restrict__ maxlen) {
  __shared__ unsigned int sh[8];
  if (threadIdx.x < 8) sh[threadIdx.x] = 0;
  __syncthreads();
  int32_t lm = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t len = row_ptr[i + 1] - row_ptr[i];
    lm = max(lm, len);
    int b = len == 0 ? 0 : len <= 2 ? 1 : len <= 4 ? 2 : len <= 8 ? 3 : len <= 16 ? 4 : len <= 32 ? 5 : len <= 64 ? 6 : 7;
    atomicAdd(&sh[b], 1u);
  }
  lm = warp_max(lm);
  if ((threadIdx.x & 31) == 0) atomicMax(maxlen, lm);
  __syncthreads();
  if (threadIdx.x < 8) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// Merge-path partition of the gather runs, one thread per tile boundary: tile g of run q starts on diagonal
// g * kMergeTile of the run's merge grid (rows of the run x its nonzeros).  The number of row ends consumed before
// diagonal d is #{ i : row_ptr[ra + i + 1] - ka <= d - i - 1 }, a prefix of the rows, found by binary search.
struct MergeRun { int32_t ra, rb, ka, kb, tile0, ntiles; };

__global__ void merge_tiles_kernel(const MergeRun* __restrict__ runs, int nruns, int total_tiles, int tile_items,
                                   const int32_t* __restrict__ row_ptr, MergeTile* __restrict__ tiles) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_tiles) return;
  int q = 0;
  {
    int lo = 0, hi = nruns - 1;  // last run whose tile0 <= g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (runs[mid].tile0 <= g) lo = mid; else hi = mid - 1; }
    q = lo;
  }
  const MergeRun run = runs[q];
  const int64_t nrows = run.rb - run.ra, nnz = run.kb - run.ka, total = nrows + nnz;
  auto split = [&](int64_t diag, int32_t* r_out, int32_t* k_out) {
    diag = diag < total ? diag : total;
    int64_t lo = diag > nnz ? diag - nnz : 0, hi = diag < nrows ? diag : nrows;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)row_ptr[run.ra + mid + 1] - run.ka <= diag - mid - 1) lo = mid + 1; else hi = mid;
    }
    *r_out = (int32_t)(run.ra + lo);
    *k_out = (int32_t)(run.ka + (diag - lo));
  };
  const int64_t i = g - run.tile0;
  MergeTile t;
  split(i * tile_items, &t.r0, &t.k0);
  split((i + 1) * tile_items, &t.r1, &t.k1);
  tiles[g] = t;
}

// CSR sanity before anything indexes with it (plan_count_kernel's bitmap, the gathers): row_ptr starts at 0, ends at nnz
// and never decreases; every column lies in [0, m).  *bad = 1 + the first kind of violation seen (any thread may win).
__global__ void __launch_bounds__(256) validate_csr_kernel(const int32_t* __restrict__ row_ptr, int64_t n, int64_t nnz,
                                                           const int32_t* __restrict__ col, int64_t m, int32_t* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t r = t0; r <= n; r += stride) {
    const int32_t v = row_ptr[r];
    if ((r == 0 && v != 0) || (r == n && v != nnz) || v < 0 || v > nnz || (r < n && row_ptr[r + 1] < v)) *bad = 1;
  }
  for (int64_t k = t0; k < nnz; k += stride) {
    const int32_t c = col[k];
    if (c < 0 || c >= m) *bad = 2;
  }
}

// ---- column reordering of the gather path (hub clustering) ------------------------------------------------------------
// A power-law matrix sends most of its x gathers to a small set of hub columns that are scattered over the index space
// (R-MAT: the columns with few 1-bits), so every gather drags a 32-byte sector through L2 for 8 useful bytes, hot and cold
// entries share 128-byte lines, and the cache holds far fewer hot entries than its size suggests (ncu on scale 25: 26 % of
// the x sectors miss L2, 5.2 GB of DRAM reads beside the 6.0 GB matrix stream; profiles/r2c_spmv_merge_rmat_ncu.md).
// The plan therefore renumbers the columns by descending reference count (stable: equal counts keep their order, never
// referenced columns go last): y = A x = (A P^T)(P x).  The gather kernel reads a relabelled copy of the column array and a
// permuted copy of x that a streaming kernel writes in front of every SpMV (only the referenced columns).  Hubs then share
// lines (L1 hits), every sector the cache holds is full of hot entries, and the cold tail is a compacted stream.
__global__ void __launch_bounds__(256) col_count_kernel(const int32_t* __restrict__ col, int64_t nnz, int32_t* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(counts + col[k], 1);
}
// Sort keys of the three numberings; never referenced columns get kColKeyEmpty and go last.
//   by_count, no owners   ascending key = descending count (hub clustering)
//   owners                (owner of the column) << 32 | (descending count): the compact numbering of a sharded plan - one
//                         segment per owning rank, the segment's hubs first
//   neither               referenced columns in column order
constexpr uint64_t kColKeyEmpty = 64ull << 32;
constexpr int kMaxOwners = 64;
struct ColOwners {
  int32_t world = 0;              // 0: no owner grouping
  int64_t first[kMaxOwners + 1];  // first column of every owner (+ m)
};
__global__ void __launch_bounds__(256) col_keys_kernel(const int32_t* __restrict__ counts, int64_t m, uint64_t* __restrict__ keys,
                                                       uint32_t* __restrict__ ids, int by_count, const ColOwners own) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    const int32_t c = counts[j];
    uint64_t key = kColKeyEmpty;
    if (c) {
      key = by_count ? (uint64_t)(0x7fffffff - c) : 0;
      if (own.world) {
        int q = 0;
        while (q + 1 < own.world && j >= own.first[q + 1]) q++;
        key |= (uint64_t)q << 32;
      }
    }
    keys[j] = key;
    ids[j] = (uint32_t)j;
  }
}
// inv[old column] = new column; *used = number of columns referenced at least once (they are the first *used new columns)
__global__ void __launch_bounds__(256) col_inverse_kernel(const uint32_t* __restrict__ perm, const uint64_t* __restrict__ keys,
                                                          int64_t m, int32_t* __restrict__ inv, int32_t* __restrict__ used) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    inv[perm[j]] = (int32_t)j;
    const bool ref = keys[j] != kColKeyEmpty;
    if (ref && (j == m - 1 || keys[j + 1] == kColKeyEmpty)) *used = (int32_t)(j + 1);
  }
}
// *sum += reference counts of the first k columns of the hub ordering (keys = 0x7fffffff - count, sorted ascending)
__global__ void __launch_bounds__(256) col_head_count_kernel(const uint64_t* __restrict__ keys, int64_t k, unsigned long long* __restrict__ sum) {
  unsigned long long acc = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < k; j += stride) acc += 0x7fffffffull - (keys[j] & 0xffffffffull);
#pragma unroll
  for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum, acc);
}
__global__ void __launch_bounds__(256) col_relabel_kernel(const int32_t* __restrict__ col, int64_t nnz, const int32_t* __restrict__ inv,
                                                          int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) out[k] = inv[col[k]];
}

}  // namespace

void free_plan(cask_b200_ctx* ctx) {
  Plan& p = ctx->plan;
  if (p.owns_csr) {
    cudaFree((void*)p.d_row_ptr);
    cudaFree((void*)p.d_col);
    cudaFree((void*)p.d_val);
  }
  cudaFree(p.d_slices); cudaFree(p.d_runs); cudaFree(p.d_ell_vals); cudaFree(p.d_ell_idx);
  cudaFree(p.d_ell_codes); cudaFree(p.d_ell_dict); cudaFree(p.d_ell_pairs);
  cudaFree(p.d_list_ell); cudaFree(p.d_list_csr);
  cudaFree(p.d_csr_items); cudaFree(p.d_split_rows); cudaFree(p.d_csr_scratch);
  cudaFree(p.d_merge_tiles); cudaFree(p.d_merge_carry);
  cudaFree(p.d_col_perm); cudaFree(p.d_perm);
  if (p.xperm_owned) cudaFree(p.d_xperm);
  p = Plan();
}

// Cuts the rows of the gather-CSR slices (in list order) into work items of bounded nonzero count.
// Rows of >= kCsrLongRow nonzeros get CTAs of their own; beyond kCsrSegment they are split into segments
// whose partial sums meet again, in order, in a fix-up kernel (deterministic, no atomics).
// Merge-path tiles for the gather slices, built ON THE DEVICE (the replacement of Spmv::preprocess's host loops,
// src/runtime/Spmv.cpp:329-365, for irregular matrices): the host only walks the slice list (32 K entries for R-MAT
// scale 25, nonzero counts per slice are already known from plan_count_kernel) to find the runs of consecutive gather
// slices; every tile boundary is then one binary search over row_ptr by one GPU thread.  Runs never cross the
// interior / halo-dependent split of the list, so either part can be launched on its own.
static int build_merge_tiles(cask_b200_ctx* ctx, int items_override = 0) {
  Plan& p = ctx->plan;
  cudaFree(p.d_merge_tiles); cudaFree(p.d_merge_carry);
  p.d_merge_tiles = nullptr; p.d_merge_carry = nullptr;
  cudaStream_t s = ctx->stream;
  std::vector<int64_t> k_at(p.nslices + 1, 0);  // row_ptr at every slice boundary = prefix sum of the slices' nonzeros
  for (int32_t i = 0; i < p.nslices; i++) k_at[i + 1] = k_at[i] + p.h_slices[i].nnz;
  std::vector<MergeRun> runs;
  p.merge_items = ctx->merge_items == 5 || ctx->merge_items == 7 || ctx->merge_items == 11 || ctx->merge_items == 17
                      ? ctx->merge_items : items_override ? items_override : kMergeItemsDefault;
  const int64_t tile_items = (int64_t)kMergeThreads * p.merge_items;
  p.h_item_begin.assign((size_t)p.n_csr + 1, -1);
  p.h_split_begin.assign((size_t)p.n_csr + 1, 0);
  int32_t tiles = 0;
  for (int32_t pos = 0; pos < p.n_csr;) {
    int32_t end = pos + 1;
    while (end < p.n_csr && end != p.n_csr_interior && p.h_list_csr[end] == p.h_list_csr[end - 1] + 1) end++;
    const SliceDesc& a = p.h_slices[p.h_list_csr[pos]];
    const SliceDesc& b = p.h_slices[p.h_list_csr[end - 1]];
    MergeRun r;
    r.ra = a.row0; r.rb = b.row0 + b.nrows;
    r.ka = (int32_t)k_at[p.h_list_csr[pos]]; r.kb = (int32_t)k_at[p.h_list_csr[end - 1] + 1];
    const int64_t total = (int64_t)(r.rb - r.ra) + (r.kb - r.ka);
    r.tile0 = tiles;
    r.ntiles = (int32_t)((total + tile_items - 1) / tile_items);
    p.h_item_begin[pos] = tiles;
    tiles += r.ntiles;
    runs.push_back(r);
    pos = end;
  }
  p.h_item_begin[p.n_csr] = tiles;
  p.n_merge_tiles = tiles;
  CB_CUDA(cudaMalloc(&p.d_merge_tiles, sizeof(MergeTile) * std::max(tiles, 1)));
  CB_CUDA(cudaMalloc(&p.d_merge_carry, sizeof(double) * std::max(tiles, 1)));
  if (tiles) {
    MergeRun* d_runs = nullptr;
    CB_CUDA(cudaMalloc(&d_runs, sizeof(MergeRun) * runs.size()));
    CB_CUDA(cudaMemcpyAsync(d_runs, runs.data(), sizeof(MergeRun) * runs.size(), cudaMemcpyHostToDevice, s));
    merge_tiles_kernel<<<(tiles + 255) / 256, 256, 0, s>>>(d_runs, (int)runs.size(), tiles, (int)tile_items, p.d_row_ptr, p.d_merge_tiles);
    ctx->launches++;
    CB_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_runs);
    CB_CUDA(cudaGetLastError());
  }
  return CASK_B200_OK;
}

int build_csr_items(cask_b200_ctx* ctx) {
  Plan& p = ctx->plan;
  cudaFree(p.d_csr_items); cudaFree(p.d_split_rows); cudaFree(p.d_csr_scratch);
  cudaFree(p.d_merge_tiles); cudaFree(p.d_merge_carry);
  p.d_csr_items = nullptr; p.d_split_rows = nullptr; p.d_csr_scratch = nullptr;
  p.d_merge_tiles = nullptr; p.d_merge_carry = nullptr;
  p.n_merge_tiles = 0;
  p.csr_merge = false;
  p.h_item_begin.assign(1, 0);
  p.h_split_begin.assign(1, 0);
  if (p.n_csr == 0) return CASK_B200_OK;
  int64_t gather_nnz = 0;
  for (int32_t pos = 0; pos < p.n_csr; pos++) gather_nnz += p.h_slices[p.h_list_csr[pos]].nnz;
  p.csr_merge = ctx->csr_kernel == 1 || (ctx->csr_kernel < 0 && !ctx->csr_stream && gather_nnz >= kMergeAutoNnz);
  p.stats.csr_kernel = p.csr_merge ? 1 : 0;
  if (p.csr_merge) {
    CB_TRY(build_merge_tiles(ctx));
    p.stats.csr_items = p.n_merge_tiles;
    return CASK_B200_OK;
  }
  cudaStream_t s = ctx->stream;
  std::vector<int32_t> rp((size_t)p.n + 1);
  CB_CUDA(cudaMemcpyAsync(rp.data(), p.d_row_ptr, sizeof(int32_t) * (p.n + 1), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  std::vector<CsrItem> items;
  std::vector<SplitRow> splits;
  int32_t n_scratch = 0;
  auto pick_vec = [&](int64_t nnz, int32_t rows) {
    if (ctx->force_csr_vec) return ctx->force_csr_vec;
    int32_t v = 2;
    const double mean = rows ? (double)nnz / rows : 0.0;
    while (v < 32 && v * 2 <= mean) v <<= 1;
    return v;
  };
  p.csr_stream = ctx->csr_stream != 0;
  p.csr_item_nnz = p.csr_stream ? std::max(1024, std::min(ctx->csr_item_nnz, 16384)) : kCsrItemNnz;
  const int32_t item_target = p.csr_item_nnz;
  for (int32_t pos = 0; pos < p.n_csr; pos++) {
    const SliceDesc& sd = p.h_slices[p.h_list_csr[pos]];
    int32_t start = sd.row0;
    auto flush = [&](int32_t end) {
      if (end > start) {
        CsrItem it{start, end - start, rp[start], rp[end], -1, pick_vec(rp[end] - rp[start], end - start), {0, 0}};
        items.push_back(it);
      }
      start = end;
    };
    for (int32_t r = sd.row0; r < sd.row0 + sd.nrows; r++) {
      const int32_t len = rp[r + 1] - rp[r];
      if (len >= kCsrLongRow) {
        flush(r);
        if (len <= kCsrSegment) {
          items.push_back(CsrItem{r, 1, rp[r], rp[r + 1], -1, 0, {0, 0}});
        } else {
          const int32_t nseg = (len + kCsrSegment - 1) / kCsrSegment;
          splits.push_back(SplitRow{r, n_scratch, nseg, 0});
          for (int32_t g = 0; g < nseg; g++)
            items.push_back(CsrItem{r, 1, rp[r] + g * kCsrSegment, std::min(rp[r + 1], rp[r] + (g + 1) * kCsrSegment),
                                    n_scratch++, 0, {0, 0}});
        }
        start = r + 1;
      } else if (rp[r + 1] - rp[start] >= item_target) {
        flush(r + 1);
      }
    }
    flush(sd.row0 + sd.nrows);
    p.h_item_begin.push_back((int32_t)items.size());
    p.h_split_begin.push_back((int32_t)splits.size());
  }
  CB_CUDA(cudaMalloc(&p.d_csr_items, sizeof(CsrItem) * std::max<size_t>(items.size(), 1)));
  CB_CUDA(cudaMemcpyAsync(p.d_csr_items, items.data(), sizeof(CsrItem) * items.size(), cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMalloc(&p.d_split_rows, sizeof(SplitRow) * std::max<size_t>(splits.size(), 1)));
  if (!splits.empty())
    CB_CUDA(cudaMemcpyAsync(p.d_split_rows, splits.data(), sizeof(SplitRow) * splits.size(), cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMalloc(&p.d_csr_scratch, sizeof(double) * std::max(n_scratch, 1)));
  CB_CUDA(cudaStreamSynchronize(s));
  p.stats.csr_items = (int32_t)items.size();
  return CASK_B200_OK;
}

// Hub clustering of the gather path (see col_count_kernel above).  Measured on R-MAT scale 25
// (profiles/r2k_rmat_reorder.md): the gather phase alone falls from 3.04 to 2.36 ms, DRAM reads from 9.6 to 7.8 GB, but with
// the default tile shape the whole SpMV only from 3.21 to 3.17 ms - with the hubs served by L1 the kernel becomes bound by
// the LSU wavefront rate its reduction phases share with the gathers, and the permutation of x costs 0.12 ms per call.
// Smaller tiles and more resident CTAs recover part of it (profiles/r2s_rmat_ctas.md).
//
// mode -1: hub clustering if it pays (the automatic rule of a single-rank plan): the merge-path kernel runs at least
// kReorderAutoNnz nonzeros, x is larger than kReorderAutoXBytes (below that every gather hits L2 anyway), and the matrix IS
// skewed - the most referenced 1/64 of the referenced columns draw at least a quarter of all gathers (R-MAT: about half;
// uniformly scattered columns: 1/64, for which the permutation of x would be pure cost).  The tiles are then re-cut for
// 5 merge items per thread and the kernel compiled for 8 resident CTAs per SM: measured on R-MAT scale 25
// (profiles/r2s_rmat_ctas.md) 2.99 ms per SpMV, permutation included, against 3.23 ms without clustering.
// mode 1: hub clustering as above (option col_reorder).  mode 2 (dist.cu, sparse exchange of a row-sharded gather plan):
// the referenced columns only - the compact local numbering of a distributed SpMV - grouped by owning rank, hubs first
// inside a group (owner_first given), or in ascending column order (single-rank option); in a sharded plan the permuted x
// is filled by the exchange (own columns by a pack kernel, the others received from their owners), not by permute_x_kernel.
int build_col_reorder(cask_b200_ctx* ctx, int mode, const int64_t* owner_first, int world) {
  Plan& p = ctx->plan;
  cudaFree(p.d_col_perm); cudaFree(p.d_perm);
  if (p.xperm_owned) cudaFree(p.d_xperm);
  p.d_col_perm = nullptr; p.d_perm = nullptr; p.d_xperm = nullptr;
  p.xperm_owned = true;
  p.cols_used = 0;
  p.xperm_external = false;
  p.hub_ordered = false;
  p.stats.col_reorder = 0;
  p.stats.cols_referenced = 0;
  p.merge_ctas = 0;
  if (mode == 0 || p.m <= 0 || p.m >= (1ll << 30)) return CASK_B200_OK;
  const bool automatic = mode < 0;
  if (automatic) {
    if (dist_active(ctx) || !p.csr_merge || p.stats.csr_nnz < kReorderAutoNnz || p.m * 8 < kReorderAutoXBytes) return CASK_B200_OK;
    mode = 1;
  }
  if (mode == 1 && (!p.csr_merge || p.nnz <= 0)) return CASK_B200_OK;
  if (mode == 2 && !ctx->dist_sparse_active && (!p.csr_merge || p.nnz <= 0)) return CASK_B200_OK;  // as an option: gather plans only
  cudaStream_t s = ctx->stream;
  dev::Exec ex = dev::exec_of(ctx);
  struct Tmp {
    void* q[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Tmp() { for (void* x : q) cudaFree(x); }
  } tmp;
  const int64_t m = p.m;
  const int grid = ctx->sm_count * 8;
  CB_CUDA(cudaMalloc(&tmp.q[0], sizeof(int32_t) * (size_t)m));       // counts, then the inverse permutation
  CB_CUDA(cudaMalloc(&tmp.q[1], sizeof(uint64_t) * (size_t)m));      // keys
  CB_CUDA(cudaMalloc(&tmp.q[2], sizeof(uint64_t) * (size_t)m));      // sorted keys
  CB_CUDA(cudaMalloc(&tmp.q[3], sizeof(uint32_t) * (size_t)m));      // column ids
  CB_CUDA(cudaMalloc(&tmp.q[4], sizeof(int32_t)));                   // referenced columns
  CB_CUDA(cudaMalloc(&p.d_perm, sizeof(int32_t) * (size_t)m));
  CB_CUDA(cudaMemsetAsync(tmp.q[0], 0, sizeof(int32_t) * (size_t)m, s));
  CB_CUDA(cudaMemsetAsync(tmp.q[4], 0, sizeof(int32_t), s));
  if (p.nnz) col_count_kernel<<<grid, 256, 0, s>>>(p.d_col, p.nnz, (int32_t*)tmp.q[0]);
  // sharded compact numbering: one segment per owning rank, hubs first inside a segment (measured on the stripes of the
  // 8- and 2-rank C3 jobs run alone, profiles/r2u2_stripe_hub.md: 7-12 % faster than column order)
  ColOwners own;
  const bool grouped = mode == 2 && owner_first && world >= 1 && world <= kMaxOwners;
  if (grouped) {
    own.world = world;
    for (int q = 0; q <= world; q++) own.first[q] = owner_first[q];
  }
  col_keys_kernel<<<grid, 256, 0, s>>>((const int32_t*)tmp.q[0], m, (uint64_t*)tmp.q[1], (uint32_t*)tmp.q[3], mode == 1 || grouped ? 1 : 0, own);
  ctx->launches += 2;
  CB_TRY(dev::sort_pairs_u64_u32(ex, (const uint64_t*)tmp.q[1], (uint64_t*)tmp.q[2], (const uint32_t*)tmp.q[3],
                                 (uint32_t*)p.d_perm, m, 39));  // stable: equal keys keep ascending column order
  col_inverse_kernel<<<grid, 256, 0, s>>>((const uint32_t*)p.d_perm, (const uint64_t*)tmp.q[2], m, (int32_t*)tmp.q[0],
                                          (int32_t*)tmp.q[4]);
  CB_CUDA(cudaMalloc(&p.d_col_perm, sizeof(int32_t) * (size_t)std::max<int64_t>(p.nnz, 1)));
  if (p.nnz) col_relabel_kernel<<<grid, 256, 0, s>>>(p.d_col, p.nnz, (const int32_t*)tmp.q[0], p.d_col_perm);
  ctx->launches += 2;
  int32_t used = 0;
  CB_CUDA(cudaMemcpyAsync(&used, tmp.q[4], sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaGetLastError());
  if (automatic) {
    // skew test: share of all gathers that go to the most referenced 1/64 of the referenced columns
    unsigned long long head = 0;
    CB_CUDA(cudaMemsetAsync(tmp.q[1], 0, sizeof(unsigned long long), s));  // the unsorted keys are no longer needed
    const int64_t k = std::max<int64_t>(1, used / 64);
    col_head_count_kernel<<<grid, 256, 0, s>>>((const uint64_t*)tmp.q[2], k, (unsigned long long*)tmp.q[1]);
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(&head, tmp.q[1], sizeof(head), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if ((double)head < 0.25 * (double)p.nnz) {
      cudaFree(p.d_col_perm); cudaFree(p.d_perm);
      p.d_col_perm = nullptr; p.d_perm = nullptr;
      return CASK_B200_OK;
    }
    if (!(ctx->merge_items == 5 || ctx->merge_items == 7 || ctx->merge_items == 11 || ctx->merge_items == 17)) {
      CB_TRY(build_merge_tiles(ctx, 5));
      p.stats.csr_items = p.n_merge_tiles;
      p.merge_ctas = 8;
    }
  }
  if (grouped && p.csr_merge && p.stats.csr_nnz >= kReorderAutoNnz &&
      !(ctx->merge_items == 5 || ctx->merge_items == 7 || ctx->merge_items == 11 || ctx->merge_items == 17)) {
    CB_TRY(build_merge_tiles(ctx, 5));  // the configuration of the hub-clustered single-rank plans
    p.stats.csr_items = p.n_merge_tiles;
    p.merge_ctas = 8;
    p.stats.merge_items = p.merge_items;
    p.stats.merge_ctas = p.merge_ctas;
  }
  p.hub_ordered = mode == 1 || grouped;
  p.cols_used = used;
  CB_CUDA(cudaMalloc(&p.d_xperm, sizeof(double) * (size_t)std::max<int64_t>(used, 2)));
  p.xperm_external = mode == 2 && ctx->dist_sparse_active;
  p.stats.col_reorder = mode;
  p.stats.cols_referenced = used;
  return CASK_B200_OK;
}

// Coded staged ELL (option value_dict): per-slice tables + 8-bit codes (valuedict_logic.inl).
//   value_dict = 2  (value, x-cache displacement) PAIR codes: 1 byte per stored nonzero, no index stream;
//   value_dict = 1  value codes beside the 16-bit indices: 3 bytes per stored nonzero.
// All staged slices or none: if any slice needs more than 255 pairs the plan falls back to value codes, and if any
// slice holds more than 256 distinct values it stays uncoded.  val_off_total = entries of the ELL arrays.
static int build_value_dict(cask_b200_ctx* ctx, int64_t val_off_total) {
  Plan& p = ctx->plan;
  p.coded = 0;
  p.dict_len = 0;
  if (!ctx->value_dict || p.n_ell == 0 || p.nslices == 0) return CASK_B200_OK;
  dev::Exec ex = dev::exec_of(ctx);
  std::vector<int64_t> h_off((size_t)p.nslices, 0);
  std::vector<int32_t> h_width((size_t)p.nslices, 0);  // 0 for gather-CSR slices: empty table, nothing scanned
  for (int32_t i = 0; i < p.nslices; i++)
    if (p.h_slices[i].kind == kSliceStagedEll) { h_off[i] = p.h_slices[i].val_off; h_width[i] = p.h_slices[i].width; }
  struct Tmp {
    void* q[3] = {nullptr, nullptr, nullptr};
    ~Tmp() { for (void* x : q) dev::release(x); }
  } tmp;
  CB_TRY(dev::alloc(&tmp.q[0], sizeof(int64_t) * (size_t)p.nslices));
  CB_TRY(dev::alloc(&tmp.q[1], sizeof(int32_t) * (size_t)p.nslices));
  CB_TRY(dev::alloc(&tmp.q[2], sizeof(int32_t) * (size_t)p.nslices));
  CB_TRY(dev::upload(ex, tmp.q[0], h_off.data(), sizeof(int64_t) * (size_t)p.nslices));
  CB_TRY(dev::upload(ex, tmp.q[1], h_width.data(), sizeof(int32_t) * (size_t)p.nslices));
  CB_CUDA(cudaMalloc(&p.d_ell_codes, (size_t)std::max<int64_t>(val_off_total, 16)));
  int32_t overflow = 0, max_entries = 0;
  if (ctx->value_dict >= 2) {
    CB_CUDA(cudaMalloc(&p.d_ell_pairs, sizeof(valuedict::PairEntry) * (size_t)valuedict::kPairStride * (size_t)p.nslices));
    CB_TRY(valuedict::build_pairs(ex, p.nslices, (const int64_t*)tmp.q[0], (const int32_t*)tmp.q[1], p.slice_rows, p.d_ell_vals,
                                  p.d_ell_idx, (valuedict::PairEntry*)p.d_ell_pairs, p.d_ell_codes, (int32_t*)tmp.q[2],
                                  &overflow, &max_entries));
    if (!overflow && max_entries > 0) {
      p.coded = 2;
      p.dict_len = (max_entries + 7) & ~7;  // records staged per slice (16 bytes each)
      return CASK_B200_OK;
    }
    cudaFree(p.d_ell_pairs);
    p.d_ell_pairs = nullptr;
  }
  CB_CUDA(cudaMalloc(&p.d_ell_dict, sizeof(double) * (size_t)valuedict::kStride * (size_t)p.nslices));
  CB_TRY(valuedict::build(ex, p.nslices, (const int64_t*)tmp.q[0], (const int32_t*)tmp.q[1], p.slice_rows, p.d_ell_vals,
                          p.d_ell_dict, p.d_ell_codes, (int32_t*)tmp.q[2], &overflow, &max_entries));
  if (overflow || max_entries == 0) {
    cudaFree(p.d_ell_codes); cudaFree(p.d_ell_dict);
    p.d_ell_codes = nullptr; p.d_ell_dict = nullptr;
    return CASK_B200_OK;
  }
  p.coded = 1;
  p.dict_len = (max_entries + 1) & ~1;  // bulk copies move multiples of 16 bytes
  return CASK_B200_OK;
}

int build_plan(cask_b200_ctx* ctx) {
  Plan& p = ctx->plan;
  const cask_b200_design& d = ctx->design;
  cudaStream_t s = ctx->stream;
  const int64_t n = p.n;
  p.slice_rows = kSliceRows;

  // ---- slices: stripes of Spmv::preprocess, cut every slice_rows rows -------------------------
  // (sharded runs: the stripe IS the rank's row range, Spmv.cpp:334-364 applied over ranks)
  std::vector<int32_t> row0s, nrows;
  {
    const int64_t P = d.num_pipes > 0 ? d.num_pipes : 1;
    const int64_t rpp = n / P;
    int64_t start = 0;
    for (int64_t q = 0; q < P; q++) {
      int64_t rows = rpp == 0 ? (q == 0 ? n : 0) : (q == P - 1 ? n - start : rpp);
      for (int64_t r = 0; r < rows; r += p.slice_rows) {
        row0s.push_back((int32_t)(start + r));
        nrows.push_back((int32_t)std::min<int64_t>(p.slice_rows, rows - r));
      }
      start += rows;
    }
  }
  if (n > 0) {
    int32_t* d_bad = nullptr;
    int32_t bad = 0;
    CB_CUDA(cudaMalloc(&d_bad, sizeof(int32_t)));
    cudaMemsetAsync(d_bad, 0, sizeof(int32_t), s);
    validate_csr_kernel<<<std::min<int64_t>((std::max(n + 1, p.nnz) + 255) / 256, 148 * 8), 256, 0, s>>>(p.d_row_ptr, n, p.nnz, p.d_col, p.m,
                                                                                                      d_bad);
    ctx->launches++;
    const cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(d_bad);
    CB_CUDA(e);
    CB_CUDA(e2);
    if (bad == 1) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: row_ptr must start at 0, end at nnz and never decrease");
    if (bad == 2) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: a column index lies outside [0, m)");
  }
  p.nslices = (int32_t)row0s.size();
  p.stats = cask_b200_plan_stats{};
  p.stats.n = n; p.stats.m = p.m; p.stats.nnz = p.nnz;
  p.stats.slice_rows = p.slice_rows;
  p.stats.num_slices = p.nslices;
  if (p.nslices == 0) return CASK_B200_OK;

  int32_t cache = d.cache_size;
  if (cache > kMaxCacheDoubles) cache = kMaxCacheDoubles;
  const int32_t max_gran = std::max(0, (cache - kZeroSlots) / kGranule);

  struct Scratch {  // device temporaries of this function: released on every exit path
    std::vector<void*> v;
    ~Scratch() { for (void* x : v) cudaFree(x); }
  } scratch;
  int32_t *d_row0 = nullptr, *d_nrows = nullptr;
  SliceCount* d_counts = nullptr;
  CB_CUDA(cudaMalloc(&d_row0, sizeof(int32_t) * p.nslices));
  scratch.v.push_back(d_row0);
  CB_CUDA(cudaMalloc(&d_nrows, sizeof(int32_t) * p.nslices));
  scratch.v.push_back(d_nrows);
  CB_CUDA(cudaMalloc(&d_counts, sizeof(SliceCount) * p.nslices));
  scratch.v.push_back(d_counts);
  CB_CUDA(cudaMemcpyAsync(d_row0, row0s.data(), sizeof(int32_t) * p.nslices, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMemcpyAsync(d_nrows, nrows.data(), sizeof(int32_t) * p.nslices, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaFuncSetAttribute(plan_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SliceSmem)));
  CB_CUDA(cudaFuncSetAttribute(plan_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SliceSmem)));
  plan_count_kernel<<<p.nslices, 256, sizeof(SliceSmem), s>>>(p.d_row_ptr, p.d_col, d_row0, d_nrows, max_gran, d_counts);
  ctx->launches++;

  unsigned long long* d_hist = nullptr;
  int32_t* d_maxlen = nullptr;
  CB_CUDA(cudaMalloc(&d_hist, sizeof(unsigned long long) * 8));
  scratch.v.push_back(d_hist);
  CB_CUDA(cudaMalloc(&d_maxlen, sizeof(int32_t)));
  scratch.v.push_back(d_maxlen);
  CB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 8, s));
  CB_CUDA(cudaMemsetAsync(d_maxlen, 0, sizeof(int32_t), s));
  row_length_histogram_kernel<<<std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(p.d_row_ptr, n, d_hist, d_maxlen);
  ctx->launches++;

  std::vector<SliceCount> counts(p.nslices);
  unsigned long long hist[8];
  int32_t maxlen = 0;
  CB_CUDA(cudaMemcpyAsync(counts.data(), d_counts, sizeof(SliceCount) * p.nslices, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof(hist), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(&maxlen, d_maxlen, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaGetLastError());

  // ---- kernel selection from the histogram and the band profile --------------------------------
  p.h_slices.assign(p.nslices, SliceDesc{});
  std::vector<int32_t> list_ell, list_csr;
  int64_t val_off = 0;
  int32_t run_off = 0;
  int64_t csr_rows = 0, csr_nnz = 0;
  p.max_xcache = 0;
  const int64_t stripe_lo = p.row0_global, stripe_hi = p.row0_global + p.n;
  (void)stripe_lo; (void)stripe_hi;
  for (int32_t i = 0; i < p.nslices; i++) {
    SliceDesc& sd = p.h_slices[i];
    const SliceCount& c = counts[i];
    sd.row0 = row0s[i]; sd.nrows = nrows[i]; sd.width = c.width; sd.nnz = c.nnz;
    sd.col_hi = (int32_t)std::min<int64_t>(c.col_hi, p.m);
    const double fill = c.width > 0 ? (double)c.nnz / ((double)c.width * nrows[i]) : 1.0;
    bool staged = c.nruns >= 0 && c.xlen <= cache && fill >= ctx->ell_min_fill && c.width <= 4096;
    if (ctx->force_kind == 1) staged = false;
    if (ctx->force_kind == 0) staged = c.nruns >= 0 && c.xlen <= cache;
    if (staged) {
      sd.kind = kSliceStagedEll;
      sd.val_off = val_off;
      sd.run_off = run_off;
      sd.nruns = c.nruns;
      sd.xcache_len = c.xlen;
      val_off += (int64_t)c.width * p.slice_rows;
      run_off += c.nruns;
      p.max_xcache = std::max(p.max_xcache, c.xlen);
      list_ell.push_back(i);
      p.stats.ell_padded_entries += (int64_t)c.width * p.slice_rows;
      p.stats.ell_nnz += c.nnz;
      p.stats.xcache_doubles_total += c.xlen;
    } else {
      sd.kind = kSliceGatherCsr;
      list_csr.push_back(i);
      csr_rows += nrows[i];
      csr_nnz += c.nnz;
    }
  }
  // lanes per row for the CSR slices: smallest power of two >= mean row length / 2, in [2, 32]
  int32_t vec = 2;
  if (csr_rows > 0) {
    const double mean = (double)csr_nnz / (double)csr_rows;
    while (vec < 32 && vec * 2 <= mean) vec <<= 1;
  }
  if (ctx->force_csr_vec) vec = ctx->force_csr_vec;
  p.csr_vec = vec;
  p.h_list_ell = list_ell;
  p.h_list_csr = list_csr;
  p.n_ell = (int32_t)list_ell.size();
  p.n_ell_interior = p.n_ell;
  p.n_csr = (int32_t)list_csr.size();
  p.n_csr_interior = p.n_csr;

  CB_CUDA(cudaMalloc(&p.d_slices, sizeof(SliceDesc) * p.nslices));
  CB_CUDA(cudaMemcpyAsync(p.d_slices, p.h_slices.data(), sizeof(SliceDesc) * p.nslices, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMalloc(&p.d_list_ell, sizeof(int32_t) * std::max(p.n_ell, 1)));
  CB_CUDA(cudaMalloc(&p.d_list_csr, sizeof(int32_t) * std::max(p.n_csr, 1)));
  if (p.n_ell) CB_CUDA(cudaMemcpyAsync(p.d_list_ell, list_ell.data(), sizeof(int32_t) * p.n_ell, cudaMemcpyHostToDevice, s));
  if (p.n_csr) CB_CUDA(cudaMemcpyAsync(p.d_list_csr, list_csr.data(), sizeof(int32_t) * p.n_csr, cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaMalloc(&p.d_runs, sizeof(Run) * std::max(run_off, 1)));
  CB_CUDA(cudaMalloc(&p.d_ell_vals, sizeof(double) * std::max<int64_t>(val_off, 1)));
  CB_CUDA(cudaMalloc(&p.d_ell_idx, sizeof(uint16_t) * std::max<int64_t>(val_off, 8)));
  if (p.n_ell) {
    plan_fill_kernel<<<p.n_ell, 256, sizeof(SliceSmem), s>>>(p.d_row_ptr, p.d_col, p.d_val, p.d_slices, p.d_list_ell,
                                                             max_gran, (int32_t)p.m, p.d_runs, p.d_ell_vals, p.d_ell_idx);
    ctx->launches++;
  }
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaGetLastError());

  CB_TRY(build_value_dict(ctx, val_off));
  CB_TRY(configure_persistent(ctx));
  CB_TRY(build_csr_items(ctx));
  p.stats.slices_staged_ell = p.n_ell;
  p.stats.slices_gather_csr = p.n_csr;
  p.stats.csr_lanes_per_row = vec;
  p.stats.max_row_length = maxlen;
  p.stats.csr_nnz = csr_nnz;
  p.stats.csr_rows = csr_rows;
  CB_TRY(build_col_reorder(ctx, ctx->col_reorder == 1 || ctx->col_reorder == 2 || ctx->col_reorder < 0 ? ctx->col_reorder : 0, nullptr, 0));
  p.stats.persist_ku = ctx->ell_kernel == 1 ? p.persist_ku : 0;
  p.stats.persist_stages = p.persist_stages;
  p.stats.persist_ctas_per_sm = p.persist_ctas_per_sm;
  p.stats.value_dict = p.coded;
  p.stats.merge_items = p.csr_merge ? p.merge_items : 0;
  p.stats.merge_ctas = p.csr_merge ? (p.merge_ctas ? p.merge_ctas : (p.merge_items > 11 ? 4 : 5)) : 0;
  for (int i = 0; i < 8; i++) p.stats.row_length_histogram[i] = (int64_t)hist[i];
  p.stats.device_bytes = val_off * (p.coded ? 11 : 10) + (p.coded == 1 ? (int64_t)p.nslices * valuedict::kStride * 8 : 0) +
                         (p.coded == 2 ? (int64_t)p.nslices * valuedict::kPairStride * 16 : 0) +
                         (int64_t)run_off * sizeof(Run) + (int64_t)p.nslices * sizeof(SliceDesc) +
                         (p.n_csr ? (csr_nnz * 12 + csr_rows * 4) : 0) +
                         (p.d_col_perm ? p.nnz * 4 + p.m * 4 + (int64_t)p.cols_used * 8 : 0);
  return CASK_B200_OK;
}

}  // namespace caskb200
