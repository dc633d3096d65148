// extern "C" entry points of libcask_b200.so (contract: include/cask_b200.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctx.cuh"

namespace caskb200 {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int ensure_device(cask_b200_ctx* ctx) {
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return fail(CASK_B200_ERR_NO_DEVICE, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  return CASK_B200_OK;
}

namespace {

int check_design(const cask_b200_design* d) {
  if (!d) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "design is null");
  if (d->num_pipes <= 0 || d->cache_size <= 0 || d->input_width <= 0)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "num_pipes, cache_size and input_width must be positive");
  if (d->arch != CASK_B200_ARCH_SIMPLE && d->arch != CASK_B200_ARCH_SKIPEMPTY)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "unknown arch");
  return CASK_B200_OK;
}

void reset_matrix(cask_b200_ctx* ctx) {
  free_precond(ctx);
  free_ref_partitions(ctx);
  free_plan(ctx);
  ctx->have_design = false;
}

int finish_preprocess(cask_b200_ctx* ctx) {
  const int rc = build_plan(ctx);
  if (rc != CASK_B200_OK) {  // nothing half-built survives (device temporaries of the plan, an owned CSR copy)
    reset_matrix(ctx);
    return rc;
  }
  ctx->have_design = true;
  if (dist_active(ctx)) CB_TRY(dist_plan_halo(ctx));
  return CASK_B200_OK;
}

int grow(double** buf, int64_t* len, int64_t need) {
  if (*len >= need) return CASK_B200_OK;
  cudaFree(*buf);
  *buf = nullptr;
  *len = 0;
  CB_CUDA(cudaMalloc(buf, sizeof(double) * std::max<int64_t>(need, 2)));
  *len = need;
  return CASK_B200_OK;
}

}  // namespace
}  // namespace caskb200

using namespace caskb200;

extern "C" {

const char* cask_b200_last_error(void) { return g_last_error.c_str(); }

int cask_b200_device_count(int* count) {
  if (!count) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "device_count: null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *count = 0; return fail(CASK_B200_ERR_NO_DEVICE, cudaGetErrorString(e)); }
  *count = c;
  return CASK_B200_OK;
}

int cask_b200_create(cask_b200_ctx** out, int device) {
  if (!out) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "create: null output");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(CASK_B200_ERR_NO_DEVICE, std::string("no CUDA device: cask_b200 has no CPU fallback (") +
                                             (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") + ")");
  if (device < 0 || device >= count) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "create: device index out of range");
  cask_b200_ctx* ctx = new cask_b200_ctx();
  ctx->device = device;
  if (ensure_device(ctx) != CASK_B200_OK) { delete ctx; return CASK_B200_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return fail(CASK_B200_ERR_CUDA, "create: stream/event creation failed");
  }
  ctx->stream = ctx->own_stream;
  // A/B switches for profiling sessions (same meaning as cask_b200_set_option)
  if (const char* e = getenv("CASK_B200_ELL_KERNEL")) ctx->ell_kernel = atoi(e);
  if (const char* e = getenv("CASK_B200_PERSIST_KU")) ctx->persist_ku = atoi(e);
  if (const char* e = getenv("CASK_B200_HOST_CHUNKS")) ctx->host_pipeline_chunks = atoi(e);
  if (const char* e = getenv("CASK_B200_HOST_STAGING")) ctx->host_staging = atoi(e);
  if (const char* e = getenv("CASK_B200_HOST_RAMP")) ctx->host_pipeline_ramp = atoi(e);
  if (const char* e = getenv("CASK_B200_PEER")) ctx->peer_mode = atoi(e);
  if (const char* e = getenv("CASK_B200_L2_KEEP")) ctx->l2_keep = atoi(e);
  if (const char* e = getenv("CASK_B200_CSR_STREAM")) ctx->csr_stream = atoi(e);
  if (const char* e = getenv("CASK_B200_CSR_ITEM_NNZ")) ctx->csr_item_nnz = atoi(e);
  if (const char* e = getenv("CASK_B200_CSR_KERNEL")) ctx->csr_kernel = atoi(e);
  if (const char* e = getenv("CASK_B200_L2_KEEP")) ctx->l2_keep = atoi(e);
  if (const char* e = getenv("CASK_B200_MERGE_ITEMS")) ctx->merge_items = atoi(e);
  if (const char* e = getenv("CASK_B200_COL_REORDER")) ctx->col_reorder = atoi(e);
  if (const char* e = getenv("CASK_B200_DIST_SPARSE")) ctx->dist_sparse = atoi(e);
  if (const char* e = getenv("CASK_B200_DIST_SPARSE_HUB")) ctx->dist_sparse_hub = atoi(e);
  if (const char* e = getenv("CASK_B200_ILU_GRAPH")) ctx->ilu_graph = atoi(e);
  if (const char* e = getenv("CASK_B200_ILU_PERSISTENT")) ctx->ilu_persistent = atoi(e);
  if (const char* e = getenv("CASK_B200_VALUE_DICT")) ctx->value_dict = atoi(e);
  if (const char* e = getenv("CASK_B200_PERSIST_CTAS")) ctx->persist_ctas = atoi(e);
  *out = ctx;
  return CASK_B200_OK;
}

int cask_b200_destroy(cask_b200_ctx* ctx) {
  if (!ctx) return CASK_B200_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  dist_free(ctx);
  free_solver_work(ctx);
  reset_matrix(ctx);
  cudaFree(ctx->d_x);
  cudaFree(ctx->d_y);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
  for (auto e : ctx->ring_events) cudaEventDestroy(e);
  delete ctx->copy_pool;
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  for (auto e : ctx->pipe_events) cudaEventDestroy(e);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
  return CASK_B200_OK;
}

int cask_b200_set_stream(cask_b200_ctx* ctx, void* cuda_stream) {
  if (!ctx) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "set_stream: null context");
  ctx->stream = (cudaStream_t)cuda_stream;  // NULL is CUDA's legacy default stream, a valid choice
  return CASK_B200_OK;
}

int cask_b200_use_own_stream(cask_b200_ctx* ctx) {
  if (!ctx) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "use_own_stream: null context");
  ctx->stream = ctx->own_stream;
  return CASK_B200_OK;
}

int cask_b200_synchronize(cask_b200_ctx* ctx) {
  if (!ctx) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "synchronize: null context");
  CB_TRY(ensure_device(ctx));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  return peer_check_error(ctx);  // a peer that stopped participating surfaces here as an error code (no-op on one rank)
}

int cask_b200_set_option(cask_b200_ctx* ctx, const char* name, double value) {
  if (!ctx || !name) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "set_option: null");
  const std::string k(name);
  if (k == "ell_min_fill") ctx->ell_min_fill = value;
  else if (k == "force_kind") ctx->force_kind = (int32_t)value;
  else if (k == "force_csr_vec") ctx->force_csr_vec = (int32_t)value;
  else if (k == "ell_kernel") ctx->ell_kernel = (int32_t)value;
  else if (k == "host_pipeline_chunks") ctx->host_pipeline_chunks = (int32_t)value;
  else if (k == "host_staging") ctx->host_staging = (int32_t)value;
  else if (k == "persist_ku") ctx->persist_ku = (int32_t)value;
  else if (k == "peer_mode") ctx->peer_mode = (int32_t)value;
  else if (k == "l2_keep") ctx->l2_keep = (int32_t)value;
  else if (k == "csr_stream") ctx->csr_stream = (int32_t)value;
  else if (k == "csr_item_nnz") ctx->csr_item_nnz = (int32_t)value;
  else if (k == "csr_kernel") ctx->csr_kernel = (int32_t)value;
  else if (k == "merge_items") ctx->merge_items = (int32_t)value;
  else if (k == "col_reorder") ctx->col_reorder = (int32_t)value;
  else if (k == "dist_sparse") ctx->dist_sparse = (int32_t)value;
  else if (k == "dist_sparse_hub") ctx->dist_sparse_hub = (int32_t)value;
  else if (k == "ilu_graph") ctx->ilu_graph = (int32_t)value;
  else if (k == "ilu_persistent") ctx->ilu_persistent = (int32_t)value;
  else if (k == "value_dict") ctx->value_dict = (int32_t)value;
  else if (k == "persist_ctas") ctx->persist_ctas = (int32_t)value;
  else return fail(CASK_B200_ERR_INVALID_ARGUMENT, "set_option: unknown option " + k);
  return CASK_B200_OK;
}

int cask_b200_launch_count(cask_b200_ctx* ctx, int64_t* count) {
  if (!ctx || !count) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "launch_count: null");
  *count = ctx->launches;
  return CASK_B200_OK;
}

// ---- preprocess -------------------------------------------------------------------------------
int cask_b200_preprocess_shard_device(cask_b200_ctx* ctx, const cask_b200_design* design, int64_t n_global,
                                      int64_t m, int64_t row0, int64_t nrows, int64_t nnz_local,
                                      const int32_t* d_row_ptr, const int32_t* d_col_ind, const double* d_values) {
  if (!ctx) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: null context");
  CB_TRY(check_design(design));
  CB_TRY(ensure_device(ctx));
  if (n_global < 0 || m < 0 || nrows < 0 || row0 < 0 || row0 + nrows > n_global || nnz_local < 0)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: bad dimensions");
  if (nrows > INT32_MAX - 2 * kSliceRows || m > INT32_MAX - 64 || nnz_local > INT32_MAX)
    return fail(CASK_B200_ERR_UNSUPPORTED, "preprocess: a stripe is limited to 2^31 rows / columns / nonzeros");
  if (nrows && (!d_row_ptr || (nnz_local && (!d_col_ind || !d_values))))
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: null CSR arrays");
  reset_matrix(ctx);
  ctx->design = *design;
  Plan& p = ctx->plan;
  p.n = nrows; p.m = m; p.nnz = nnz_local; p.n_global = n_global; p.row0_global = row0;
  p.d_row_ptr = d_row_ptr; p.d_col = d_col_ind; p.d_val = d_values; p.owns_csr = false;
  if (dist_active(ctx)) {
    // one stripe per rank: the rank IS the pipe
    ctx->design.num_pipes = 1;
  }
  return finish_preprocess(ctx);
}

int cask_b200_preprocess_device(cask_b200_ctx* ctx, const cask_b200_design* design, int64_t n, int64_t m, int64_t nnz,
                                const int32_t* d_row_ptr, const int32_t* d_col_ind, const double* d_values) {
  return cask_b200_preprocess_shard_device(ctx, design, n, m, 0, n, nnz, d_row_ptr, d_col_ind, d_values);
}

int cask_b200_preprocess(cask_b200_ctx* ctx, const cask_b200_design* design, int64_t n, int64_t m, int64_t nnz,
                         const int32_t* row_ptr, const int32_t* col_ind, const double* values) {
  if (!ctx) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: null context");
  CB_TRY(check_design(design));
  CB_TRY(ensure_device(ctx));
  if (n < 0 || m < 0 || nnz < 0 || n > INT32_MAX - 2 * kSliceRows || nnz > INT32_MAX)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: bad dimensions");
  if (!row_ptr || (nnz && (!col_ind || !values))) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: null CSR arrays");
  if (row_ptr[n] != nnz) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess: row_ptr[n] != nnz");
  reset_matrix(ctx);
  // the plan owns the device copy from the first allocation on: an early return below leaks nothing (free_plan)
  Plan& p = ctx->plan;
  p.owns_csr = true;
  int32_t *d_rp = nullptr, *d_ci = nullptr;
  double* d_va = nullptr;
  CB_CUDA(cudaMalloc(&d_rp, sizeof(int32_t) * (n + 1)));
  p.d_row_ptr = d_rp;
  CB_CUDA(cudaMalloc(&d_ci, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
  p.d_col = d_ci;
  CB_CUDA(cudaMalloc(&d_va, sizeof(double) * std::max<int64_t>(nnz, 1)));
  p.d_val = d_va;
  CB_CUDA(cudaMemcpyAsync(d_rp, row_ptr, sizeof(int32_t) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    CB_CUDA(cudaMemcpyAsync(d_ci, col_ind, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaMemcpyAsync(d_va, values, sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream));
  }
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->design = *design;
  p.n = n; p.m = m; p.nnz = nnz; p.n_global = n; p.row0_global = 0;
  return finish_preprocess(ctx);
}

int cask_b200_plan_get_stats(cask_b200_ctx* ctx, cask_b200_plan_stats* out) {
  if (!ctx || !out || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "plan_get_stats: preprocess first");
  *out = ctx->plan.stats;
  return CASK_B200_OK;
}

int cask_b200_plan_estimate(const cask_b200_plan_stats* st, double hbm_gbs, double l2_bytes, double sm_clock_hz, int32_t sms,
                            double* bytes, double* seconds) {
  if (!st) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "plan_estimate: null stats");
  if (hbm_gbs <= 0) hbm_gbs = 6456.5;
  if (l2_bytes <= 0) l2_bytes = 126.0e6;
  if (sm_clock_hz <= 0) sm_clock_hz = 1.965e9;
  if (sms <= 0) sms = 148;
  const double nnz_csr = (double)(st->nnz - st->ell_nnz);
  const double rows_csr = (double)st->slices_gather_csr * st->slice_rows;
  const double miss = std::max(0.0, 1.0 - l2_bytes / std::max(8.0 * (double)st->m, 1.0));
  const double b = 10.0 * (double)st->ell_padded_entries + 12.0 * nnz_csr + 4.0 * rows_csr + 8.0 * (double)st->m + 8.0 * (double)st->n +
                   32.0 * nnz_csr * miss;
  // a scattered 8-byte gather is one L1 wavefront per distinct line of the warp (about 28 of 32), the coalesced stream,
  // the product round trip through shared memory and the merge-path walk add 11 per 32 nonzeros; the gather kernels keep
  // the LSU pipe 0.64 busy (R-MAT scale 25: 3.2 ms measured, 1.92 ms at one wavefront per clock per SM)
  const double gather_s = nnz_csr * (39.0 / 32.0) / ((double)sms * sm_clock_hz * 0.64);
  if (bytes) *bytes = b;
  if (seconds) *seconds = std::max(b / (hbm_gbs * 1e9), gather_s);
  return CASK_B200_OK;
}

int cask_b200_plan_value_dict(cask_b200_ctx* ctx, int32_t* active, int32_t* max_entries, int64_t* matrix_bytes_per_spmv) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "plan_value_dict: preprocess first");
  const Plan& p = ctx->plan;
  const int on = p.coded && ctx->ell_kernel == 1 && p.persist_ku != 0 ? p.coded : 0;  // only the persistent kernel reads the codes
  if (active) *active = on;  // 0 uncoded, 1 value codes, 2 pair codes
  if (max_entries) *max_entries = on ? p.dict_len : 0;
  if (matrix_bytes_per_spmv) {
    const int64_t csr = p.stats.nnz - p.stats.ell_nnz;
    int64_t csr_rows = 0;
    for (const SliceDesc& sd : p.h_slices) if (sd.kind != kSliceStagedEll) csr_rows += sd.nrows;
    const int64_t entry = on == 2 ? 1 : on == 1 ? 3 : 10, table = on == 2 ? 16 : on == 1 ? 8 : 0;
    *matrix_bytes_per_spmv = p.stats.ell_padded_entries * entry + (int64_t)p.n_ell * p.dict_len * table + csr * 12 + csr_rows * 4;
  }
  return CASK_B200_OK;
}

// ---- parity hook --------------------------------------------------------------------------------
int cask_b200_partition_get_info(cask_b200_ctx* ctx, int32_t pipe, cask_b200_partition_info* out) {
  if (!ctx || !out || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "partition_get_info: preprocess first");
  CB_TRY(ensure_device(ctx));
  CB_TRY(build_ref_partitions(ctx));
  if (pipe < 0 || pipe >= (int32_t)ctx->ref_parts.size()) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "partition index out of range");
  *out = ctx->ref_parts[pipe].info;
  return CASK_B200_OK;
}

int cask_b200_partition_export(cask_b200_ctx* ctx, int32_t pipe, int32_t* colptr, void* pairs) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "partition_export: preprocess first");
  CB_TRY(ensure_device(ctx));
  CB_TRY(build_ref_partitions(ctx));
  if (pipe < 0 || pipe >= (int32_t)ctx->ref_parts.size()) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "partition index out of range");
  const RefPartition& q = ctx->ref_parts[pipe];
  if (colptr && q.info.len_colptr)
    CB_CUDA(cudaMemcpyAsync(colptr, q.d_colptr, sizeof(int32_t) * q.info.len_colptr, cudaMemcpyDeviceToHost, ctx->stream));
  if (pairs && q.info.len_pairs)
    CB_CUDA(cudaMemcpyAsync(pairs, q.d_pairs, 12 * q.info.len_pairs, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CASK_B200_OK;
}

// ---- SpMV ---------------------------------------------------------------------------------------
static int check_spmv(cask_b200_ctx* ctx) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "spmv: preprocess a matrix first");
  const cask_b200_design& d = ctx->design;
  // argument checks of Spmv::spmv, src/runtime/Spmv.cpp:189-232, same wording
  if (d.dram_reduction_enabled) {
    if (ctx->plan.n_global < 35000)
      return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Matrix is too small! Minimum supported rows with DRAM reduction: 35000 actual rows: " +
                                                      std::to_string(ctx->plan.n_global));
  } else if (d.max_rows > 0 && d.max_rows < ctx->plan.n_global) {
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Matrix is too large! Maximum supported rows: " + std::to_string(d.max_rows) +
                                                    " actual rows: " + std::to_string(ctx->plan.n_global));
  }
  if (d.num_controllers > 0) {
    if (d.num_pipes % d.num_controllers != 0) return fail(CASK_B200_ERR_RUNTIME, "numPipes should be a multiple of numControllers");
    if (d.num_controllers > d.num_pipes) return fail(CASK_B200_ERR_RUNTIME, "numPipes should be larger than numControllers");
  }
  return ensure_device(ctx);
}

int cask_b200_spmv_device(cask_b200_ctx* ctx, const double* d_x, double* d_y) {
  CB_TRY(check_spmv(ctx));
  if (!d_x || (!d_y && ctx->plan.n)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "spmv: null vector");
  if (dist_active(ctx)) {
    // d_x is the full-layout vector (global length); halo entries are refreshed in place
    const int ch = peer_channel_of(ctx, d_x);
    if (ch >= 0 && peer_ready(ctx) && spmv_single_launch(ctx)) {
      // x lives in the symmetric arena (cask_b200_dist_vector): ONE persistent launch over all slices, interior first.
      // Its last CTA stores the boundary rows straight into the neighbours' copies before it turns to its slices
      // (flow-controlled by acknowledgements); the producer warps acquire the neighbours' epoch flags only when they
      // reach the first halo-dependent slice
      HaloWait hw;
      CB_TRY(peer_acked_wait(ctx, ch, &hw));
      SpmvFusion f;
      f.pdl = true;
      return launch_spmv(ctx, d_x, d_y, 0, ctx->stream, &f, &hw);
    }
    CB_TRY(dist_exchange_begin(ctx, const_cast<double*>(d_x), ctx->stream));
    CB_TRY(launch_spmv(ctx, d_x, d_y, 1, ctx->stream, nullptr));
    CB_TRY(dist_exchange_wait(ctx, ctx->stream));
    return launch_spmv(ctx, d_x, d_y, 2, ctx->stream, nullptr);
  }
  return launch_spmv(ctx, d_x, d_y, 0, ctx->stream, nullptr);
}

static int stage_in(cask_b200_ctx* ctx, const double* x, int64_t m, int64_t n) {
  CB_TRY(grow(&ctx->d_x, &ctx->d_x_len, m));
  CB_TRY(grow(&ctx->d_y, &ctx->d_y_len, n));
  if (m) CB_CUDA(cudaMemcpyAsync(ctx->d_x, x, sizeof(double) * m, cudaMemcpyHostToDevice, ctx->stream));
  return CASK_B200_OK;
}

// Host-buffer SpMV as a three-stage pipeline over chunks of consecutive slices: the x columns a chunk needs
// are uploaded on one stream, its kernels run on the context's stream, its y rows are downloaded on a
// third — PCIe is full duplex, so for banded matrices the call costs about max(H2D, D2H) instead of their
// sum.  A chunk waits only for the upload that covers its largest referenced column (SliceDesc::col_hi);
// matrices whose slices reference far columns degrade gracefully to upload-all-then-overlap-download.
//
// PAGEABLE caller buffers (std::vector<double>, i.e. every cask::Vector; hostcopy.hpp): the same pipeline, but both
// directions go through rings of pinned 4 MB chunks that the library fills / drains itself with several host threads,
// instead of leaving the staging to the driver's single-threaded bounce buffer.  x: copy into a free ring slot, DMA
// from it, the slot is free again when that DMA's event has completed.  y: DMA into a ring slot, a drain thread waits
// for the event and copies the rows out while the issuing thread carries on with the next uploads.
static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

constexpr size_t kRingChunk = 4u << 20;  // bytes per pinned ring slot
constexpr int kRingSlots = 4;            // per direction

static int ensure_ring(cask_b200_ctx* ctx) {
  if (!ctx->h_ring) {
    CB_CUDA(cudaHostAlloc((void**)&ctx->h_ring, kRingChunk * kRingSlots * 2, cudaHostAllocDefault));
    for (int i = 0; i < 2 * kRingSlots; i++) {
      cudaEvent_t e;
      CB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->ring_events.push_back(e);
    }
  }
  if (!ctx->copy_pool) {
    const unsigned hw = std::thread::hardware_concurrency();
    const int t = getenv("CASK_B200_HOST_COPY_THREADS") ? atoi(getenv("CASK_B200_HOST_COPY_THREADS")) : (int)std::min(8u, std::max(2u, hw / 2));
    ctx->copy_pool = new HostCopyPool(std::max(1, t));
  }
  return CASK_B200_OK;
}

static int spmv_host_pipelined(cask_b200_ctx* ctx, const double* x, double* y, bool stage_x, bool stage_y) {
  const Plan& p = ctx->plan;
  // merge-path tiles cross slice boundaries: such plans run as one chunk (upload, all kernels, download)
  const int K = p.csr_merge ? 1 : std::max(1, std::min<int>(ctx->host_pipeline_chunks, p.nslices / 64));
  CB_TRY(grow(&ctx->d_x, &ctx->d_x_len, p.m));
  CB_TRY(grow(&ctx->d_y, &ctx->d_y_len, p.n));
  if (!ctx->h2d_stream) CB_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) CB_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  while ((int)ctx->pipe_events.size() < 2 * K + 1) {
    cudaEvent_t e;
    CB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_events.push_back(e);
  }
  if (stage_x || stage_y) CB_TRY(ensure_ring(ctx));
  cudaStream_t cs = ctx->stream;

  // y side of the staged path: the drain thread lives exactly as long as this call
  DrainQueue drain;
  struct Joiner {
    DrainQueue* q;
    std::thread t;
    ~Joiner() {
      if (t.joinable()) {
        q->close();
        t.join();
      }
    }
  } joiner{&drain, {}};
  if (stage_y) joiner.t = std::thread([&] { drain.run(ctx->copy_pool); });
  int64_t up_slot = 0, dn_slot = 0;  // ring positions (monotonic)
  unsigned char* up_ring = ctx->h_ring;
  unsigned char* dn_ring = ctx->h_ring ? ctx->h_ring + kRingChunk * kRingSlots : nullptr;

  // x[lo, hi) to the device on the upload stream
  auto upload = [&](int64_t lo, int64_t hi) -> int {
    if (!stage_x) {
      CB_CUDA(cudaMemcpyAsync(ctx->d_x + lo, x + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, ctx->h2d_stream));
      return CASK_B200_OK;
    }
    const int64_t per = (int64_t)(kRingChunk / sizeof(double));
    for (int64_t a = lo; a < hi; a += per) {
      const int64_t b = std::min(hi, a + per);
      const int slot = (int)(up_slot % kRingSlots);
      if (up_slot >= kRingSlots) CB_CUDA(cudaEventSynchronize(ctx->ring_events[slot]));  // the DMA that last read this slot
      unsigned char* buf = up_ring + kRingChunk * slot;
      ctx->copy_pool->copy(buf, x + a, sizeof(double) * (b - a));
      CB_CUDA(cudaMemcpyAsync(ctx->d_x + a, buf, sizeof(double) * (b - a), cudaMemcpyHostToDevice, ctx->h2d_stream));
      CB_CUDA(cudaEventRecord(ctx->ring_events[slot], ctx->h2d_stream));
      up_slot++;
    }
    return CASK_B200_OK;
  };
  // rows [lo, hi) of y to the host on the download stream (which already waits for their kernels)
  auto download = [&](int64_t lo, int64_t hi) -> int {
    if (!stage_y) {
      CB_CUDA(cudaMemcpyAsync(y + lo, ctx->d_y + lo, sizeof(double) * (hi - lo), cudaMemcpyDeviceToHost, ctx->d2h_stream));
      return CASK_B200_OK;
    }
    const int64_t per = (int64_t)(kRingChunk / sizeof(double));
    for (int64_t a = lo; a < hi; a += per) {
      const int64_t b = std::min(hi, a + per);
      const int slot = (int)(dn_slot % kRingSlots);
      drain.wait_in_flight(kRingSlots - 1);  // the drain thread has emptied this slot
      if (drain.error() != cudaSuccess) return fail(CASK_B200_ERR_CUDA, std::string("spmv download: ") + cudaGetErrorString(drain.error()));
      unsigned char* buf = dn_ring + kRingChunk * slot;
      CB_CUDA(cudaMemcpyAsync(buf, ctx->d_y + a, sizeof(double) * (b - a), cudaMemcpyDeviceToHost, ctx->d2h_stream));
      CB_CUDA(cudaEventRecord(ctx->ring_events[kRingSlots + slot], ctx->d2h_stream));
      drain.push(DrainItem{y + a, buf, sizeof(double) * (size_t)(b - a), ctx->ring_events[kRingSlots + slot]});
      dn_slot++;
    }
    return CASK_B200_OK;
  };

  // uploads must not start before earlier work queued on the compute stream (e.g. a previous async call) is done
  CB_CUDA(cudaEventRecord(ctx->pipe_events[2 * K], cs));
  CB_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, ctx->pipe_events[2 * K], 0));
  int64_t uploaded = 0;
  size_t ie = 0, ic = 0;  // cursors into the (ascending) slice lists
  // Chunk sizes grow geometrically from both ends (weights 1, 2, 4, 8, 16, 16, ... 16, 8, 4, 2, 1): the first upload and
  // the last download overlap nothing, so they are kept small, while the middle of the vector moves in few large
  // copies - every chunk costs about 17 us of copy-engine and launch latency (profiles/r2n_stripes_chunks_ilu.md).
  std::vector<int> bound((size_t)K + 1, 0);
  {
    std::vector<int64_t> w((size_t)K);
    int64_t total = 0;
    for (int c = 0; c < K; c++) {
      const int e = std::min(std::min(c, K - 1 - c), 4);
      w[c] = ctx->host_pipeline_ramp && !stage_x && !stage_y ? (int64_t)1 << e : 1;  // staged vectors move in 4 MB ring slots anyway
      total += w[c];
    }
    int64_t run = 0;
    for (int c = 0; c < K; c++) {
      run += w[c];
      bound[c + 1] = (int)((int64_t)p.nslices * run / total);
    }
  }
  for (int c = 0; c < K; c++) {
    const int s_lo = bound[c], s_hi = bound[c + 1];
    if (s_hi == s_lo) continue;
    int64_t need = 0;
    for (int s = s_lo; s < s_hi; s++) need = std::max<int64_t>(need, p.h_slices[s].col_hi);
    need = std::min<int64_t>(std::max<int64_t>(need, 0), p.m);
    if (c == K - 1 && !p.csr_merge) need = std::max(need, uploaded);  // nothing beyond `need` is ever read
    if (p.csr_merge) need = p.m;                                      // hub-clustered plans permute all of x
    if (need > uploaded) {
      CB_TRY(upload(uploaded, need));
      uploaded = need;
    }
    CB_CUDA(cudaEventRecord(ctx->pipe_events[2 * c], ctx->h2d_stream));
    CB_CUDA(cudaStreamWaitEvent(cs, ctx->pipe_events[2 * c], 0));
    const size_t e_lo = ie, c_lo = ic;
    while (ie < p.h_list_ell.size() && p.h_list_ell[ie] < s_hi) ie++;
    while (ic < p.h_list_csr.size() && p.h_list_csr[ic] < s_hi) ic++;
    CB_TRY(launch_spmv_range(ctx, ctx->d_x, ctx->d_y, (int)e_lo, (int)ie, (int)c_lo, (int)ic, cs, nullptr));
    CB_CUDA(cudaEventRecord(ctx->pipe_events[2 * c + 1], cs));
    CB_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->pipe_events[2 * c + 1], 0));
    const int64_t r_lo = p.h_slices[s_lo].row0, r_hi = (int64_t)p.h_slices[s_hi - 1].row0 + p.h_slices[s_hi - 1].nrows;
    CB_TRY(download(r_lo, r_hi));
  }
  if (stage_y) {
    drain.close();
    joiner.t.join();
    if (drain.error() != cudaSuccess) return fail(CASK_B200_ERR_CUDA, std::string("spmv download: ") + cudaGetErrorString(drain.error()));
  }
  CB_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
  CB_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
  CB_CUDA(cudaStreamSynchronize(cs));
  return CASK_B200_OK;
}

int cask_b200_spmv(cask_b200_ctx* ctx, const double* x, double* y) {
  CB_TRY(check_spmv(ctx));
  if (dist_active(ctx)) return fail(CASK_B200_ERR_UNSUPPORTED, "spmv (host buffers, whole vectors) is single-rank; sharded: cask_b200_spmv_shard");
  const Plan& p = ctx->plan;
  if ((!x && p.m) || (!y && p.n)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "spmv: null vector");
  // vectors of at least 4 MB in pageable memory are staged through the library's own pinned rings (hostcopy.hpp)
  const bool big = (p.m + p.n) * (int64_t)sizeof(double) >= (int64_t)kRingChunk && ctx->host_staging != 0;
  const bool stage_x = big && p.m && is_pageable(x), stage_y = big && p.n && is_pageable(y);
  // merge-path tiles cross slice boundaries, so the chunked pipeline (which launches slice ranges) is for plans without them
  if ((ctx->host_pipeline_chunks > 1 && p.nslices >= 128 && !p.csr_merge) || stage_x || stage_y) {
    try {  // the staged path starts host threads: std::system_error / bad_alloc must not cross the C ABI
      return spmv_host_pipelined(ctx, x, y, stage_x, stage_y);
    } catch (const std::exception& e) {
      cudaStreamSynchronize(ctx->stream);
      return fail(CASK_B200_ERR_RUNTIME, std::string("spmv (host buffers): ") + e.what());
    }
  }
  CB_TRY(stage_in(ctx, x, p.m, p.n));
  CB_TRY(launch_spmv(ctx, ctx->d_x, ctx->d_y, 0, ctx->stream, nullptr));
  if (p.n) CB_CUDA(cudaMemcpyAsync(y, ctx->d_y, sizeof(double) * p.n, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CASK_B200_OK;
}

int cask_b200_dist_vector(cask_b200_ctx* ctx, int32_t channel, double** d_vector) {
  if (!ctx || !d_vector || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_vector: preprocess_shard first");
  if (channel < 0 || channel >= kHaloChannels) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_vector: channel is 0 or 1");
  *d_vector = nullptr;
  if (!dist_active(ctx)) return CASK_B200_OK;
  CB_TRY(ensure_device(ctx));
  CB_TRY(peer_ensure_arena(ctx, ctx->plan.m));  // collective
  if (peer_ready(ctx) && spmv_single_launch(ctx)) *d_vector = peer_vector(ctx, channel);
  return CASK_B200_OK;
}

int cask_b200_spmv_shard(cask_b200_ctx* ctx, const double* x_slice, double* y_slice) {
  CB_TRY(check_spmv(ctx));
  const Plan& p = ctx->plan;
  if (!dist_active(ctx)) return cask_b200_spmv(ctx, x_slice, y_slice);
  if (p.n_global != p.m) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "spmv_shard: x is sharded like the rows, the system must be square");
  if ((!x_slice || !y_slice) && p.n) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "spmv: null vector");
  double* full = nullptr;
  CB_TRY(cask_b200_dist_vector(ctx, 0, &full));
  if (!full) {  // NCCL path: an own full-layout buffer
    CB_TRY(grow(&ctx->d_x, &ctx->d_x_len, p.m));
    full = ctx->d_x;
  }
  CB_TRY(grow(&ctx->d_y, &ctx->d_y_len, p.n));
  if (p.n) CB_CUDA(cudaMemcpyAsync(full + p.row0_global, x_slice, sizeof(double) * p.n, cudaMemcpyHostToDevice, ctx->stream));
  CB_TRY(cask_b200_spmv_device(ctx, full, ctx->d_y));
  if (p.n) CB_CUDA(cudaMemcpyAsync(y_slice, ctx->d_y, sizeof(double) * p.n, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  return peer_check_error(ctx);
}

int cask_b200_spmv_refformat(cask_b200_ctx* ctx, const double* x, double* y) {
  CB_TRY(check_spmv(ctx));
  const Plan& p = ctx->plan;
  if ((!x && p.m) || (!y && p.n)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "spmv: null vector");
  CB_TRY(stage_in(ctx, x, p.m, p.n));
  CB_TRY(spmv_refformat_device(ctx, ctx->d_x, ctx->d_y));
  if (p.n) CB_CUDA(cudaMemcpyAsync(y, ctx->d_y, sizeof(double) * p.n, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CASK_B200_OK;
}

// ---- solvers with host buffers -------------------------------------------------------------------
int cask_b200_cg(cask_b200_ctx* ctx, const double* rhs, double* x, int32_t maxiters, double tol, int32_t* iterations,
                 int32_t* converged, double* rs_final) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "cg: preprocess a matrix first");
  if (dist_active(ctx)) return fail(CASK_B200_ERR_UNSUPPORTED, "cg (host buffers) is single-rank; use cg_device when sharded");
  CB_TRY(ensure_device(ctx));
  const int64_t n = ctx->plan.n;
  if (!rhs || !x) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "cg: null vector");
  CB_TRY(grow(&ctx->d_x, &ctx->d_x_len, n));
  CB_TRY(grow(&ctx->d_y, &ctx->d_y_len, n));
  CB_CUDA(cudaMemcpyAsync(ctx->d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CB_CUDA(cudaMemcpyAsync(ctx->d_y, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CB_TRY(cask_b200_cg_device(ctx, ctx->d_y, ctx->d_x, maxiters, tol, iterations, converged, rs_final, nullptr));
  CB_CUDA(cudaMemcpyAsync(x, ctx->d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CASK_B200_OK;
}

int cask_b200_bicgstab(cask_b200_ctx* ctx, const double* b, double* x, int32_t* iters, double* tol_error) {
  if (!ctx || !ctx->have_design) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "bicgstab: preprocess a matrix first");
  if (dist_active(ctx)) return fail(CASK_B200_ERR_UNSUPPORTED, "bicgstab (host buffers) is single-rank; use bicgstab_device");
  CB_TRY(ensure_device(ctx));
  const int64_t n = ctx->plan.n;
  if (!b || !x) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "bicgstab: null vector");
  CB_TRY(grow(&ctx->d_x, &ctx->d_x_len, n));
  CB_TRY(grow(&ctx->d_y, &ctx->d_y_len, n));
  CB_CUDA(cudaMemcpyAsync(ctx->d_y, b, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  CB_TRY(cask_b200_bicgstab_device(ctx, ctx->d_y, ctx->d_x, iters, tol_error));
  CB_CUDA(cudaMemcpyAsync(x, ctx->d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  return CASK_B200_OK;
}

}  // extern "C"
