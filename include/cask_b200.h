/* cask_b200 — C ABI of the B200-native CASK SpMV / CG hot path.
 *
 * This header is the drop-in boundary: plain C, plain pointers and sizes, no C++ or torch types.
 * Every entry point names the reference interface (caskorg/cask @ 9e561d7, paths relative to the
 * reference root) that it stands in for.  All functions return CASK_B200_OK (0) or an error code;
 * the message is available from cask_b200_last_error() (thread-local).  No exception crosses this
 * boundary; the C++ mirror in cask_b200/host/ rethrows the reference's exception types.
 *
 * Ownership: host and device pointers passed in are borrowed for the duration of the call, except
 * the CSR arrays given to cask_b200_preprocess_device(), which must stay valid until the next
 * preprocess or cask_b200_destroy().  One context = one device + one stream; a context is not
 * thread-safe (the reference's GeneratedSpmvImplementation is not either,
 * src/runtime/GeneratedImplSupport.hpp:51-95).
 *
 * There is NO CPU fallback: every compute entry point fails with CASK_B200_ERR_NO_DEVICE when no
 * CUDA device is usable.
 */
#ifndef CASK_B200_H
#define CASK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CASK_B200_OK 0
#define CASK_B200_ERR_INVALID_ARGUMENT 1 /* std::invalid_argument in the reference (Spmv.cpp:195-206) */
#define CASK_B200_ERR_RUNTIME 2          /* std::runtime_error in the reference (Spmv.cpp:222-232)    */
#define CASK_B200_ERR_CUDA 3
#define CASK_B200_ERR_NO_DEVICE 4
#define CASK_B200_ERR_UNSUPPORTED 5
#define CASK_B200_ERR_NCCL 6

#define CASK_B200_ARCH_SIMPLE 0    /* cask::spmv::Spmv,              src/runtime/Spmv.hpp:49-202  */
#define CASK_B200_ARCH_SKIPEMPTY 1 /* cask::spmv::SkipEmptyRowsSpmv, src/runtime/Spmv.hpp:211-259 */

typedef struct cask_b200_ctx cask_b200_ctx;

/* The architecture parameters of GeneratedSpmvImplementation (GeneratedImplSupport.hpp:58).
 * Meaning on B200:
 *   num_pipes       row stripes, split exactly as Spmv::preprocess does (Spmv.cpp:334-364)
 *   cache_size      capacity, in doubles, of the on-chip x cache: column-block width of the exported
 *                   reference format AND the shared-memory x-cache budget of a row slice
 *   input_width     pair-stream padding of the exported reference format (Spmv.cpp:81-82)
 *   max_rows        capacity check of Spmv::spmv (Spmv.cpp:201-207); <= 0 disables it
 *   num_controllers must divide num_pipes (Spmv.cpp:226-232)
 */
typedef struct {
  int32_t num_pipes;
  int32_t cache_size;
  int32_t input_width;
  int32_t max_rows;
  int32_t num_controllers;
  int32_t dram_reduction_enabled;
  int32_t arch; /* CASK_B200_ARCH_* */
} cask_b200_design;

/* struct Partition scalars (src/runtime/Spmv.hpp:25-29) + the two array lengths. */
typedef struct {
  int32_t nBlocks, n, paddingCycles, totalCycles, vector_load_cycles, outSize;
  int32_t reductionCycles, emptyCycles;
  int32_t m_colptr_unpaddedLength, m_indptr_values_unpaddedLength;
  int64_t len_colptr, len_pairs;
} cask_b200_partition_info;

/* What the row-length histogram / band profile selected (the B200 analogue of the DSE record,
 * src/runtime/Dse.cpp:32-74). */
typedef struct {
  int64_t n, m, nnz;
  int32_t slice_rows;           /* rows per slice */
  int32_t num_slices;
  int32_t slices_staged_ell;    /* x window staged in shared memory by TMA bulk copy, thread-per-row */
  int32_t slices_gather_csr;    /* x gathered through L2, vector-per-row */
  int32_t csr_lanes_per_row;    /* sub-warp width chosen from the row-length histogram */
  int32_t max_row_length;
  int64_t ell_padded_entries;   /* stored entries incl. padding in staged slices */
  int64_t ell_nnz;              /* true nonzeros in staged slices */
  int64_t xcache_doubles_total; /* sum over staged slices of staged x doubles */
  int64_t device_bytes;         /* bytes of the compute format resident in HBM */
  int64_t row_length_histogram[8]; /* rows with length 0, 1-2, 3-4, 5-8, 9-16, 17-32, 33-64, >64 */
  int64_t csr_nnz, csr_rows;    /* nonzeros / rows of the gather slices */
  int32_t csr_items;            /* work items of the gather kernel (merge-path tiles or row groups) */
  int32_t csr_kernel;           /* 0 row-group items (spmv_csr_items_kernel), 1 merge-path tiles (spmv_csr_merge_kernel) */
  int32_t persist_ku;           /* persistent staged-ELL kernel: ELL columns per ring stage (0: kernel not in use) */
  int32_t persist_stages, persist_ctas_per_sm;
  int32_t value_dict;           /* format in use: 0 uncoded, 1 value codes, 2 pair codes */
  int32_t col_reorder;          /* gather kernel's column numbering: 0 original, 1 by descending reference count, 2 compact */
  int64_t cols_referenced;      /* columns referenced at least once (valid when col_reorder is not 0) */
  int32_t merge_items;          /* merge-path tiles: merge items per thread (tile = 256 x merge_items) */
  int32_t merge_ctas;           /* resident CTAs per SM the merge kernel instantiation is compiled for */
} cask_b200_plan_stats;

/* ---- context ---------------------------------------------------------------------------- */
int cask_b200_device_count(int* count);
int cask_b200_create(cask_b200_ctx** out, int device);
int cask_b200_destroy(cask_b200_ctx* ctx);
const char* cask_b200_last_error(void);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream; NULL = CUDA's legacy default
 * stream) instead of the context's own non-blocking stream; use_own_stream switches back.  A context
 * on its own stream is NOT ordered with work the caller queued on other streams: device-pointer
 * callers either share a stream this way or synchronize before and after. */
int cask_b200_set_stream(cask_b200_ctx* ctx, void* cuda_stream);
int cask_b200_use_own_stream(cask_b200_ctx* ctx);
int cask_b200_synchronize(cask_b200_ctx* ctx);
/* Tuning knobs of the kernel selector (tests and sweeps): "ell_min_fill" (default 0.75),
 * "force_kind" (-1 auto, 0 staged ELL wherever the x windows fit, 1 gather CSR everywhere),
 * "force_csr_vec" (0 auto, else 2/4/8/16/32 lanes per row), "peer_mode" (row-sharded solvers: 1 = halo pushes
 * and scalar all-reduces by the library's own kernels over IPC-mapped peer memory, 0 = NCCL send/recv and
 * all-reduce; every rank must use the same value), "value_dict" (0 default, 1 / 2 = coded staged ELL, see
 * cask_b200_plan_value_dict), "persist_ctas" (coded format: CTAs per SM of the persistent kernel, 0 auto),
 * "col_reorder" (gather path; -1 default = automatic: hub clustering for single-rank gather plans of at least 2^24
 * nonzeros whose column reference counts are skewed, 0 = off, 1 = columns renumbered by descending reference count so
 * that the hub columns of a power-law matrix share cache lines, 2 = referenced columns only, in column order; x is
 * permuted by a streaming kernel in front of every SpMV), "dist_sparse" (row-sharded gather plans, default 1: every rank receives only the x
 * entries its rows reference, packed by their owners; 0: every slice of x is broadcast to all ranks),
 * "host_staging" (default 1: pageable caller vectors of cask_b200_spmv go through the library's pinned rings and copy
 * threads; 0: left to the driver), "ilu_persistent" (default 1: all dependency levels of an ILU application run in one
 * cooperative kernel with grid barriers), "ilu_graph" (when not persistent; default 1: the per-level launches are
 * replayed as one CUDA graph).
 * Takes effect at the next preprocess. */
int cask_b200_set_option(cask_b200_ctx* ctx, const char* name, double value);

/* ---- preprocess: replaces Spmv::preprocess(const CsrMatrix&), src/runtime/Spmv.cpp:329-365 ---- */
/* Host CSR (0-based, row_ptr has n+1 entries: cask::CsrMatrix, SparseMatrix.hpp:272-319). */
int cask_b200_preprocess(cask_b200_ctx* ctx, const cask_b200_design* design, int64_t n, int64_t m,
                         int64_t nnz, const int32_t* row_ptr, const int32_t* col_ind,
                         const double* values);
/* Same with CSR arrays already resident in device memory (borrowed until the next preprocess). */
int cask_b200_preprocess_device(cask_b200_ctx* ctx, const cask_b200_design* design, int64_t n,
                                int64_t m, int64_t nnz, const int32_t* d_row_ptr,
                                const int32_t* d_col_ind, const double* d_values);
int cask_b200_plan_get_stats(cask_b200_ctx* ctx, cask_b200_plan_stats* out);
/* Time model of one SpMV on the plan described by *stats - the B200 stand-in for the reference's cycle model
 * (countComputeCycles, src/runtime/Spmv.cpp:25-40, and getEstimatedClockCycles, Spmv.hpp:103-109), used by the
 * architecture selector (host/include/Dse.hpp) and validated against measured kernel times by profiles/dse_validate.py.
 * No GPU needed.  *bytes = what the format streams from HBM: 10 B per stored staged-ELL entry, 12 B per gathered nonzero
 * + 4 B per gathered row, x and y once, and one 32-byte sector per gathered nonzero for the share of x that cannot stay in
 * L2.  *seconds = max(bytes / hbm_gbs, gather time), gather time = 39/32 LSU wavefronts per gathered nonzero at 0.64 of
 * one wavefront per clock per SM (measured: profiles/r2k_rmat_reorder.md).  hbm_gbs <= 0: 6456.5 (measured copy peak). */
int cask_b200_plan_estimate(const cask_b200_plan_stats* stats, double hbm_gbs, double l2_bytes, double sm_clock_hz,
                            int32_t sms, double* bytes, double* seconds);
/* Coded staged ELL (option "value_dict", off by default; set before preprocess).
 *   1  When every staged slice holds at most 256 distinct fp64 bit patterns - constant-coefficient stencils hold 2-6 -
 *      the values are stored as 8-bit codes into a per-slice table that travels with the slice's x windows: 3 bytes
 *      per stored nonzero instead of 10.
 *   2  When every staged slice holds at most 255 distinct (value, displacement) pairs - displacement = position of the
 *      entry in the slice's x cache minus its row inside the slice, constant along a diagonal, so a stencil needs one
 *      pair per stencil point - the 8-bit code names the pair and the 16-bit index stream disappears: 1 byte per
 *      stored nonzero.  Falls back to 1, then to the uncoded format, when a slice does not qualify.
 * Either way the same doubles are multiplied by the same x entries in the same order: y is bit-identical to the uncoded
 * kernel.  *active = the format the current plan runs (0 uncoded, 1, 2), *max_entries = table entries staged per slice,
 * *matrix_bytes_per_spmv = bytes of the matrix stream one SpMV reads from HBM in the format in use (x and y not
 * included).  Pointers may be NULL. */
int cask_b200_plan_value_dict(cask_b200_ctx* ctx, int32_t* active, int32_t* max_entries,
                              int64_t* matrix_bytes_per_spmv);

/* ---- parity hook: the reference's partition arrays, produced on the GPU --------------------- */
/* Partition p of Spmv::partitions (Spmv.hpp:51) as do_blocking builds it (Spmv.cpp:42-107).
 * colptr receives info.len_colptr int32, pairs receives info.len_pairs packed 12-byte
 * indptr_value records (Spmv.hpp:13-20).  Either pointer may be NULL. */
int cask_b200_partition_get_info(cask_b200_ctx* ctx, int32_t pipe, cask_b200_partition_info* out);
int cask_b200_partition_export(cask_b200_ctx* ctx, int32_t pipe, int32_t* colptr, void* pairs);

/* ---- y = A x: replaces cask::Vector Spmv::spmv(const cask::Vector&), Spmv.cpp:185-328 --------- */
/* Host buffers: x has m doubles, y has n doubles; includes H2D of x and D2H of y. */
int cask_b200_spmv(cask_b200_ctx* ctx, const double* x, double* y);
/* Device buffers (x 16-byte aligned); asynchronous on the context's stream. */
int cask_b200_spmv_device(cask_b200_ctx* ctx, const double* d_x, double* d_y);
/* y computed from the exported reference-format arrays exactly as the dataflow engine consumes
 * them: proves the emitted format is a complete description of A. Host buffers. */
int cask_b200_spmv_refformat(cask_b200_ctx* ctx, const double* x, double* y);

/* ---- solvers ------------------------------------------------------------------------------ */
/* pcg<double, IdentityPreconditioner>, src/runtime/SparseLinearSolvers.hpp:162-239, on the matrix
 * given to preprocess (full symmetric CSR).  x: initial guess in, solution out.  *iterations follows
 * the reference's convention (index of the last non-converged iteration; untouched if the loop
 * converges in its first two iterations), so pass the caller's initial value in.  tol is the
 * reference's 1e-5 (test: r.r <= tol*tol, absolute); maxiters its 2000. */
int cask_b200_cg(cask_b200_ctx* ctx, const double* rhs, double* x, int32_t maxiters, double tol,
                 int32_t* iterations, int32_t* converged, double* rs_final);
int cask_b200_cg_device(cask_b200_ctx* ctx, const double* d_rhs, double* d_x, int32_t maxiters,
                        double tol, int32_t* iterations, int32_t* converged, double* rs_final,
                        int32_t* loop_trips);
/* Eigen::BiCGSTAB<SparseMatrix<double>> with its default Jacobi preconditioner, as called by
 * solveBICG / EigenSolver::solve, src/runtime/SparseLinearSolvers.cpp:18-26,62-67.
 * In: *iters = max iterations (<=0: 2*n), *tol_error = tolerance (<=0: DBL_EPSILON).
 * Out: iterations performed and ||r||/||b||.  x is overwritten (initial guess 0, as Eigen's solve()). */
int cask_b200_bicgstab(cask_b200_ctx* ctx, const double* b, double* x, int32_t* iters,
                       double* tol_error);
int cask_b200_bicgstab_device(cask_b200_ctx* ctx, const double* d_b, double* d_x, int32_t* iters,
                              double* tol_error);

/* ---- one process per GPU: row-sharded execution over NCCL ----------------------------------- */
/* Row stripes follow Spmv::preprocess (rows_per = n / world, remainder to the last rank). */
int cask_b200_nccl_unique_id(void* out_128_bytes);
int cask_b200_dist_init(cask_b200_ctx* ctx, int32_t rank, int32_t world, const void* unique_id_128_bytes);
/* Host-side shard arithmetic (no GPU needed): rows [row0, row0+nrows) owned by `rank`. */
int cask_b200_shard_rows(int64_t n, int32_t world, int32_t rank, int64_t* row0, int64_t* nrows);
/* Host-side halo plan (no GPU needed; the same routine the GPU path runs after it has built a stripe's x windows):
 * given the contiguous x windows [run_col0[i], run_col0[i] + run_len[i]) that `rank` stages, returns the column
 * ranges it must receive per peer, ranges merged and split by owner under the striping above.  *out_count
 * receives the number of (peer, col0, len) triples; the call fails if it exceeds `capacity`. */
int cask_b200_halo_plan_host(int64_t n_global, int32_t world, int32_t rank, int64_t nruns, const int64_t* run_col0,
                             const int64_t* run_len, int64_t capacity, int32_t* out_peer, int64_t* out_col0,
                             int64_t* out_len, int64_t* out_count);

/* Host-side arithmetic of the SPARSE exchange of a row-sharded gather plan (no GPU needed; the same routines the GPU path
 * runs, exercised by world-size 2/3 gloo tests on CPU).  A rank renumbers the columns its rows reference compactly, in
 * column order: need[0 .. count) ascending.  bounds[q] = first row (= first column) rank q owns, bounds[world] = n.
 *   segments:  seg[q] = first position of need[] that belongs to owner q (seg[world] = count): segment q of the rank's
 *              compact x holds the entries it receives from rank q (its own segment is filled locally).
 *   send plan: all_seg = every rank's seg array, rank-major (world x (world + 1)).  For rank `rank`:
 *              send_off[q] .. send_off[q + 1] = positions of its send list that go to rank q (nothing to itself), and
 *              dst_off[q] = where those entries start inside rank q's compact x. */
int cask_b200_sparse_segments_host(const int64_t* bounds, int32_t world, const int32_t* need, int64_t count, int64_t* seg);
int cask_b200_sparse_send_plan_host(const int64_t* all_seg, int32_t world, int32_t rank, int64_t* send_off, int64_t* dst_off);
/* Local stripe of the global n x m matrix: d_row_ptr has nrows+1 entries rebased to 0 (exactly
 * CsrMatrix::sliceRows, SparseMatrix.hpp:426-443), column indices stay global.  Collective.  The ranks' row ranges must be
 * contiguous, in rank order and cover all rows; cask_b200_shard_rows gives the reference's partition (equal row counts),
 * but any other is accepted - a power-law matrix is better cut into stripes of equal NONZERO count (bench.py does that
 * for R-MAT: under the reference rule rank 0 of 8 owns 44 % of the nonzeros).  Vectors are sharded like the rows. */
int cask_b200_preprocess_shard_device(cask_b200_ctx* ctx, const cask_b200_design* design,
                                      int64_t n_global, int64_t m, int64_t row0, int64_t nrows,
                                      int64_t nnz_local, const int32_t* d_row_ptr,
                                      const int32_t* d_col_ind, const double* d_values);
/* Host-side halo plan of the last preprocess_shard: for each peer, the number of x doubles this
 * rank receives from it per SpMV. counts has `world` entries. */
int cask_b200_dist_halo_counts(cask_b200_ctx* ctx, int64_t* recv_counts);
/* 1 if the row-sharded solvers of this context run on the peer-memory path (halo entries stored into the
 * neighbours' vectors by the producing kernel, scalar all-reduces by a kernel over IPC-mapped control blocks),
 * 0 if they use NCCL send/recv + all-reduce (option peer_mode = 0, irregular halo, or no IPC on this machine).
 * Meaningful after the first solver call that followed a preprocess_shard. */
int cask_b200_dist_peer_active(cask_b200_ctx* ctx, int32_t* active);
/* COLLECTIVE (every rank, after preprocess_shard).  Full-layout vector `channel` (0 or 1; m doubles, this rank's slice at
 * [row0, row0 + nrows)) of the library's symmetric arena, the allocation every peer has mapped over NVLink.  A sharded
 * caller that keeps x THERE gets the one-launch SpMV: in cask_b200_spmv_device(ctx, that pointer, y) the persistent
 * kernel's last CTA stores this rank's boundary rows straight into the neighbours' copies (flow-controlled by
 * acknowledgements) while the others stream interior slices, and the producer warps acquire the neighbours' epoch flags
 * when they reach the first halo-dependent slice - no NCCL call, no second launch.  *d_vector = NULL (status OK) when the peer-memory path is not available for this plan (gather slices,
 * irregular halo, no IPC, peer_mode 0): use an own buffer then, spmv_device exchanges it over NCCL.  The solvers use the
 * same two vectors as scratch: the contents do not survive a cg / bicgstab call. */
int cask_b200_dist_vector(cask_b200_ctx* ctx, int32_t channel, double** d_vector);
/* Sharded counterpart of cask_b200_spmv (Spmv::spmv(const Vector&), Spmv.cpp:185-328, where the reference writes x to
 * every pipe itself, :165-170,234-258): host buffers, x_slice = the nrows entries of x this rank owns (needs a square
 * system), y_slice = its nrows results; the x exchange with the neighbours happens inside.  Collective. */
int cask_b200_spmv_shard(cask_b200_ctx* ctx, const double* x_slice, double* y_slice);

/* ---- Matrix Market ingest: file -> CSR on the device ------------------------- */
/* Replaces io::readHeader / readDokMatrix / readMatrix / readSymMatrix / readVector (src/runtime/IO.hpp:60-176) and
 * the DokMatrix -> CsrMatrix conversion behind them (SparseMatrix.hpp:156-189, 289-305).  The text is tokenised on
 * the host by all cores (mm_* below: no GPU needed); the dictionary-of-keys build - last value wins for a repeated
 * key, explicitSymmetric mirroring with its "Matrix is not symmetric" check, rows in ascending column order - is a
 * radix sort and four data-parallel passes on the GPU (ingest_*), and the result can be handed to preprocess
 * without ever visiting the host. */
typedef struct {
  char type[16], format[16], data_type[16], symmetry[16]; /* the header words, struct MmInfo IO.hpp:39-58 */
  int64_t n, m, entries;                                  /* the size line: N M L (coordinate) or N M (array) */
} cask_b200_mm_info;
/* io::readHeader (IO.hpp:60-71) + the size line.  Same acceptance rule and messages as the reference. */
int cask_b200_mm_read_info(const char* path, cask_b200_mm_info* info);
/* The entries of a coordinate file in file order, indices 1-based as stored.  *count receives L. */
int cask_b200_mm_read_coo(const char* path, int64_t capacity, int32_t* rows, int32_t* cols, double* vals, int64_t* count);
/* io::readVector (IO.hpp:73-115), including its quirk: coordinate vectors are not rebased (v[a] = val). */
int cask_b200_mm_read_vector(const char* path, int64_t capacity, double* out, int64_t* n);

typedef struct cask_b200_csr cask_b200_csr; /* a CSR matrix resident on the device of the context that built it */
#define CASK_B200_INGEST_ONE_BASED 1  /* indices are 1-based (Matrix Market) */
#define CASK_B200_INGEST_SYMMETRIC 2  /* DokMatrix::explicitSymmetric: every (i, j, v) also gives (j, i, v) */
#define CASK_B200_INGEST_DROP_UPPER 4 /* ignore entries above the diagonal (what mkl_dcsrsymv('l') reads); with
                                         SYMMETRIC this expands a stored lower triangle to the full matrix */
/* COO (host or device arrays, `count` entries in file order) -> CSR.  Fails with CASK_B200_ERR_INVALID_ARGUMENT and
 * the reference's message "Matrix is not symmetric" when SYMMETRIC meets a stored transpose pair with different
 * values, or when an index lies outside the n x m matrix. */
int cask_b200_ingest_coo(cask_b200_ctx* ctx, int64_t n, int64_t m, int64_t count, const int32_t* rows,
                         const int32_t* cols, const double* vals, int32_t flags, cask_b200_csr** out);
int cask_b200_ingest_coo_device(cask_b200_ctx* ctx, int64_t n, int64_t m, int64_t count, const int32_t* d_rows,
                                const int32_t* d_cols, const double* d_vals, int32_t flags, cask_b200_csr** out);
/* mode 0: io::readMatrix (IO.hpp:151-163; symmetric files are expanded).  mode 1: io::readSymMatrix (IO.hpp:165-176;
 * the stored triangle is kept as it is; fails with the reference's message if the file is not symmetric). */
int cask_b200_read_matrix(cask_b200_ctx* ctx, const char* path, int32_t mode, cask_b200_csr** out);
/* nnz = entries stored; nnzs_field = what the reference's CsrMatrix::nnzs / row_ptr[n] holds (it counts every
 * DokMatrix::set call, so it exceeds nnz when a file repeats a key or stores both (i, j) and (j, i)). */
int cask_b200_csr_get_info(const cask_b200_csr* csr, int64_t* n, int64_t* m, int64_t* nnz, int64_t* nnzs_field);
int cask_b200_csr_export(cask_b200_ctx* ctx, const cask_b200_csr* csr, int32_t* row_ptr, int32_t* col_ind, double* values);
int cask_b200_csr_device_arrays(const cask_b200_csr* csr, const int32_t** d_row_ptr, const int32_t** d_col_ind,
                                const double** d_values);
int cask_b200_csr_free(cask_b200_csr* csr);
/* Spmv::preprocess on an ingested matrix; the CSR stays owned by `csr`, which must outlive the next preprocess. */
int cask_b200_preprocess_csr(cask_b200_ctx* ctx, const cask_b200_design* design, const cask_b200_csr* csr);

/* ---- preconditioned CG --------------------------------------------------------- */
/* pcg<double, Precon>, src/runtime/SparseLinearSolvers.hpp:162-239, with the preconditioner left in.  Single rank.
 *   IDENTITY  IdentityPreconditioner (:64-73); cask_b200_cg is the tuned loop for this case
 *   ILU       ILUPreconditioner (:77-156): ILU(0) in IKJ order on the pattern of the matrix, applied as the reference
 *             applies it - mkl_dcsrtrsv with diag = 'N' for BOTH factors (MklLayer.hpp:66-84), i.e. the lower solve
 *             divides by U's diagonal.  M = (D + L)(D + U) is then not symmetric and the reference's PCG stalls on
 *             SPD stencils; reproduced because it is what the reference computes (test/LinearSolvers.cpp:54-77)
 *   JACOBI    z = r / a_ii (1 where a_ii is absent or 0).  Not in the reference
 *   ILU_UNIT  the same ILU(0) factors with a unit lower solve, M = (I + L)(D + U): the textbook preconditioner.
 *             Not in the reference
 * ILU needs rows in strictly ascending column order (CsrMatrix(DokMatrix) / the ingest path give that); the
 * factorisation and both triangular solves are level-scheduled on the GPU, bit-identical to the sequential loops.
 * Arguments as cask_b200_cg; x holds the iterate the loop stopped at, converged or not.  A zero or missing pivot
 * met by an ILU solve is reported as CASK_B200_ERR_RUNTIME after the loop. */
#define CASK_B200_PRECON_IDENTITY 0
#define CASK_B200_PRECON_ILU 1
#define CASK_B200_PRECON_JACOBI 2
#define CASK_B200_PRECON_ILU_UNIT 3
int cask_b200_pcg(cask_b200_ctx* ctx, const double* rhs, double* x, int32_t maxiters, double tol, int32_t precon,
                  int32_t* iterations, int32_t* converged, double* rs_final);
int cask_b200_pcg_device(cask_b200_ctx* ctx, const double* d_rhs, double* d_x, int32_t maxiters, double tol,
                         int32_t precon, int32_t* iterations, int32_t* converged, double* rs_final);
/* Build the preconditioner from `csr` instead of the matrix given to preprocess (NULL: back to that matrix).  The
 * reference's pcg constructs Precon{a} from the array it was handed - the stored LOWER TRIANGLE in its tests - while
 * the product uses the implied symmetric matrix; this call reproduces that pairing.  Same dimensions as A; `csr` is
 * borrowed until the next preprocess (which clears the override) or cask_b200_destroy. */
int cask_b200_precond_set_matrix(cask_b200_ctx* ctx, const cask_b200_csr* csr);
/* Parity hooks.  ilu_factor: ILUPreconditioner's `pc` (:89-140) in the pattern of the matrix, nnz doubles (pc may be
 * NULL), and the number of dependency levels of the lower / upper solve.  ilu_apply: ILUPreconditioner::apply
 * (:142-150) on host vectors; unit_lower selects the ILU_UNIT variant. */
int cask_b200_ilu_factor(cask_b200_ctx* ctx, double* pc, int32_t* levels_lower, int32_t* levels_upper);
int cask_b200_ilu_apply(cask_b200_ctx* ctx, int32_t unit_lower, const double* x, double* z, int32_t* zero_pivot);

/* ---- synthetic matrices of BASELINE.json, generated on the device ----------------------------- */
#define CASK_B200_SYNTH_POISSON2D 0   /* 5-point, N x N grid  */
#define CASK_B200_SYNTH_POISSON3D27 1 /* 27-point, N^3 grid   */
#define CASK_B200_SYNTH_CONVDIFF3D7 2 /* 7-point upwind convection-diffusion, N^3 grid */
int cask_b200_synth_rows(int32_t kind, int32_t N, int64_t* n);
int cask_b200_synth_nnz(int32_t kind, int32_t N, int64_t row0, int64_t nrows, int64_t* nnz);
int cask_b200_synth_device(int32_t kind, int32_t N, int64_t row0, int64_t nrows, int32_t* d_row_ptr,
                           int32_t* d_col_ind, double* d_values, void* cuda_stream);

/* ---- the reference's own device boundary, for the UNMODIFIED reference Spmv::spmv ------------- */
/* Flat byte-addressed memory per controller + a blocking run, exactly what SLiC generates for a design
 * (src/spmv/src/SpmvDeviceInterface.h:21-73; mock versions src/runtime/GeneratedImplSupport.hpp:31-49).
 * The plugin cask_b200/host/libSpmv_b200.so wraps these three into the `void` callbacks that
 * GeneratedSpmvImplementation stores (GeneratedImplSupport.hpp:59-61); num_pipes, num_controllers and
 * input_width are the design's build parameters, which a SLiC design has baked in.  Arrays of
 * legacy_run have num_pipes entries, those of write/read num_controllers entries with exactly one
 * non-zero size (msinglearray, Spmv.cpp:109-114). */
int cask_b200_legacy_write(int32_t num_controllers, int64_t size_bytes_cpu, const int64_t* size_bytes_memory_ctl,
                           const int64_t* start_bytes_memory_ctl, const uint8_t* instream_fromcpu);
int cask_b200_legacy_read(int32_t num_controllers, int64_t size_bytes_cpu, const int64_t* size_bytes_memory_ctl,
                          const int64_t* start_bytes_memory_ctl, uint8_t* outstream_tocpu);
int cask_b200_legacy_run(int32_t num_pipes, int32_t num_controllers, int32_t input_width, int64_t nIterations,
                         int64_t nPartitions, int64_t vectorLoadCycles, const int64_t* colPtrStartAddresses,
                         const int32_t* colptrSizes, const int64_t* indptrValuesAddresses,
                         const int32_t* indptrValuesSizes, const int32_t* nrows, const int64_t* outStartAddresses,
                         const int32_t* reductionCycles, const int32_t* totalCycles, const int64_t* vStartAddresses);
int cask_b200_legacy_reset(void);
int cask_b200_legacy_launch_count(int64_t* count);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
int cask_b200_launch_count(cask_b200_ctx* ctx, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* CASK_B200_H */
