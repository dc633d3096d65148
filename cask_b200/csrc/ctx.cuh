// Internal declarations shared by the translation units of libcask_b200.so.
// Nothing here is part of the ABI; the ABI is include/cask_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/cask_b200.h"
#include "hostcopy.hpp"

namespace caskb200 {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define CB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::caskb200::fail(CASK_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define CB_TRY(expr)                 \
  do {                               \
    int _rc = (expr);                \
    if (_rc != CASK_B200_OK) return _rc; \
  } while (0)

// ---- compute format ---------------------------------------------------
constexpr int kGranuleShift = 4;               // x is staged in granules of 16 doubles (128 B)
constexpr int kGranule = 1 << kGranuleShift;
constexpr int kZeroSlots = 2;                  // xs[0..1] = 0.0: target of padding entries, keeps runs 16-B aligned
constexpr int kEllThreads = 256;               // threads per CTA of the staged kernel
constexpr int kEllRowsPerThread = 4;           // rows per thread -> 128-bit value loads
constexpr int kSliceRows = kEllThreads * kEllRowsPerThread;  // 1024 rows per slice
constexpr int kBitmapWords = 8192;             // granule bitmap of a slice: 8192*32 granules = 4M columns of span
constexpr int kMaxRuns = 1024;                 // contiguous x runs staged per slice
constexpr int kMaxCacheDoubles = 24576;        // 192 KB of shared memory for the x cache at most

enum SliceKind : int32_t { kSliceStagedEll = 0, kSliceGatherCsr = 1 };

struct Run {           // one contiguous window of x staged by a single bulk copy
  int32_t col0;        // first column (multiple of kGranule)
  int32_t len;         // doubles (clipped at m)
  int32_t local_base;  // position in the slice's shared-memory x cache
  int32_t pad_;
};

constexpr int kInlineRuns = 6;  // runs carried inside the descriptor: one load gives a producer everything

struct SliceDesc {     // one row slice (<= kSliceRows consecutive rows of one stripe)
  int32_t row0;        // first row, local to this rank's stripe
  int32_t nrows;       // valid rows
  int32_t width;       // ELL width (max row length in the slice)
  int32_t kind;        // SliceKind
  int64_t val_off;     // entry offset into ell_vals / ell_idx
  int32_t run_off;     // offset into runs[]
  int32_t nruns;
  int32_t xcache_len;  // doubles staged in shared memory, including the zero slots
  int32_t nnz;
  int32_t remote;      // 1 if some staged column lies outside this rank's own x slice (dist only)
  int32_t col_hi;      // one past the largest column the slice references (host-side pipelining of x uploads)
  Run inl[kInlineRuns]; // copy of runs[run_off ..] when nruns <= kInlineRuns (written by plan_fill_kernel)
};

// Unit of work of the gather-CSR path: a run of consecutive rows holding ~8K nonzeros, or one segment of a
// single long row (R-MAT hubs) so that no CTA ever owns more than kCsrSegment nonzeros.
struct CsrItem {
  int32_t row0, nrows;  // local rows
  int32_t k_lo, k_hi;   // nonzero range (used by single-row items)
  int32_t scratch;      // >= 0: partial sum of a split row goes to scratch[scratch]; -1: write y directly
  int32_t vec;          // lanes per row (2..32) for multi-row items; 0: the whole CTA works on one row
  int32_t pad_[2];
};
struct SplitRow { int32_t row, first, nseg, pad_; };
constexpr int kCsrItemNnz = 8192;    // target nonzeros per multi-row item
constexpr int kCsrLongRow = 2048;    // rows at least this long get a CTA (or several) of their own
constexpr int kCsrSegment = 16384;   // nonzeros per segment of a split row

// Merge-path tiles of the gather path (spmv.cu: spmv_csr_merge_kernel; plan.cu: build_merge_tiles): the rows of a run
// of consecutive gather slices and their nonzeros are ONE merged sequence (a row end after the row's last nonzero),
// cut every kMergeTile items, so a tile is bounded in rows AND nonzeros whatever the row-length distribution.
struct MergeTile { int32_t r0, k0, r1, k1; };  // merge coordinates (local row, nonzero) of the tile's start and end
constexpr int kMergeThreads = 256;
constexpr int kMergeItemsDefault = 7;  // merge items per thread (5, 7, 11 or 17: odd, so that neighbouring threads start
                                       // in different banks); the tile's products and row ends live in shared memory
                                       // (12 bytes per item), and what shared memory takes L1 loses - the x gathers in
                                       // flight are bounded by the L1 that is left (profiles/r2b_spmv_merge_rmat_ncu.md)
constexpr int64_t kMergeAutoNnz = 1 << 20;  // gather slices holding at least this many nonzeros run the merge kernel
// automatic hub clustering of a single-rank gather plan (plan.cu: build_col_reorder): from this many gathered nonzeros,
// when x is larger than this (and the column reference counts are skewed)
constexpr int64_t kReorderAutoNnz = 1 << 24;
constexpr int64_t kReorderAutoXBytes = 64 << 20;

struct RefPartition {  // reference-format partition resident on the device
  cask_b200_partition_info info{};
  int32_t* d_colptr = nullptr;
  uint8_t* d_pairs = nullptr;  // packed 12-byte records
  int64_t row0 = 0;
  bool values_zeroed = false;
};

struct Plan {
  int64_t n = 0, m = 0, nnz = 0;       // local rows, global columns, local nnz
  int64_t n_global = 0, row0_global = 0;
  const int32_t* d_row_ptr = nullptr;  // CSR (owned or borrowed)
  const int32_t* d_col = nullptr;
  const double* d_val = nullptr;
  bool owns_csr = false;

  int32_t slice_rows = kSliceRows;
  int32_t nslices = 0;
  SliceDesc* d_slices = nullptr;
  Run* d_runs = nullptr;
  double* d_ell_vals = nullptr;
  uint16_t* d_ell_idx = nullptr;
  // coded staged ELL (option value_dict; plan.cu: build_value_dict, valuedict_logic.inl): 8-bit codes into a per-slice
  // table of the slice's distinct values - 3 bytes per stored nonzero instead of 10, same doubles multiplied
  // pair-coded staged ELL (value_dict = 2; valuedict_logic.inl: BuildPairs): the code names a (value, x-cache
  // displacement) pair, position = displacement + row inside the slice - 1 byte per stored nonzero, no index stream
  int32_t coded = 0;                  // 0 uncoded, 1 value codes (+ 16-bit indices), 2 pair codes
  uint8_t* d_ell_codes = nullptr;     // same indexing as d_ell_vals
  double* d_ell_dict = nullptr;       // valuedict::kStride doubles per slice id
  void* d_ell_pairs = nullptr;        // pair codes: valuedict::kPairStride 16-byte {value, displacement} records per slice id
  int32_t dict_len = 0;               // table entries staged per slice: the largest table, rounded up to a multiple of
                                      // 2 (value codes) or 8 (pair codes) so that every bulk copy moves whole 16 bytes
  int32_t* d_list_ell = nullptr;      // slice ids, staged ELL, interior first then halo-dependent
  int32_t* d_list_csr = nullptr;
  int32_t n_ell = 0, n_ell_interior = 0;
  int32_t n_csr = 0, n_csr_interior = 0;
  int32_t csr_vec = 4;
  bool csr_stream = false;        // gather-CSR items run the CSR-stream variant (products staged in shared memory)
  int32_t csr_item_nnz = kCsrItemNnz;
  CsrItem* d_csr_items = nullptr;
  SplitRow* d_split_rows = nullptr;
  double* d_csr_scratch = nullptr;
  std::vector<int32_t> h_item_begin, h_split_begin;  // per position of the CSR slice list (+1 sentinel)
  // merge-path variant: h_item_begin[pos] = first tile of the run that starts at list position pos, -1 inside a run
  bool csr_merge = false;
  MergeTile* d_merge_tiles = nullptr;
  double* d_merge_carry = nullptr;    // per tile: partial sum of the row the tile ends in (0 if it ends on a row boundary)
  int32_t n_merge_tiles = 0;
  int32_t merge_items = kMergeItemsDefault;   // merge items per thread the tiles were cut for
  int32_t merge_ctas = 0;                     // resident CTAs per SM the merge kernel instantiation is compiled for (0: default of the item count)
  // hub clustering (plan.cu: build_col_reorder): columns renumbered by descending reference count for the gather kernel
  int32_t* d_col_perm = nullptr;      // relabelled copy of d_col (all local nonzeros); nullptr: not reordered
  int32_t* d_perm = nullptr;          // [m] new column -> old column
  double* d_xperm = nullptr;          // [cols_used] x in the new numbering, rewritten in front of every SpMV
  int32_t cols_used = 0;              // columns referenced at least once = the first cols_used new columns
  bool xperm_external = false;        // d_xperm is filled by the sparse exchange of a sharded plan (dist.cu), not by permute_x_kernel
  bool hub_ordered = false;           // the renumbering puts the most referenced columns first: gathers go through L1
  bool xperm_owned = true;            // false: d_xperm lives in the symmetric arena (peers store into it over NVLink)
  int32_t max_xcache = 0;
  // persistent staged-ELL kernel configuration (spmv.cu: configure_persistent)
  int32_t persist_ku = 0;       // ELL columns per ring stage (0: persistent kernel not usable)
  int32_t persist_stages = 0;
  int32_t persist_ctas_per_sm = 0;
  int32_t persist_xbuf = 0;     // doubles per x buffer
  size_t persist_smem = 0;
  cask_b200_plan_stats stats{};
  std::vector<SliceDesc> h_slices;
  std::vector<int32_t> h_list_ell, h_list_csr;  // host copies of the slice lists (ascending slice id unless sharded)
};

struct DistState;     // dist.cu
struct PrecondState;  // precond.cu

// ---- peer-memory layer (dist.cu) --------------------------------------------------
// Every rank owns one "symmetric arena" (one cudaMalloc, exported with cudaIpcGetMemHandle and mapped by all
// peers over NVLink): a control block followed by the full-layout vectors the solvers feed to the SpMV.
// Kernels store halo entries and reduction scalars straight into the peers' arenas.
constexpr int kMaxPeers = 8;       // one NVSwitch box
constexpr int kMaxPush = 8;        // contiguous row ranges a rank pushes per vector (stencils: 2)
constexpr int kHaloChannels = 2;   // full-layout vectors living in the arena (CG: p; BiCGStab: y, z)

struct PeerCtrl {
  // written by peers (slot = sequence number & 3; a rank can be at most one collective ahead of another)
  unsigned long long ar_flag[4][kMaxPeers];
  double ar_val[4][kMaxPeers][2];
  unsigned long long halo_flag[kHaloChannels][kMaxPeers];  // [channel][source rank] = epoch whose halo has landed
  // low-latency slots of the in-kernel all-reduce: [slot][source rank][quantity][half] = {32 data bits, 32-bit
  // sequence number} in ONE 8-byte store - data and flag arrive together, no fence, one NVLink traversal
  unsigned long long ar_ll[4][kMaxPeers][2][2];
  // flow control of the plain sharded SpMV (peer.cuh: acked_push): [channel][rank q] = number of SpMVs on
  // that channel rank q has FINISHED, i.e. the halo rows this rank pushed for them may be overwritten
  unsigned long long halo_ack[kHaloChannels][kMaxPeers];
  // [channel][source rank] = number of plain sharded SpMVs whose boundary rows have landed here (host-counted epochs:
  // the solvers' device-counted halo_flag / push_seq epochs are a separate space and stay untouched)
  unsigned long long halo_flag_acked[kHaloChannels][kMaxPeers];
  // local
  unsigned long long ar_seq;                    // all-reduces performed
  unsigned long long push_seq[kHaloChannels];   // halo epochs pushed per channel (== the epoch the next SpMV expects)
  unsigned int push_ticket[kHaloChannels];      // last-CTA detection of the pushing kernels
  int error;                                    // a wait on a peer flag timed out
};

struct PushDesc {       // by-value kernel argument of every kernel that produces a full-layout vector
  int32_t nsend = 0;    // ranges to push; the peer path is in use iff ctrl != nullptr
  int32_t channel = 0;
  int32_t me = 0;
  int32_t pad_ = 0;
  int64_t lo[kMaxPush], hi[kMaxPush];  // local row range [lo, hi) that a peer stages
  double* dst[kMaxPush];               // peer's copy of the vector, rebased: dst[i] is its entry for local row i
  PeerCtrl* peer_ctrl[kMaxPush];       // that peer's control block
  PeerCtrl* ctrl = nullptr;            // this rank's control block
};

struct SparsePushDesc {  // by-value kernel argument of sparse_push_kernel (dist.cu): the sparse exchange over peer memory
  PeerCtrl* ctrl = nullptr;              // this rank's control block
  PeerCtrl* peers[kMaxPeers] = {nullptr};
  double* dst[kMaxPeers] = {nullptr};    // rank q's compact x, rebased to the segment this rank fills there
  int64_t send_off[kMaxPeers + 1] = {0}; // send list positions by destination rank
  uint32_t send_mask = 0, recv_mask = 0; // ranks this rank stores entries for / receives entries from
  int32_t me = 0, world = 1;
  unsigned long long done = 0;           // sparse exchanges this rank has completed (host-counted, equal on all ranks)
};

struct AckDesc {        // device-resident (one per channel): the plain sharded SpMV's in-kernel boundary-row push
  PeerCtrl* peers[kMaxPeers] = {nullptr};  // control blocks of all ranks
  uint32_t send_mask = 0;          // ranks that stage rows of this rank's slice
  int32_t me = 0, world = 1;
  int32_t pad_ = 0;
  PushDesc push;                   // what goes where
};

struct ReduceDesc {     // by-value kernel argument: the grid's last CTA finishes a dot product (and its all-reduce)
  double* partials = nullptr;      // [nq][stride] per-CTA partial sums; nullptr: no reduction in this launch
  int32_t stride = 0, nq = 0;
  unsigned int* ticket = nullptr;  // zero between launches
  double* out = nullptr;           // out[0 .. nq) receives the sums
  PeerCtrl* ctrl = nullptr;        // non-null: one-shot all-reduce over the peers' control blocks before writing out
  PeerCtrl* peers[kMaxPeers] = {nullptr};
  int32_t me = 0, world = 1;
  unsigned int* gen = nullptr;     // non-null: grid barrier - no CTA returns before out[] is written (all CTAs resident)
  int32_t publish_only = 0;        // peer path: store the local sums into the peers' slots and return; the CONSUMER
  int32_t pad_ = 0;                // kernel gathers the W contributions (peer_gather_sum) - hides the NVLink round trip
};

struct GatherDesc {     // by-value kernel argument of the consumer of a publish-only reduction
  PeerCtrl* ctrl = nullptr;        // nullptr: the value is in the scalar slot already
  int32_t world = 1;
  int32_t pad_ = 0;
};

struct HaloUpdate {     // by-value kernel argument: halo rows of a full-layout vector that this rank recomputes itself
  int32_t nrecv = 0;               // from halo entries of ANOTHER vector that the peers pushed (CG: p_halo = r_halo + beta p_halo)
  int32_t channel = 0;             // channel of the pushed vector
  uint32_t peer_mask = 0;
  int32_t pad_ = 0;
  int64_t lo[kMaxPush], hi[kMaxPush];  // ranges as offsets from the LOCAL base pointer (negative / >= n: outside the own slice)
  PeerCtrl* ctrl = nullptr;
};

struct HaloWait {       // by-value kernel argument of the persistent SpMV kernel
  const PeerCtrl* ctrl = nullptr;  // nullptr: nothing to wait for
  PeerCtrl* ctrl_rw = nullptr;     // for the timeout flag
  int32_t channel = 0;
  int32_t first_item = 0;          // list position of the first slice that stages remote x
  uint32_t peer_mask = 0;          // ranks this rank receives halo entries from
  int32_t pad_ = 0;
  const int32_t* skip0 = nullptr;  // the kernel returns at once if *skip0 or *skip1 is set (solver flags:
  const int32_t* skip1 = nullptr;  // converged / restart pending) - the producer of x skipped its push as well
  // plain sharded SpMV (x in the symmetric arena): the grid's LAST CTA first sends this rank's acknowledgement, waits for
  // the receivers' acknowledgements, stores the boundary rows of x into the neighbours' copies and publishes epoch
  // acked_want + 1; the halo wait then uses the host-counted acked epochs instead of the solvers' device counters
  const AckDesc* ack = nullptr;    // nullptr: solver path (the producer of x pushed)
  unsigned long long acked_want = 0;  // plain sharded SpMVs this rank has finished on the channel
  int64_t own_row0 = 0;            // first global row of this rank (its slice inside the full-layout x)
};

struct SolverWork {  // device scratch of the CG / BiCGStab loops
  double* d_vec[8] = {nullptr};
  int64_t vec_len = 0;
  double* d_scalars = nullptr;   // see solvers.cu
  double* d_partials = nullptr;
  int64_t partials_len = 0;
  unsigned int* d_counters = nullptr;
  unsigned int* d_tickets = nullptr;  // last-CTA tickets of the in-kernel reductions
  unsigned long long* d_trace = nullptr;  // CASK_B200_TRACE: kTraceIters x 3 kernels x 3 timestamps
  int32_t* h_flags = nullptr;    // pinned
  double* h_scalars = nullptr;   // pinned
};

}  // namespace caskb200

struct cask_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  int sm_count = 148;
  int64_t launches = 0;

  cask_b200_design design{};
  bool have_design = false;
  caskb200::Plan plan;
  std::vector<caskb200::RefPartition> ref_parts;
  bool ref_built = false;

  // tuning knobs (cask_b200_set_option)
  double ell_min_fill = 0.75;
  int32_t force_kind = -1;   // -1 auto, 0 staged ELL wherever possible, 1 gather CSR everywhere
  int32_t force_csr_vec = 0;
  int32_t ell_kernel = 1;    // 1 persistent warp-specialised kernel, 0 one CTA per slice
  int32_t persist_ku = 0;    // 0 auto, else 2 or 4
  int32_t persist_ctas = 0;  // coded format only: 0 auto, else CTAs per SM of the persistent kernel (1..3)
  int32_t value_dict = 0;    // 1: staged-ELL values as 8-bit codes into per-slice tables when every slice has <= 256 distinct values;
                             // 2: (value, displacement) pair codes where every slice has <= 255 pairs, else as 1
  int64_t l2_persist_bytes = -1;  // persisting L2 set-aside claimed for evict-last vector accesses (-1: not asked yet)
  int32_t csr_stream = 0;    // 1: CSR-stream variant of the gather-CSR kernel (products staged in shared memory)
  int32_t csr_item_nnz = 4096;  // nonzeros per work item of the CSR-stream variant (its shared-memory footprint)
  int32_t merge_items = 0;   // merge-path tiles: merge items per thread (0: default; 5, 7, 11, 17)
  int32_t col_reorder = -1;  // gather path: -1 auto (hub clustering for large skewed single-rank plans), 0 off, 1 = columns renumbered
                             // by descending reference count (hub clustering), 2 = referenced columns only, in column order (the
                             // numbering of the sharded sparse exchange)
  bool dist_sparse_active = false;  // set by dist.cu while it builds the compact numbering of a sparse exchange
  int32_t csr_kernel = -1;   // gather slices: -1 auto (merge-path tiles from kMergeAutoNnz nonzeros), 0 row-group items, 1 merge-path tiles
  int32_t l2_keep = -1;      // -1 auto (vectors of a solver iteration fit L2), 0 never, 1 always: evict-last on vector accesses
  int32_t dist_sparse_hub = 1;  // sparse exchange: inside every owner's segment of the compact x the most referenced columns come first
  int32_t dist_sparse = 1;   // row-sharded gather plans: 1 = sparse exchange (each rank receives only the x entries its rows
                             // reference, packed by their owners), 0 = every slice broadcast to all
  int32_t ilu_persistent = 1;  // ILU(0) application: 1 = all levels of both solves in ONE cooperative kernel with grid barriers
  int32_t ilu_graph = 1;     // ILU(0) application (when not persistent): 1 = the per-level launches replayed as one CUDA graph, 0 = launched one by one
  int32_t peer_mode = 1;     // 1: halo pushes and scalar all-reduces by own kernels over mapped peer memory; 0: NCCL

  // host-call staging buffers
  double* d_x = nullptr; int64_t d_x_len = 0;
  double* d_y = nullptr; int64_t d_y_len = 0;
  double* h_pinned = nullptr; int64_t h_pinned_len = 0;
  // host-buffer SpMV pipeline: x upload, kernels and y download overlap chunk by chunk
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> pipe_events;
  int32_t host_pipeline_chunks = 16;
  int32_t host_pipeline_ramp = 1;  // chunk sizes grow geometrically from both ends (0: equal chunks)
  // pageable caller buffers (hostcopy.hpp): pinned rings + host copy threads, created on first use
  int32_t host_staging = 1;  // 0: leave pageable buffers to the driver's own staging
  unsigned char* h_ring = nullptr;
  std::vector<cudaEvent_t> ring_events;
  caskb200::HostCopyPool* copy_pool = nullptr;

  caskb200::SolverWork work;
  caskb200::DistState* dist = nullptr;
  caskb200::PrecondState* precond = nullptr;  // preconditioners + work vectors of cask_b200_pcg (precond.cu)
};

namespace caskb200 {

// refformat.cu
int build_ref_partitions(cask_b200_ctx* ctx);
void free_ref_partitions(cask_b200_ctx* ctx);
int spmv_refformat_device(cask_b200_ctx* ctx, const double* d_x, double* d_y);
int refformat_stripe(cudaStream_t s, int64_t* launches, const int32_t* d_colptr, int64_t len, const uint8_t* d_pairs,
                     int32_t ns, int32_t nb, int32_t cache, int32_t w, int64_t m, const double* d_x, double* d_y);

// plan.cu
int build_plan(cask_b200_ctx* ctx);
int build_csr_items(cask_b200_ctx* ctx);
// mode 0 none, -1 automatic, 1 hub clustering, 2 compact numbering of the referenced columns (grouped by owning rank,
// hubs first inside a group, when the owners' first columns are given; else in column order)
int build_col_reorder(cask_b200_ctx* ctx, int mode, const int64_t* owner_first, int world);
void free_plan(cask_b200_ctx* ctx);

// spmv.cu
struct SpmvFusion {           // optional fused epilogue: partial dot products per CTA
  const double* d_dot_with = nullptr;  // if set: partial sum of y[r] * dot_with[r] (local rows)
  double* d_partials = nullptr;        // one partial per CTA of the launch (+offset)
  int fuse_self_dot = 0;               // also accumulate y[r]*y[r]
  ReduceDesc reduce;                   // persistent kernel only: the last CTA sums the partials (and all-reduces them)
  unsigned long long* trace = nullptr; // persistent kernel only: 3 timeline slots (peer.cuh: trace_min / trace_max)
  bool pdl = false;                    // launch with programmatic stream serialization (solver loops)
  int keep_vectors = 0;                // persistent kernel only: x windows, y and dot_with with an L2 evict-last policy
};
// true if one persistent staged-ELL launch covers the whole SpMV (in-kernel reduction / peer halo wait possible)
bool spmv_single_launch(const cask_b200_ctx* ctx);
int launch_spmv(cask_b200_ctx* ctx, const double* d_x_full, double* d_y, int part /*0 all,1 interior,2 boundary*/,
                cudaStream_t stream, const SpmvFusion* fusion, const HaloWait* wait = nullptr);
int launch_spmv_range(cask_b200_ctx* ctx, const double* d_x, double* d_y, int ell_lo, int ell_hi, int csr_lo, int csr_hi,
                      cudaStream_t stream, const SpmvFusion* fusion, const HaloWait* wait = nullptr);
int spmv_num_ctas(cask_b200_ctx* ctx, int part);
int configure_persistent(cask_b200_ctx* ctx);

// dist.cu
int dist_exchange_begin(cask_b200_ctx* ctx, double* d_x_full, cudaStream_t after);
int dist_exchange_wait(cask_b200_ctx* ctx, cudaStream_t consumer);
int dist_allreduce_sum(cask_b200_ctx* ctx, const double* d_local, double* d_global, int count, cudaStream_t stream);
void dist_free(cask_b200_ctx* ctx);
bool dist_active(const cask_b200_ctx* ctx);
int dist_plan_halo(cask_b200_ctx* ctx);
// peer-memory path: usable once the arena is mapped and the halo plan is a handful of contiguous ranges
bool peer_ready(const cask_b200_ctx* ctx);
int peer_ensure_arena(cask_b200_ctx* ctx, int64_t len_full);         // collective
double* peer_vector(cask_b200_ctx* ctx, int channel);                // full-layout vector `channel` of the arena
PushDesc peer_push_desc(cask_b200_ctx* ctx, int channel);            // ctrl == nullptr when the peer path is off
void peer_fill_reduce(cask_b200_ctx* ctx, ReduceDesc* rd);           // adds the all-reduce part when the peer path is on
HaloWait peer_halo_wait(cask_b200_ctx* ctx, int channel);
HaloUpdate peer_halo_update(cask_b200_ctx* ctx, int channel);
int peer_push(cask_b200_ctx* ctx, int channel, cudaStream_t stream); // standalone push of the channel's own slice
int peer_acked_wait(cask_b200_ctx* ctx, int channel, HaloWait* hw);  // fills the in-kernel push part of a plain sharded SpMV
int peer_channel_of(const cask_b200_ctx* ctx, const double* d_full); // arena vector index of a pointer, or -1
int peer_allreduce_partials(cask_b200_ctx* ctx, const double* d_partials, int count, int stride, int nq, double* d_scal,
                            int slot0, const int32_t* d_skip0, const int32_t* d_skip1, cudaStream_t stream);
int peer_check_error(cask_b200_ctx* ctx);

// solvers.cu
void free_solver_work(cask_b200_ctx* ctx);

// precond.cu: everything derived from the current matrix (factors, levels, the override set by precond_set_matrix)
void free_precond(cask_b200_ctx* ctx);

// helpers
int ensure_device(cask_b200_ctx* ctx);

}  // namespace caskb200
