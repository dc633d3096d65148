// One process per GPU: row-sharded SpMV / solvers over NCCL (NVLink 5 / NVSwitch).
//
// The reference replicates x into every pipe's memory (src/runtime/Spmv.cpp:165-170,247).  Here every
// rank owns the row stripe Spmv::preprocess would give pipe `rank` (Spmv.cpp:334-364) and the matching
// slice of every vector.  Vectors that feed an SpMV live in "full layout" (global length, own slice in
// place), so received halo entries land at their global positions and the kernels need no index
// translation.  What must travel is derived from the staged x windows of the plan:
//   HALO mode       each rank receives only the column ranges its slices stage from peers
//                   (stencils: one plane per neighbour), grouped ncclSend/ncclRecv straight from / into
//                   the vectors (no packing), on a communication stream that overlaps the interior slices;
//   ALLGATHER mode  irregular matrices (gather-CSR slices present or too many ranges): every rank's
//                   slice is broadcast to all;
//   SPARSE mode     the same matrices when every slice of every rank runs the merge-path gather kernel (default,
//                   option dist_sparse): a rank receives only the x entries its rows reference.  The plan renumbers the
//                   referenced columns compactly (plan.cu: build_col_reorder mode 2), every owner packs the entries each
//                   peer asked for at preprocess time into one send buffer (sparse_pack_kernel), and grouped
//                   ncclSend/ncclRecv move them straight into the segments of the receiver's compact x.  R-MAT scale
//                   25 over 8 ranks: a rank references about 5 M of the 33.5 M columns, so it receives tens of MB
//                   instead of 235 MB and its gathers run on an L2-resident vector.
// NCCL is bound at run time with dlopen("libnccl.so.2") so that the library shares the NCCL that the
// hosting process (e.g. torch) already loaded, and so that single-GPU users need no NCCL at all.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "ctx.cuh"
#include "peer.cuh"

namespace caskb200 {

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclInt8 = 0, kNcclInt32 = 2, kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommSplit)(NcclComm, int, int, NcclComm*, void*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return CASK_B200_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(CASK_B200_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define CB_SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                 \
  if (!g_nccl.field) return fail(CASK_B200_ERR_NCCL, std::string("libnccl lacks ") + name)
  CB_SYM(GetUniqueId, "ncclGetUniqueId");
  CB_SYM(CommInitRank, "ncclCommInitRank");
  CB_SYM(CommSplit, "ncclCommSplit");
  CB_SYM(CommDestroy, "ncclCommDestroy");
  CB_SYM(AllReduce, "ncclAllReduce");
  CB_SYM(AllGather, "ncclAllGather");
  CB_SYM(Broadcast, "ncclBroadcast");
  CB_SYM(Send, "ncclSend");
  CB_SYM(Recv, "ncclRecv");
  CB_SYM(GroupStart, "ncclGroupStart");
  CB_SYM(GroupEnd, "ncclGroupEnd");
  CB_SYM(GetErrorString, "ncclGetErrorString");
#undef CB_SYM
  g_nccl.lib = h;
  return CASK_B200_OK;
}

#define CB_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    int _r = (expr);                                                                                    \
    if (_r != 0) return fail(CASK_B200_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

struct Range { int64_t col0, len; };

}  // namespace

struct DistState {
  int rank = 0, world = 1;
  NcclComm comm_halo = nullptr, comm_red = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  bool allgather = false;
  // SPARSE mode (see the header comment): segment q of the compact x = the referenced columns owned by rank q
  bool sparse = false;
  bool ell_all_local = true;       // staged slices (if any) read own columns only: they need nothing from the exchange
  std::vector<int64_t> seg;        // [world + 1] segment boundaries inside the compact x (plan.d_xperm)
  std::vector<int64_t> send_off;   // [world + 1] offsets into the send list / send buffer, by destination rank
  int32_t* d_send_list = nullptr;  // global columns (all inside the own slice) the peers asked for, by destination
  double* d_sendbuf = nullptr;
  bool sparse_peer = false;        // the exchange runs as ONE kernel that stores into the peers' compact x over NVLink
  SparsePushDesc sparse_push;      // (peer-memory layer, channel 1 of the arena); false: grouped ncclSend / ncclRecv
  std::vector<std::vector<Range>> recv_from;  // [peer] ranges of x this rank receives
  std::vector<std::vector<Range>> send_to;    // [peer] ranges of its own slice this rank sends
  int64_t n_global = 0;
  // peer-memory layer
  bool peer_mapped = false;      // arenas of all ranks are mapped into this process
  bool peer_failed = false;      // IPC mapping was tried and is not available on this machine: NCCL from now on
  bool peer_plan_ok = false;     // every rank's halo plan fits a PushDesc and runs on the persistent kernel
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0, vec_stride = 0;
  int64_t arena_len = 0;         // doubles per full-layout vector
  unsigned char* peer_base[kMaxPeers] = {nullptr};
  uint32_t recv_mask = 0;
  unsigned long long acked_pushes[kHaloChannels] = {0};  // flow-controlled pushes launched per channel (same on every rank)
  // first row of every rank's stripe (+ n_global): the reference rule rows_per = n / world (Spmv.cpp:334-364) unless the
  // caller handed preprocess_shard another contiguous partition (e.g. stripes of equal nonzero count for a power-law matrix)
  std::vector<int64_t> bounds;
  // in-kernel push of the plain sharded SpMV: device copy of its descriptor per channel, rebuilt when stale
  AckDesc* d_ack[kHaloChannels] = {nullptr};
  unsigned long long plan_stamp = 0, ack_plan_stamp[kHaloChannels] = {0};
  unsigned char* ack_arena[kHaloChannels] = {nullptr};
};

constexpr size_t kCtrlBytes = 4096;
static_assert(sizeof(PeerCtrl) <= kCtrlBytes, "control block must fit its reservation");

static void peer_unmap(DistState* d) {
  for (int q = 0; q < kMaxPeers; q++) {
    if (d->peer_base[q] && q != d->rank) cudaIpcCloseMemHandle(d->peer_base[q]);
    d->peer_base[q] = nullptr;
  }
  if (d->arena) cudaFree(d->arena);
  d->arena = nullptr;
  d->arena_bytes = 0;
  d->arena_len = 0;
  d->peer_mapped = false;
  for (auto& a : d->acked_pushes) a = 0;  // the control block (and its acknowledgement counters) went with the arena
}

bool dist_active(const cask_b200_ctx* ctx) { return ctx->dist && ctx->dist->world > 1; }

void dist_free(cask_b200_ctx* ctx) {
  if (!ctx->dist) return;
  DistState* d = ctx->dist;
  peer_unmap(d);
  for (auto& a : d->d_ack) { cudaFree(a); a = nullptr; }
  cudaFree(d->d_send_list);
  cudaFree(d->d_sendbuf);
  if (d->comm_red && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm_red);
  if (d->comm_halo && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm_halo);
  if (d->ev_ready) cudaEventDestroy(d->ev_ready);
  if (d->ev_done) cudaEventDestroy(d->ev_done);
  delete d;
  ctx->dist = nullptr;
}

static void stripe_of(int64_t n, int world, int r, int64_t* row0, int64_t* nrows) {
  const int64_t rpp = n / world;  // Spmv.cpp:334
  *row0 = rpp * r;
  *nrows = r == world - 1 ? n - rpp * (world - 1) : rpp;  // Spmv.cpp:362-364: remainder to the last
}

// Pure host arithmetic (no CUDA, no NCCL; exported as cask_b200_halo_plan_host and exercised by the world-size-2
// gloo tests on CPU): the x windows (runs) this rank stages, clipped to the columns it does NOT own, merged, and
// split by owner according to the reference's row striping.  Returns the number of (peer, range) pairs.
static void owner_range(const int64_t* bounds, int64_t n_global, int world, int q, int64_t* q0, int64_t* qn) {
  if (bounds) { *q0 = bounds[q]; *qn = bounds[q + 1] - bounds[q]; }
  else stripe_of(n_global, world, q, q0, qn);
}

static int64_t halo_ranges_from_runs(int64_t n_global, int world, int me, const int64_t* run_col0, const int64_t* run_len,
                                     int64_t nruns, std::vector<std::vector<Range>>* recv_from, const int64_t* bounds = nullptr) {
  int64_t own_lo, own_n;
  owner_range(bounds, n_global, world, me, &own_lo, &own_n);
  const int64_t own_hi = own_lo + own_n;
  std::vector<Range> remote;
  for (int64_t i = 0; i < nruns; i++) {
    const int64_t lo = run_col0[i], hi = run_col0[i] + run_len[i];
    if (lo < own_lo) remote.push_back({lo, std::min(hi, own_lo) - lo});
    if (hi > own_hi) { const int64_t b = std::max(lo, own_hi); remote.push_back({b, hi - b}); }
  }
  std::sort(remote.begin(), remote.end(), [](const Range& a, const Range& b) { return a.col0 < b.col0; });
  std::vector<Range> merged;
  for (auto& r : remote) {
    if (r.len <= 0) continue;
    if (!merged.empty() && r.col0 <= merged.back().col0 + merged.back().len) {
      merged.back().len = std::max(merged.back().len, r.col0 + r.len - merged.back().col0);
    } else merged.push_back(r);
  }
  recv_from->assign(world, {});
  int64_t nranges = 0;
  for (auto& r : merged) {
    int64_t lo = r.col0, hi = r.col0 + r.len;
    for (int q = 0; q < world && lo < hi; q++) {
      int64_t q0, qn;
      owner_range(bounds, n_global, world, q, &q0, &qn);
      const int64_t a = std::max(lo, q0), b = std::min(hi, q0 + qn);
      if (a < b && q != me) { (*recv_from)[q].push_back({a, b - a}); nranges++; }
    }
  }
  return nranges;
}


// ---- SPARSE mode ---------------------------------------------------------------------------------------------------
// sendbuf[i] = x[send_list[i]] for the entries the peers asked for, and the own segment of the compact x directly.
// Inside a destination the list is in the RECEIVER's order (its hubs first, then ascending runs of equally often
// referenced columns), so beyond the hubs the reads are a handful of forward sweeps over the own slice.
__global__ void __launch_bounds__(256) sparse_pack_kernel(const double* __restrict__ x_full, const int32_t* __restrict__ send_list,
                                                          int64_t n_send, double* __restrict__ sendbuf,
                                                          const int32_t* __restrict__ own_cols, int64_t n_own,
                                                          double* __restrict__ xp_own) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_send + n_own; i += stride) {
    if (i < n_send) sendbuf[i] = __ldg(x_full + send_list[i]);
    else xp_own[i - n_send] = __ldg(x_full + own_cols[i - n_send]);
  }
}

// The same exchange without NCCL (peer_mode 1, IPC arena mapped): the compact x lives in channel 1 of the symmetric arena
// and ONE kernel per SpMV does everything.  Protocol (host-counted epochs, as the acknowledged boundary-row push of the
// plain sharded SpMV, peer.cuh: acked_push):
//   1. CTA 0 tells every peer how many exchanges this rank has finished (stream order: the SpMV that read the previous
//      epoch is complete when this kernel runs); every CTA waits until the ranks it stores for have said the same -
//      only then may their compact x be overwritten.  Acknowledgements leave before any wait: no deadlock.
//   2. all CTAs: entries the peers asked for go straight into their compact x (coalesced 8-byte stores over NVLink),
//      the own segment into the local one.
//   3. every CTA fences its stores system-wide and takes a ticket; the last CTA publishes epoch done + 1 to the ranks it
//      stored for, then waits for the same epoch from the ranks it receives from - the gather kernel queued behind this
//      kernel starts with all of x in place.
// The grid is at most 2 CTAs per SM (always co-resident), every wait carries the peer timeout.
__global__ void __launch_bounds__(256) sparse_push_kernel(const double* __restrict__ x_full, const int32_t* __restrict__ send_list,
                                                          const int32_t* __restrict__ own_cols, int64_t n_own,
                                                          double* __restrict__ xp_own, const SparsePushDesc sd) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (blockIdx.x == 0)
      for (int q = 0; q < sd.world; q++)
        if (q != sd.me) st_volatile_u64(&sd.peers[q]->halo_ack[1][sd.me], sd.done);
    if (sd.done)
      for (int q = 0; q < sd.world; q++)
        if (sd.send_mask & (1u << q)) peer_wait_ge(&sd.ctrl->halo_ack[1][q], sd.done, &sd.ctrl->error);
  }
  __syncthreads();
  const int64_t n_send = sd.send_off[sd.world];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n_send + n_own; i += stride) {
    if (i < n_send) {
      int q = 0;
      while (i >= sd.send_off[q + 1]) q++;
      sd.dst[q][i - sd.send_off[q]] = __ldg(x_full + send_list[i]);
    } else {
      xp_own[i - n_send] = __ldg(x_full + own_cols[i - n_send]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();  // this CTA's peer stores (ordered before by the barrier) are visible system-wide
    s_last = atomicAdd(&sd.ctrl->push_ticket[1], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) {
    sd.ctrl->push_ticket[1] = 0;
    __threadfence_system();
    for (int q = 0; q < sd.world; q++)
      if (sd.send_mask & (1u << q)) st_volatile_u64(&sd.peers[q]->halo_flag_acked[1][sd.me], sd.done + 1);
  }
  if (tid < sd.world && ((sd.recv_mask >> tid) & 1u)) peer_wait_ge(&sd.ctrl->halo_flag_acked[1][tid], sd.done + 1, &sd.ctrl->error);
}

// Pure host arithmetic of the sparse exchange (exported as cask_b200_sparse_segments_host / cask_b200_sparse_send_plan_host
// and exercised by the gloo tests on CPU).
static void sparse_segments(const int64_t* bounds, int world, const int32_t* need, int64_t count, int64_t* seg) {
  for (int q = 0; q <= world; q++)
    seg[q] = std::lower_bound(need, need + count, bounds[q], [](int32_t a, int64_t b) { return (int64_t)a < b; }) - need;
}
static void sparse_send_plan(const int64_t* all_seg, int world, int me, int64_t* send_off, int64_t* dst_off) {
  const size_t cap = (size_t)world + 1;
  send_off[0] = 0;
  for (int q = 0; q < world; q++) {
    send_off[q + 1] = send_off[q] + (q == me ? 0 : all_seg[cap * q + me + 1] - all_seg[cap * q + me]);
    if (dst_off) dst_off[q] = all_seg[cap * q + me];
  }
}

// Collective.  Decides (all ranks together) whether the sparse exchange applies, renumbers the columns, and tells every
// owner which of its entries each peer needs.
static int dist_plan_sparse(cask_b200_ctx* ctx) {
  DistState* d = ctx->dist;
  Plan& p = ctx->plan;
  cudaStream_t s = ctx->stream;
  const int W = d->world, me = d->rank;
  d->sparse = false;
  d->sparse_peer = false;
  cudaFree(d->d_send_list); cudaFree(d->d_sendbuf);
  d->d_send_list = nullptr; d->d_sendbuf = nullptr;
  struct Tmp {
    int64_t* q = nullptr;
    ~Tmp() { cudaFree(q); }
  } tmp;
  const size_t cap = (size_t)W + 1;
  CB_CUDA(cudaMalloc(&tmp.q, sizeof(int64_t) * cap * (size_t)(W + 1)));
  std::vector<int64_t> mine(cap, 0), all(cap * W, 0);
  auto gather_all = [&]() -> int {
    CB_CUDA(cudaMemcpyAsync(tmp.q + cap * W, mine.data(), sizeof(int64_t) * cap, cudaMemcpyHostToDevice, s));
    CB_NCCL(g_nccl.AllGather(tmp.q + cap * W, tmp.q, cap, kNcclInt64, d->comm_halo, s));
    CB_CUDA(cudaMemcpyAsync(all.data(), tmp.q, sizeof(int64_t) * cap * W, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    return CASK_B200_OK;
  };
  // 1. every slice that reads remote columns runs the merge-path gather kernel, on every rank (a rank without rows
  //    qualifies trivially; staged slices that stage own columns only read the full-layout x as before)
  mine[0] = d->allgather && ctx->dist_sparse != 0 && d->ell_all_local && (p.n_csr == 0 || p.csr_merge) && p.m < (1ll << 30) ? 1 : 0;
  CB_TRY(gather_all());
  for (int q = 0; q < W; q++)
    if (!all[cap * q]) return CASK_B200_OK;
  // 2. compact column numbering; segment q = referenced columns owned by rank q
  ctx->dist_sparse_active = true;
  const int rc_reorder = build_col_reorder(ctx, 2, ctx->dist_sparse_hub != 0 ? d->bounds.data() : nullptr, W);
  ctx->dist_sparse_active = false;
  CB_TRY(rc_reorder);
  if (!p.d_perm) return fail(CASK_B200_ERR_RUNTIME, "sparse exchange: the column renumbering was refused");
  std::vector<int32_t> need((size_t)p.cols_used);
  if (p.cols_used) CB_CUDA(cudaMemcpyAsync(need.data(), p.d_perm, sizeof(int32_t) * need.size(), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  d->seg.assign((size_t)W + 1, 0);
  sparse_segments(d->bounds.data(), W, need.data(), (int64_t)need.size(), d->seg.data());
  // 3. every rank's segment boundaries: what rank q needs from owner r is its segment r, so one all-gather of the
  //    boundaries tells every owner how much it sends to whom, and (peer-memory variant) where its entries start
  //    inside every peer's compact x
  for (int q = 0; q <= W; q++) mine[q] = d->seg[q];
  CB_TRY(gather_all());
  auto seg_of = [&](int q, int r) -> int64_t { return all[cap * q + r]; };  // boundary r of rank q's compact x
  d->send_off.assign((size_t)W + 1, 0);
  sparse_send_plan(all.data(), W, me, d->send_off.data(), nullptr);
  int64_t max_used = 0;
  for (int q = 0; q < W; q++) max_used = std::max(max_used, seg_of(q, W));
  const int64_t total_send = d->send_off[W];
  CB_CUDA(cudaMalloc(&d->d_send_list, sizeof(int32_t) * (size_t)std::max<int64_t>(total_send, 1)));
  CB_CUDA(cudaMalloc(&d->d_sendbuf, sizeof(double) * (size_t)std::max<int64_t>(total_send, 1)));
  // 4. the requests themselves: ONE all-gather of the ranks' referenced-column lists (padded to the longest; d_perm holds
  //    m entries, the tail beyond cols_used is never looked at), from which every owner cuts the pieces addressed to it.
  //    An all-gather runs over the communicator's warmed ring; point-to-point sends between all pairs would pay NCCL's
  //    lazy connection set-up here (measured: 5.5 s at 8 ranks).
  if (max_used > 0) {
    struct TmpI {
      int32_t* q = nullptr;
      ~TmpI() { cudaFree(q); }
    } lists;
    CB_CUDA(cudaMalloc(&lists.q, sizeof(int32_t) * (size_t)max_used * (size_t)W));
    CB_NCCL(g_nccl.AllGather(p.d_perm, lists.q, (size_t)max_used, kNcclInt32, d->comm_halo, s));
    for (int q = 0; q < W; q++) {
      const int64_t n_in = d->send_off[q + 1] - d->send_off[q];
      if (n_in)
        CB_CUDA(cudaMemcpyAsync(d->d_send_list + d->send_off[q], lists.q + (size_t)q * (size_t)max_used + seg_of(q, me),
                                sizeof(int32_t) * (size_t)n_in, cudaMemcpyDeviceToDevice, s));
    }
    CB_CUDA(cudaStreamSynchronize(s));
  }
  d->sparse = true;
  // 5. peer-memory variant: the compact x moves into channel 1 of the symmetric arena (gather plans use no other
  //    channel: cask_b200_dist_vector and the solvers' peer path are for staged-ELL plans)
  if (ctx->peer_mode != 0 && W <= kMaxPeers) {
    CB_TRY(peer_ensure_arena(ctx, p.m));  // collective; all ranks agree on peer_mapped
    if (d->peer_mapped) {
      if (p.xperm_owned) cudaFree(p.d_xperm);
      p.d_xperm = peer_vector(ctx, 1);
      p.xperm_owned = false;
      SparsePushDesc sp;
      sp.ctrl = reinterpret_cast<PeerCtrl*>(d->arena);
      sp.me = me;
      sp.world = W;
      for (int q = 0; q < W; q++) {
        sp.peers[q] = reinterpret_cast<PeerCtrl*>(d->peer_base[q]);
        sp.dst[q] = reinterpret_cast<double*>(d->peer_base[q] + kCtrlBytes + d->vec_stride) + seg_of(q, me);
        sp.send_off[q + 1] = d->send_off[q + 1];
        if (d->send_off[q + 1] > d->send_off[q]) sp.send_mask |= 1u << q;
        if (q != me && d->seg[q + 1] > d->seg[q]) sp.recv_mask |= 1u << q;
      }
      d->sparse_push = sp;
      d->sparse_peer = true;
    }
  }
  return CASK_B200_OK;
}

// Derives, from the plan's staged runs, which x ranges come from which peer; exchanges the requests so
// every rank also knows what to send; splits the slice lists into interior / halo-dependent.
int dist_plan_halo(cask_b200_ctx* ctx) {
  if (!dist_active(ctx)) return CASK_B200_OK;
  DistState* d = ctx->dist;
  Plan& p = ctx->plan;
  cudaStream_t s = ctx->stream;
  const int W = d->world, me = d->rank;
  d->n_global = p.n_global;
  d->plan_stamp++;
  d->recv_from.assign(W, {});
  d->send_to.assign(W, {});
  const int64_t own_lo = p.row0_global, own_hi = p.row0_global + p.n;
  {
    // every rank's first row: the partition the callers chose (contiguous, in rank order, covering all rows)
    int64_t* d_b = nullptr;
    CB_CUDA(cudaMalloc(&d_b, sizeof(int64_t) * (size_t)(W + 1)));
    CB_CUDA(cudaMemcpyAsync(d_b + W, &own_lo, sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CB_NCCL(g_nccl.AllGather(d_b + W, d_b, 1, kNcclInt64, d->comm_halo, s));
    d->bounds.assign((size_t)W + 1, 0);
    CB_CUDA(cudaMemcpyAsync(d->bounds.data(), d_b, sizeof(int64_t) * (size_t)W, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_b);
    d->bounds[W] = p.n_global;
    bool ok = d->bounds[0] == 0 && d->bounds[me + 1] == own_hi;
    for (int q = 0; q < W; q++) ok = ok && d->bounds[q] <= d->bounds[q + 1];
    if (!ok) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess_shard: the ranks' row ranges must be contiguous, in rank order, and cover all rows");
  }

  // runs of all staged slices
  int total_runs = 0;
  for (auto& sd : p.h_slices) if (sd.kind == kSliceStagedEll) total_runs = std::max(total_runs, sd.run_off + sd.nruns);
  std::vector<Run> runs(total_runs);
  if (total_runs) CB_CUDA(cudaMemcpyAsync(runs.data(), p.d_runs, sizeof(Run) * total_runs, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));

  d->ell_all_local = true;  // no staged slice stages a column of another rank (e.g. the empty slices of a power-law matrix)
  for (auto& sd : p.h_slices) {
    sd.remote = 0;
    if (sd.kind != kSliceStagedEll) { sd.remote = 1; continue; }
    for (int i = 0; i < sd.nruns; i++) {
      const Run& r = runs[sd.run_off + i];
      if (r.col0 < own_lo || (int64_t)r.col0 + r.len > own_hi) sd.remote = 1;
    }
    if (sd.remote) d->ell_all_local = false;
  }
  std::vector<int64_t> run_col0(runs.size()), run_len(runs.size());
  for (size_t i = 0; i < runs.size(); i++) { run_col0[i] = runs[i].col0; run_len[i] = runs[i].len; }
  const int64_t nranges = halo_ranges_from_runs(p.n_global, W, me, run_col0.data(), run_len.data(), (int64_t)runs.size(),
                                                &d->recv_from, d->bounds.data());
  bool want_allgather = p.n_csr > 0 || nranges > 512;
  // agree on the mode and exchange the request lists through NCCL itself
  int64_t* d_buf = nullptr;
  const size_t cap = 2 * (size_t)W + 2;
  CB_CUDA(cudaMalloc(&d_buf, sizeof(int64_t) * cap * (size_t)(W + 1)));
  std::vector<int64_t> mine(cap, 0), all(cap * W, 0);
  mine[0] = want_allgather ? 1 : 0;
  for (int q = 0; q < W; q++) mine[1 + q] = (int64_t)d->recv_from[q].size();
  CB_CUDA(cudaMemcpyAsync(d_buf + cap * W, mine.data(), sizeof(int64_t) * cap, cudaMemcpyHostToDevice, s));
  CB_NCCL(g_nccl.AllGather(d_buf + cap * W, d_buf, cap, kNcclInt64, d->comm_halo, s));
  CB_CUDA(cudaMemcpyAsync(all.data(), d_buf, sizeof(int64_t) * cap * W, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  bool any_allgather = false;
  for (int q = 0; q < W; q++) any_allgather |= all[cap * q] != 0;
  d->allgather = any_allgather;
  cudaFree(d_buf);
  if (!d->allgather) {
    // ranges requested from me by q: all[cap*q + 1 + me] of them.  The request lists travel in ONE all-gather (every rank's
    // list, ordered by owner, padded to the longest) from which each owner cuts its piece: point-to-point sends here
    // would make NCCL build its send/recv connections between all pairs at preprocess time (about a second at 8 ranks),
    // and the peer-memory path never needs them.
    int64_t max_req = 0;
    for (int q = 0; q < W; q++) {
      int64_t t = 0;
      for (int r = 0; r < W; r++) t += all[cap * q + 1 + r];
      max_req = std::max(max_req, t);
    }
    if (max_req > 0) {
      std::vector<int64_t> req((size_t)max_req * 2, 0), got((size_t)max_req * 2 * (size_t)W, 0);
      size_t k = 0;
      for (int q = 0; q < W; q++)
        for (auto& r : d->recv_from[q]) { req[k++] = r.col0; req[k++] = r.len; }
      int64_t* d_req = nullptr;
      CB_CUDA(cudaMalloc(&d_req, sizeof(int64_t) * (size_t)max_req * 2 * (size_t)(W + 1)));
      cudaError_t e = cudaMemcpyAsync(d_req + (size_t)max_req * 2 * (size_t)W, req.data(), sizeof(int64_t) * req.size(), cudaMemcpyHostToDevice, s);
      int nrc = 0;
      if (e == cudaSuccess) nrc = g_nccl.AllGather(d_req + (size_t)max_req * 2 * (size_t)W, d_req, (size_t)max_req * 2, kNcclInt64, d->comm_halo, s);
      if (e == cudaSuccess && nrc == 0) e = cudaMemcpyAsync(got.data(), d_req, sizeof(int64_t) * got.size(), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess && nrc == 0) e = cudaStreamSynchronize(s);
      cudaFree(d_req);
      if (nrc != 0) return fail(CASK_B200_ERR_NCCL, std::string("ncclAllGather (halo requests): ") + g_nccl.GetErrorString(nrc));
      CB_CUDA(e);
      for (int q = 0; q < W; q++) {
        if (q == me) continue;
        int64_t off = 0;
        for (int r = 0; r < me; r++) off += all[cap * q + 1 + r];
        const int64_t n_in = all[cap * q + 1 + me];
        d->send_to[q].resize((size_t)n_in);
        for (int64_t i = 0; i < n_in; i++) {
          const size_t at = ((size_t)q * (size_t)max_req + (size_t)(off + i)) * 2;
          d->send_to[q][(size_t)i] = Range{got[at], got[at + 1]};
        }
      }
    }
  } else {
    for (auto& sd : p.h_slices) sd.remote = 1;
  }
  // interior slices first
  std::vector<int32_t> ell, csr;
  int n_ell_int = 0, n_csr_int = 0;
  for (int pass = 0; pass < 2; pass++)
    for (int32_t i = 0; i < p.nslices; i++) {
      const SliceDesc& sd = p.h_slices[i];
      if ((sd.remote != 0) != (pass == 1)) continue;
      if (sd.kind == kSliceStagedEll) { ell.push_back(i); if (!pass) n_ell_int++; }
      else { csr.push_back(i); if (!pass) n_csr_int++; }
    }
  p.n_ell_interior = n_ell_int;
  p.n_csr_interior = n_csr_int;
  if (!ell.empty()) CB_CUDA(cudaMemcpyAsync(p.d_list_ell, ell.data(), sizeof(int32_t) * ell.size(), cudaMemcpyHostToDevice, s));
  if (!csr.empty()) CB_CUDA(cudaMemcpyAsync(p.d_list_csr, csr.data(), sizeof(int32_t) * csr.size(), cudaMemcpyHostToDevice, s));
  CB_CUDA(cudaStreamSynchronize(s));
  p.h_list_ell = ell;
  p.h_list_csr = csr;
  CB_TRY(build_csr_items(ctx));  // items follow the (re-ordered) slice list

  // Peer-memory path: every rank's sends must fit one PushDesc and every rank must run the persistent staged-ELL
  // kernel on all its slices; the ranks agree through one all-reduce.
  size_t nsend = 0, nrecv = 0;
  d->recv_mask = 0;
  for (int q = 0; q < W; q++) {
    nsend += d->send_to[q].size();
    nrecv += d->recv_from[q].size();
    if (!d->recv_from[q].empty()) d->recv_mask |= 1u << q;
  }
  const bool mine_ok = !d->allgather && W <= kMaxPeers && nsend <= (size_t)kMaxPush && nrecv <= (size_t)kMaxPush && p.n_csr == 0 &&
                       ctx->ell_kernel == 1 && p.persist_ku != 0 && ctx->peer_mode != 0;
  int64_t* d_ok = nullptr;
  CB_CUDA(cudaMalloc(&d_ok, sizeof(int64_t)));
  const int64_t bad = mine_ok ? 0 : 1;
  int64_t bad_all = 1;
  CB_CUDA(cudaMemcpyAsync(d_ok, &bad, sizeof(bad), cudaMemcpyHostToDevice, s));
  CB_NCCL(g_nccl.AllReduce(d_ok, d_ok, 1, kNcclInt64, kNcclSum, d->comm_halo, s));
  CB_CUDA(cudaMemcpyAsync(&bad_all, d_ok, sizeof(bad_all), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_ok);
  d->peer_plan_ok = bad_all == 0;
  return dist_plan_sparse(ctx);
}

// ---- peer-memory layer: arena, halo push, scalar all-reduce --------------------------------------------------
bool peer_ready(const cask_b200_ctx* ctx) {
  return dist_active(ctx) && ctx->peer_mode != 0 && ctx->dist->peer_mapped && ctx->dist->peer_plan_ok;
}

// Collective.  (Re)allocates the symmetric arena so that each of its kHaloChannels vectors holds len_full doubles,
// exchanges the IPC handles over NCCL and maps every peer's arena.  If the machine does not allow the mapping the
// ranks agree to stay on NCCL (peer_failed).
int peer_ensure_arena(cask_b200_ctx* ctx, int64_t len_full) {
  if (!dist_active(ctx) || ctx->peer_mode == 0) return CASK_B200_OK;
  DistState* d = ctx->dist;
  if (d->peer_failed || d->world > kMaxPeers) return CASK_B200_OK;
  if (d->peer_mapped && d->arena_len >= len_full) return CASK_B200_OK;
  cudaStream_t s = ctx->stream;
  const int W = d->world, me = d->rank;
  CB_CUDA(cudaStreamSynchronize(s));
  CB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  peer_unmap(d);
  d->vec_stride = (sizeof(double) * (size_t)std::max<int64_t>(len_full, 2) + 255) & ~(size_t)255;
  d->arena_bytes = kCtrlBytes + kHaloChannels * d->vec_stride;
  CB_CUDA(cudaMalloc(&d->arena, d->arena_bytes));
  CB_CUDA(cudaMemsetAsync(d->arena, 0, d->arena_bytes, s));
  cudaIpcMemHandle_t mine;
  int64_t bad = cudaIpcGetMemHandle(&mine, d->arena) == cudaSuccess ? 0 : 1;
  if (bad) cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  unsigned char* d_h = nullptr;
  CB_CUDA(cudaMalloc(&d_h, 64 * (size_t)(W + 1) + 8));
  std::vector<cudaIpcMemHandle_t> all(W);
  CB_CUDA(cudaMemcpyAsync(d_h + 64 * (size_t)W, &mine, 64, cudaMemcpyHostToDevice, s));
  CB_NCCL(g_nccl.AllGather(d_h + 64 * (size_t)W, d_h, 64, kNcclInt8, d->comm_halo, s));
  CB_CUDA(cudaMemcpyAsync(all.data(), d_h, 64 * (size_t)W, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  d->peer_base[me] = d->arena;
  for (int q = 0; q < W && !bad; q++) {
    if (q == me) continue;
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      bad = 1;
    } else {
      d->peer_base[q] = static_cast<unsigned char*>(ptr);
    }
  }
  int64_t* d_bad = reinterpret_cast<int64_t*>(d_h + 64 * (size_t)(W + 1));
  int64_t bad_all = 1;
  CB_CUDA(cudaMemcpyAsync(d_bad, &bad, sizeof(bad), cudaMemcpyHostToDevice, s));
  CB_NCCL(g_nccl.AllReduce(d_bad, d_bad, 1, kNcclInt64, kNcclSum, d->comm_halo, s));
  CB_CUDA(cudaMemcpyAsync(&bad_all, d_bad, sizeof(bad_all), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_h);
  if (bad_all) {
    peer_unmap(d);
    d->peer_failed = true;
    return CASK_B200_OK;
  }
  d->arena_len = len_full;
  d->peer_mapped = true;
  return CASK_B200_OK;
}

double* peer_vector(cask_b200_ctx* ctx, int channel) {
  DistState* d = ctx->dist;
  return reinterpret_cast<double*>(d->arena + kCtrlBytes + (size_t)channel * d->vec_stride);
}

PushDesc peer_push_desc(cask_b200_ctx* ctx, int channel) {
  PushDesc pd;
  for (int i = 0; i < kMaxPush; i++) { pd.lo[i] = pd.hi[i] = 0; pd.dst[i] = nullptr; pd.peer_ctrl[i] = nullptr; }
  if (!peer_ready(ctx)) return pd;
  DistState* d = ctx->dist;
  const int64_t own_lo = ctx->plan.row0_global;
  pd.channel = channel;
  pd.me = d->rank;
  pd.ctrl = reinterpret_cast<PeerCtrl*>(d->arena);
  int k = 0;
  for (int q = 0; q < d->world; q++)
    for (auto& r : d->send_to[q]) {
      pd.lo[k] = r.col0 - own_lo;
      pd.hi[k] = r.col0 + r.len - own_lo;
      pd.dst[k] = reinterpret_cast<double*>(d->peer_base[q] + kCtrlBytes + (size_t)channel * d->vec_stride) + own_lo;
      pd.peer_ctrl[k] = reinterpret_cast<PeerCtrl*>(d->peer_base[q]);
      k++;
    }
  pd.nsend = k;
  return pd;
}

void peer_fill_reduce(cask_b200_ctx* ctx, ReduceDesc* rd) {
  if (!peer_ready(ctx)) return;
  DistState* d = ctx->dist;
  rd->ctrl = reinterpret_cast<PeerCtrl*>(d->arena);
  for (int q = 0; q < kMaxPeers; q++) rd->peers[q] = q < d->world ? reinterpret_cast<PeerCtrl*>(d->peer_base[q]) : nullptr;
  rd->me = d->rank;
  rd->world = d->world;
}

HaloWait peer_halo_wait(cask_b200_ctx* ctx, int channel) {
  HaloWait w;
  if (!peer_ready(ctx)) return w;
  DistState* d = ctx->dist;
  if (d->recv_mask == 0) return w;
  w.ctrl = reinterpret_cast<const PeerCtrl*>(d->arena);
  w.ctrl_rw = reinterpret_cast<PeerCtrl*>(d->arena);
  w.channel = channel;
  w.first_item = ctx->plan.n_ell_interior;
  w.peer_mask = d->recv_mask;
  return w;
}

HaloUpdate peer_halo_update(cask_b200_ctx* ctx, int channel) {
  HaloUpdate h;
  for (int i = 0; i < kMaxPush; i++) h.lo[i] = h.hi[i] = 0;
  if (!peer_ready(ctx)) return h;
  DistState* d = ctx->dist;
  const int64_t own_lo = ctx->plan.row0_global;
  int k = 0;
  for (int q = 0; q < d->world; q++)
    for (auto& r : d->recv_from[q]) {
      if (k == kMaxPush) break;
      h.lo[k] = r.col0 - own_lo;
      h.hi[k] = r.col0 + r.len - own_lo;
      k++;
    }
  h.nrecv = k;
  h.channel = channel;
  h.peer_mask = d->recv_mask;
  h.ctrl = reinterpret_cast<PeerCtrl*>(d->arena);
  return h;
}

namespace {

// standalone push: copies the pushed ranges of the channel's own slice into the peers' vectors, then signals
__global__ void __launch_bounds__(256) halo_push_kernel(const double* __restrict__ own, const PushDesc pd) {
  // blockIdx.y = range, blockIdx.x strides over it
  const int s = blockIdx.y;
  if (s < pd.nsend) {
    double* dst = pd.dst[s];
    for (int64_t i = pd.lo[s] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pd.hi[s]; i += (int64_t)gridDim.x * blockDim.x)
      dst[i] = own[i];
  }
  push_signal(pd);
}

struct PeerCtrlPtrs { PeerCtrl* p[kMaxPeers]; };

// Deterministic sum of per-CTA partials (index order) followed by a one-shot all-reduce over peer memory: every
// rank stores its local sums into slot [seq & 3][me] of every rank's control block, publishes the sequence number
// with a release store, acquires the W flags of its own block and adds the W contributions in rank order - the
// same order on every rank, so all ranks hold bit-identical scalars and take identical convergence decisions.
__global__ void __launch_bounds__(1024)
peer_allreduce_kernel(const double* __restrict__ partials, int count, int stride, int nq, double* __restrict__ scal,
                      int slot0, const int32_t* __restrict__ skip0, const int32_t* __restrict__ skip1, PeerCtrl* ctrl,
                      const PeerCtrlPtrs peers, int me, int world) {
  if ((skip0 && *skip0) || (skip1 && *skip1)) return;
  __shared__ double red[32];
  __shared__ double local[2];
  __shared__ double contrib[kMaxPeers][2];
  __shared__ unsigned long long seq_s;
  const int tid = threadIdx.x;
  for (int q = 0; q < nq; q++) {
    double v = 0.0;
    for (int i = tid; i < count; i += blockDim.x) v += partials[(size_t)q * stride + i];
#pragma unroll
    for (int dlt = 16; dlt; dlt >>= 1) v += __shfl_xor_sync(0xffffffffu, v, dlt);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
      local[q] = t;
    }
    __syncthreads();
  }
  if (tid == 0) {
    seq_s = ctrl->ar_seq + 1;
    ctrl->ar_seq = seq_s;
  }
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int slot = (int)(seq & 3ull);
  if (tid < world) {
    PeerCtrl* pc = peers.p[tid];
    for (int q = 0; q < nq; q++) pc->ar_val[slot][me][q] = local[q];
    __threadfence_system();
    st_release_sys_u64(&pc->ar_flag[slot][me], seq);
    peer_wait_ge(&ctrl->ar_flag[slot][tid], seq, &ctrl->error);
    for (int q = 0; q < nq; q++) contrib[tid][q] = *reinterpret_cast<volatile double*>(&ctrl->ar_val[slot][tid][q]);
  }
  __syncthreads();
  if (tid < nq) {
    double t = 0.0;
    for (int r = 0; r < world; r++) t += contrib[r][tid];
    scal[slot0 + tid] = t;
  }
}

}  // namespace

int peer_push(cask_b200_ctx* ctx, int channel, cudaStream_t stream) {
  const PushDesc pd = peer_push_desc(ctx, channel);
  if (pd.ctrl == nullptr) return CASK_B200_OK;
  int64_t longest = 1;
  for (int i = 0; i < pd.nsend; i++) longest = std::max(longest, pd.hi[i] - pd.lo[i]);
  const dim3 grid((unsigned)std::min<int64_t>((longest + 255) / 256, 64), (unsigned)std::max(pd.nsend, 1));
  halo_push_kernel<<<grid, 256, 0, stream>>>(peer_vector(ctx, channel) + ctx->plan.row0_global, pd);
  ctx->launches++;
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

// Plain sharded SpMV with x in the arena: the descriptor of the in-kernel boundary-row push (device-resident, rebuilt
// when the plan or the arena changes) and this call's host-counted epoch.
int peer_acked_wait(cask_b200_ctx* ctx, int channel, HaloWait* hw) {
  if (!peer_ready(ctx)) return fail(CASK_B200_ERR_RUNTIME, "acknowledged push needs the peer-memory path");
  DistState* d = ctx->dist;
  if (!d->d_ack[channel]) CB_CUDA(cudaMalloc(&d->d_ack[channel], sizeof(AckDesc)));
  if (d->ack_plan_stamp[channel] != d->plan_stamp || d->ack_arena[channel] != d->arena) {
    AckDesc ad;
    for (int q = 0; q < kMaxPeers; q++) ad.peers[q] = q < d->world ? reinterpret_cast<PeerCtrl*>(d->peer_base[q]) : nullptr;
    for (int q = 0; q < d->world; q++)
      if (!d->send_to[q].empty()) ad.send_mask |= 1u << q;
    ad.me = d->rank;
    ad.world = d->world;
    ad.push = peer_push_desc(ctx, channel);
    CB_CUDA(cudaMemcpyAsync(d->d_ack[channel], &ad, sizeof(ad), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaStreamSynchronize(ctx->stream));  // `ad` is a stack object
    d->ack_plan_stamp[channel] = d->plan_stamp;
    d->ack_arena[channel] = d->arena;
  }
  // every rank with at least one neighbour takes part, receivers and senders alike
  hw->ctrl = reinterpret_cast<const PeerCtrl*>(d->arena);
  hw->ctrl_rw = reinterpret_cast<PeerCtrl*>(d->arena);
  hw->channel = channel;
  hw->first_item = ctx->plan.n_ell_interior;
  hw->peer_mask = d->recv_mask;
  hw->ack = d->d_ack[channel];
  hw->acked_want = d->acked_pushes[channel]++;
  hw->own_row0 = ctx->plan.row0_global;
  return CASK_B200_OK;
}

int peer_channel_of(const cask_b200_ctx* ctx, const double* d_full) {
  if (!ctx->dist || !ctx->dist->peer_mapped || !d_full) return -1;
  const DistState* d = ctx->dist;
  for (int ch = 0; ch < kHaloChannels; ch++)
    if (d_full == reinterpret_cast<const double*>(d->arena + kCtrlBytes + (size_t)ch * d->vec_stride)) return ch;
  return -1;
}

int peer_allreduce_partials(cask_b200_ctx* ctx, const double* d_partials, int count, int stride, int nq, double* d_scal,
                            int slot0, const int32_t* d_skip0, const int32_t* d_skip1, cudaStream_t stream) {
  DistState* d = ctx->dist;
  PeerCtrlPtrs pp;
  for (int q = 0; q < kMaxPeers; q++) pp.p[q] = q < d->world ? reinterpret_cast<PeerCtrl*>(d->peer_base[q]) : nullptr;
  peer_allreduce_kernel<<<1, 1024, 0, stream>>>(d_partials, count, stride, nq, d_scal, slot0, d_skip0, d_skip1,
                                                reinterpret_cast<PeerCtrl*>(d->arena), pp, d->rank, d->world);
  ctx->launches++;
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

// after a solve / at a synchronisation point: did any wait on a peer time out?  (halo pushes, in-kernel all-reduces and
// the sparse exchange all raise the same flag of this rank's control block)
int peer_check_error(cask_b200_ctx* ctx) {
  if (!dist_active(ctx) || !ctx->dist->peer_mapped || !ctx->dist->arena) return CASK_B200_OK;
  int err = 0;
  CB_CUDA(cudaMemcpy(&err, ctx->dist->arena + offsetof(PeerCtrl, error), sizeof(int), cudaMemcpyDeviceToHost));
  if (err) return fail(CASK_B200_ERR_RUNTIME, "peer-memory wait timed out: a rank stopped participating");
  return CASK_B200_OK;
}

// Starts the exchange of x entries on the communication stream once everything queued on `after`
// (the producer of the local slice) has finished.
int dist_exchange_begin(cask_b200_ctx* ctx, double* d_full, cudaStream_t after) {
  DistState* d = ctx->dist;
  const Plan& p = ctx->plan;
  cudaStream_t cs = ctx->comm_stream;
  const int W = d->world, me = d->rank;
  if (d->sparse && d->sparse_peer) {
    const int64_t n_send = d->send_off[W], n_own = d->seg[me + 1] - d->seg[me];
    SparsePushDesc sp = d->sparse_push;
    sp.done = d->acked_pushes[1]++;
    const int64_t ctas = std::max<int64_t>(1, (n_send + n_own + 255) / 256);
    sparse_push_kernel<<<(int)std::min<int64_t>(ctas, (int64_t)ctx->sm_count * 2), 256, 0, after>>>(
        d_full, d->d_send_list, p.d_perm + d->seg[me], n_own, p.d_xperm + d->seg[me], sp);
    ctx->launches++;
    CB_CUDA(cudaGetLastError());
    return CASK_B200_OK;
  }
  if (d->sparse) {
    // own entries into the send buffer and into the own segment of the compact x, then one grouped exchange
    const int64_t n_send = d->send_off[W], n_own = d->seg[me + 1] - d->seg[me];
    if (n_send + n_own) {
      const int64_t ctas = (n_send + n_own + 255) / 256;
      sparse_pack_kernel<<<(int)std::min<int64_t>(ctas, (int64_t)ctx->sm_count * 8), 256, 0, after>>>(
          d_full, d->d_send_list, n_send, d->d_sendbuf, p.d_perm + d->seg[me], n_own, p.d_xperm + d->seg[me]);
      ctx->launches++;
      CB_CUDA(cudaGetLastError());
    }
    CB_CUDA(cudaEventRecord(d->ev_ready, after));
    CB_CUDA(cudaStreamWaitEvent(cs, d->ev_ready, 0));
    CB_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < W; q++) {
      if (q == me) continue;
      const int64_t n_out = d->send_off[q + 1] - d->send_off[q], n_in = d->seg[q + 1] - d->seg[q];
      if (n_out) CB_NCCL(g_nccl.Send(d->d_sendbuf + d->send_off[q], (size_t)n_out, kNcclFloat64, q, d->comm_halo, cs));
      if (n_in) CB_NCCL(g_nccl.Recv(p.d_xperm + d->seg[q], (size_t)n_in, kNcclFloat64, q, d->comm_halo, cs));
    }
    CB_NCCL(g_nccl.GroupEnd());
    CB_CUDA(cudaEventRecord(d->ev_done, cs));
    ctx->launches++;
    return CASK_B200_OK;
  }
  CB_CUDA(cudaEventRecord(d->ev_ready, after));
  CB_CUDA(cudaStreamWaitEvent(cs, d->ev_ready, 0));
  CB_NCCL(g_nccl.GroupStart());
  if (d->allgather) {
    // every rank's slice to everybody, in place.  Stripes may differ widely in rows (equal-nonzero stripes of a power-law
    // matrix: the last rank of 8 owns half of x), so W point-to-point copies per slice would make its owner's NVLink
    // egress the bottleneck (7 x 134 MB for R-MAT scale 25: 1.2 ms); one broadcast per root moves each slice once
    // through NCCL's ring / NVSwitch multicast instead.
    for (int q = 0; q < W; q++) {
      int64_t q0, qn;
      owner_range(d->bounds.data(), p.n_global, W, q, &q0, &qn);
      if (qn) CB_NCCL(g_nccl.Broadcast(d_full + q0, d_full + q0, (size_t)qn, kNcclFloat64, q, d->comm_halo, cs));
    }
  } else {
    for (int q = 0; q < W; q++) {
      if (q == me) continue;
      for (auto& r : d->send_to[q]) CB_NCCL(g_nccl.Send(d_full + r.col0, (size_t)r.len, kNcclFloat64, q, d->comm_halo, cs));
      for (auto& r : d->recv_from[q]) CB_NCCL(g_nccl.Recv(d_full + r.col0, (size_t)r.len, kNcclFloat64, q, d->comm_halo, cs));
    }
  }
  CB_NCCL(g_nccl.GroupEnd());
  CB_CUDA(cudaEventRecord(d->ev_done, cs));
  ctx->launches++;
  return CASK_B200_OK;
}

int dist_exchange_wait(cask_b200_ctx* ctx, cudaStream_t consumer) {
  if (ctx->dist->sparse && ctx->dist->sparse_peer) return CASK_B200_OK;  // the push kernel on the consumer's stream did the waiting
  CB_CUDA(cudaStreamWaitEvent(consumer, ctx->dist->ev_done, 0));
  return CASK_B200_OK;
}

int dist_allreduce_sum(cask_b200_ctx* ctx, const double* d_local, double* d_global, int count, cudaStream_t stream) {
  CB_NCCL(g_nccl.AllReduce(d_local, d_global, (size_t)count, kNcclFloat64, kNcclSum, ctx->dist->comm_red, stream));
  ctx->launches++;
  return CASK_B200_OK;
}

}  // namespace caskb200

using namespace caskb200;

extern "C" int cask_b200_nccl_unique_id(void* out_128_bytes) {
  if (!out_128_bytes) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "nccl_unique_id: null buffer");
  CB_TRY(load_nccl());
  NcclUniqueId id;
  CB_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(out_128_bytes, &id, sizeof(id));
  return CASK_B200_OK;
}

extern "C" int cask_b200_dist_init(cask_b200_ctx* ctx, int32_t rank, int32_t world, const void* unique_id) {
  if (!ctx || world < 1 || rank < 0 || rank >= world) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_init: bad rank/world");
  CB_TRY(ensure_device(ctx));
  dist_free(ctx);
  DistState* d = new DistState();
  d->rank = rank;
  d->world = world;
  ctx->dist = d;
  if (world == 1) return CASK_B200_OK;
  if (!unique_id) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_init: unique id required for world > 1");
  CB_TRY(load_nccl());
  NcclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  CB_NCCL(g_nccl.CommInitRank(&d->comm_halo, world, id, rank));
  CB_NCCL(g_nccl.CommSplit(d->comm_halo, 0, rank, &d->comm_red, nullptr));
  CB_CUDA(cudaEventCreateWithFlags(&d->ev_ready, cudaEventDisableTiming));
  CB_CUDA(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
  // NCCL builds its channels and connections at the first collective of a communicator (0.8 s at 2 ranks, 2.9 s at 8 on
  // an NVSwitch box): paid here, once, instead of inside the first preprocess_shard (round 1's "preprocess_s grows with N")
  {
    int64_t* d_one = nullptr;
    CB_CUDA(cudaMalloc(&d_one, sizeof(int64_t) * (size_t)(world + 1)));
    CB_CUDA(cudaMemsetAsync(d_one, 0, sizeof(int64_t) * (size_t)(world + 1), ctx->stream));
    CB_NCCL(g_nccl.AllReduce(d_one, d_one, 1, kNcclInt64, kNcclSum, d->comm_halo, ctx->stream));
    CB_NCCL(g_nccl.AllGather(d_one + world, d_one, 1, kNcclInt64, d->comm_halo, ctx->stream));
    CB_NCCL(g_nccl.AllReduce(d_one, d_one, 1, kNcclInt64, kNcclSum, d->comm_red, ctx->stream));
    CB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_one);
  }
  return CASK_B200_OK;
}

extern "C" int cask_b200_shard_rows(int64_t n, int32_t world, int32_t rank, int64_t* row0, int64_t* nrows) {
  if (world < 1 || rank < 0 || rank >= world || n < 0 || !row0 || !nrows)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "shard_rows: bad arguments");
  stripe_of(n, world, rank, row0, nrows);
  return CASK_B200_OK;
}

extern "C" int cask_b200_halo_plan_host(int64_t n_global, int32_t world, int32_t rank, int64_t nruns, const int64_t* run_col0,
                                        const int64_t* run_len, int64_t capacity, int32_t* out_peer, int64_t* out_col0,
                                        int64_t* out_len, int64_t* out_count) {
  if (world < 1 || rank < 0 || rank >= world || n_global < 0 || nruns < 0 || (nruns && (!run_col0 || !run_len)) || !out_count)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "halo_plan_host: bad arguments");
  std::vector<std::vector<Range>> recv;
  const int64_t total = halo_ranges_from_runs(n_global, world, rank, run_col0, run_len, nruns, &recv);
  *out_count = total;
  if (total > capacity) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "halo_plan_host: output capacity too small");
  int64_t k = 0;
  for (int q = 0; q < world; q++)
    for (auto& r : recv[q]) {
      if (out_peer) out_peer[k] = q;
      if (out_col0) out_col0[k] = r.col0;
      if (out_len) out_len[k] = r.len;
      k++;
    }
  return CASK_B200_OK;
}

extern "C" int cask_b200_sparse_segments_host(const int64_t* bounds, int32_t world, const int32_t* need, int64_t count, int64_t* seg) {
  if (!bounds || world < 1 || count < 0 || (count && !need) || !seg) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "sparse_segments_host: bad arguments");
  for (int q = 0; q < world; q++)
    if (bounds[q] > bounds[q + 1]) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "sparse_segments_host: bounds must not decrease");
  // need[] is grouped by owner, owners ascending (column order is the special case); the order inside a group is free
  int q = 0;
  for (int64_t i = 0; i < count; i++) {
    if (need[i] < bounds[0] || need[i] >= bounds[world]) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "sparse_segments_host: column outside the owners' ranges");
    if (need[i] < bounds[q]) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "sparse_segments_host: need[] must be grouped by owner, owners ascending");
    while (need[i] >= bounds[q + 1]) q++;
  }
  sparse_segments(bounds, world, need, count, seg);
  return CASK_B200_OK;
}

extern "C" int cask_b200_sparse_send_plan_host(const int64_t* all_seg, int32_t world, int32_t rank, int64_t* send_off, int64_t* dst_off) {
  if (!all_seg || world < 1 || rank < 0 || rank >= world || !send_off || !dst_off)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "sparse_send_plan_host: bad arguments");
  sparse_send_plan(all_seg, world, rank, send_off, dst_off);
  return CASK_B200_OK;
}

extern "C" int cask_b200_dist_peer_active(cask_b200_ctx* ctx, int32_t* active) {
  if (!ctx || !active) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_peer_active: null");
  *active = peer_ready(ctx) ? 1 : 0;
  return CASK_B200_OK;
}

extern "C" int cask_b200_dist_halo_counts(cask_b200_ctx* ctx, int64_t* recv_counts) {
  if (!ctx || !ctx->dist || !recv_counts) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_halo_counts: not initialised");
  DistState* d = ctx->dist;
  for (int q = 0; q < d->world; q++) {
    int64_t c = 0;
    if (d->sparse) {
      c = q == d->rank ? 0 : d->seg[q + 1] - d->seg[q];
    } else if (d->allgather) {
      int64_t q0, qn;
      owner_range(d->bounds.empty() ? nullptr : d->bounds.data(), d->n_global, d->world, q, &q0, &qn);
      c = q == d->rank ? 0 : qn;
    } else if (q < (int)d->recv_from.size()) {
      for (auto& r : d->recv_from[q]) c += r.len;
    }
    recv_counts[q] = c;
  }
  return CASK_B200_OK;
}
