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
//                   slice is broadcast to all.
// NCCL is bound at run time with dlopen("libnccl.so.2") so that the library shares the NCCL that the
// hosting process (e.g. torch) already loaded, and so that single-GPU users need no NCCL at all.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "ctx.cuh"

namespace caskb200 {

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommSplit)(NcclComm, int, int, NcclComm*, void*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
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
  std::vector<std::vector<Range>> recv_from;  // [peer] ranges of x this rank receives
  std::vector<std::vector<Range>> send_to;    // [peer] ranges of its own slice this rank sends
  int64_t n_global = 0;
};

bool dist_active(const cask_b200_ctx* ctx) { return ctx->dist && ctx->dist->world > 1; }

void dist_free(cask_b200_ctx* ctx) {
  if (!ctx->dist) return;
  DistState* d = ctx->dist;
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

// Derives, from the plan's staged runs, which x ranges come from which peer; exchanges the requests so
// every rank also knows what to send; splits the slice lists into interior / halo-dependent.
int dist_plan_halo(cask_b200_ctx* ctx) {
  if (!dist_active(ctx)) return CASK_B200_OK;
  DistState* d = ctx->dist;
  Plan& p = ctx->plan;
  cudaStream_t s = ctx->stream;
  const int W = d->world, me = d->rank;
  d->n_global = p.n_global;
  d->recv_from.assign(W, {});
  d->send_to.assign(W, {});
  const int64_t own_lo = p.row0_global, own_hi = p.row0_global + p.n;

  // runs of all staged slices
  int total_runs = 0;
  for (auto& sd : p.h_slices) if (sd.kind == kSliceStagedEll) total_runs = std::max(total_runs, sd.run_off + sd.nruns);
  std::vector<Run> runs(total_runs);
  if (total_runs) CB_CUDA(cudaMemcpyAsync(runs.data(), p.d_runs, sizeof(Run) * total_runs, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));

  std::vector<Range> remote;
  for (auto& sd : p.h_slices) {
    sd.remote = 0;
    if (sd.kind != kSliceStagedEll) { sd.remote = 1; continue; }
    for (int i = 0; i < sd.nruns; i++) {
      const Run& r = runs[sd.run_off + i];
      const int64_t lo = r.col0, hi = (int64_t)r.col0 + r.len;
      if (lo < own_lo) { remote.push_back({lo, std::min(hi, own_lo) - lo}); sd.remote = 1; }
      if (hi > own_hi) { const int64_t b = std::max(lo, own_hi); remote.push_back({b, hi - b}); sd.remote = 1; }
    }
  }
  // merge
  std::sort(remote.begin(), remote.end(), [](const Range& a, const Range& b) { return a.col0 < b.col0; });
  std::vector<Range> merged;
  for (auto& r : remote) {
    if (r.len <= 0) continue;
    if (!merged.empty() && r.col0 <= merged.back().col0 + merged.back().len) {
      merged.back().len = std::max(merged.back().len, r.col0 + r.len - merged.back().col0);
    } else merged.push_back(r);
  }
  // split by owner
  int64_t nranges = 0;
  for (auto& r : merged) {
    int64_t lo = r.col0, hi = r.col0 + r.len;
    for (int q = 0; q < W && lo < hi; q++) {
      int64_t q0, qn;
      stripe_of(p.n_global, W, q, &q0, &qn);
      const int64_t a = std::max(lo, q0), b = std::min(hi, q0 + qn);
      if (a < b && q != me) { d->recv_from[q].push_back({a, b - a}); nranges++; }
    }
  }
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
    // ranges requested from me by q: all[cap*q + 1 + me] of them
    std::vector<int64_t*> d_send(W, nullptr), d_recv(W, nullptr);
    std::vector<std::vector<int64_t>> h_recv(W);
    CB_NCCL(g_nccl.GroupStart());
    for (int q = 0; q < W; q++) {
      if (q == me) continue;
      const size_t n_out = d->recv_from[q].size(), n_in = (size_t)all[cap * q + 1 + me];
      if (n_out) {
        CB_CUDA(cudaMalloc(&d_send[q], sizeof(int64_t) * 2 * n_out));
        CB_CUDA(cudaMemcpyAsync(d_send[q], d->recv_from[q].data(), sizeof(int64_t) * 2 * n_out, cudaMemcpyHostToDevice, s));
        CB_NCCL(g_nccl.Send(d_send[q], 2 * n_out, kNcclInt64, q, d->comm_halo, s));
      }
      if (n_in) {
        CB_CUDA(cudaMalloc(&d_recv[q], sizeof(int64_t) * 2 * n_in));
        CB_NCCL(g_nccl.Recv(d_recv[q], 2 * n_in, kNcclInt64, q, d->comm_halo, s));
      }
    }
    CB_NCCL(g_nccl.GroupEnd());
    for (int q = 0; q < W; q++) {
      const size_t n_in = q == me ? 0 : (size_t)all[cap * q + 1 + me];
      if (n_in) {
        d->send_to[q].resize(n_in);
        CB_CUDA(cudaMemcpyAsync(d->send_to[q].data(), d_recv[q], sizeof(int64_t) * 2 * n_in, cudaMemcpyDeviceToHost, s));
      }
    }
    CB_CUDA(cudaStreamSynchronize(s));
    for (int q = 0; q < W; q++) { cudaFree(d_send[q]); cudaFree(d_recv[q]); }
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
  return CASK_B200_OK;
}

// Starts the exchange of x entries on the communication stream once everything queued on `after`
// (the producer of the local slice) has finished.
int dist_exchange_begin(cask_b200_ctx* ctx, double* d_full, cudaStream_t after) {
  DistState* d = ctx->dist;
  const Plan& p = ctx->plan;
  cudaStream_t cs = ctx->comm_stream;
  CB_CUDA(cudaEventRecord(d->ev_ready, after));
  CB_CUDA(cudaStreamWaitEvent(cs, d->ev_ready, 0));
  const int W = d->world, me = d->rank;
  CB_NCCL(g_nccl.GroupStart());
  if (d->allgather) {
    int64_t my0, myn;
    stripe_of(p.n_global, W, me, &my0, &myn);
    for (int q = 0; q < W; q++) {
      if (q == me) continue;
      int64_t q0, qn;
      stripe_of(p.n_global, W, q, &q0, &qn);
      if (myn) CB_NCCL(g_nccl.Send(d_full + my0, (size_t)myn, kNcclFloat64, q, d->comm_halo, cs));
      if (qn) CB_NCCL(g_nccl.Recv(d_full + q0, (size_t)qn, kNcclFloat64, q, d->comm_halo, cs));
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
  return CASK_B200_OK;
}

extern "C" int cask_b200_shard_rows(int64_t n, int32_t world, int32_t rank, int64_t* row0, int64_t* nrows) {
  if (world < 1 || rank < 0 || rank >= world || n < 0 || !row0 || !nrows)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "shard_rows: bad arguments");
  stripe_of(n, world, rank, row0, nrows);
  return CASK_B200_OK;
}

extern "C" int cask_b200_dist_halo_counts(cask_b200_ctx* ctx, int64_t* recv_counts) {
  if (!ctx || !ctx->dist || !recv_counts) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "dist_halo_counts: not initialised");
  DistState* d = ctx->dist;
  for (int q = 0; q < d->world; q++) {
    int64_t c = 0;
    if (d->allgather) {
      int64_t q0, qn;
      stripe_of(d->n_global, d->world, q, &q0, &qn);
      c = q == d->rank ? 0 : qn;
    } else if (q < (int)d->recv_from.size()) {
      for (auto& r : d->recv_from[q]) c += r.len;
    }
    recv_counts[q] = c;
  }
  return CASK_B200_OK;
}
