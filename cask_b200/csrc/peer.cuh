// Device-side helpers of the peer-memory layer (structs in ctx.cuh, host side in dist.cu).
//
// The reference replicates x into every pipe's DRAM with one host write per pipe (src/runtime/Spmv.cpp:165-170,247).
// Here the kernel that PRODUCES a vector stores the entries its neighbours stage straight into their copy of the
// vector over NVLink (plain st.global on IPC-mapped pointers), the grid's last CTA publishes an epoch number with a
// system-scope release store, and the consuming SpMV's producer warp acquires it before issuing the TMA copies of
// the first halo-dependent slice.  No NCCL call, no extra launch, no host involvement per iteration.
#pragma once

#include "ctx.cuh"

namespace caskb200 {

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// L2 residency control.  When the vectors a solver touches per iteration fit the 126 MB L2 (row-sharded runs: the
// per-rank slices of x, r, p, Ap), they are loaded and stored with an evict-last policy while the matrix streams
// through with evict-first: the vector kernels and the SpMV's x windows then run out of L2 instead of HBM.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ld_keep(const double* p, unsigned long long pol, bool keep) {
  double v;
  if (keep) asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  else v = *p;
  return v;
}
__device__ __forceinline__ void st_keep(double* p, double v, unsigned long long pol, bool keep) {
  if (keep) asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
  else *p = v;
}

constexpr unsigned long long kPeerTimeoutNs = 8000000000ull;  // a peer that stays silent for 8 s is gone: report, never hang

// Spins until *flag >= want.  Returns false (and raises *err) on timeout so that a lost peer surfaces as an error
// code on the host instead of a hung GPU.
__device__ __forceinline__ bool peer_wait_ge(const unsigned long long* flag, unsigned long long want, int* err) {
  if (ld_acquire_sys_u64(flag) >= want) return true;
  const unsigned long long t0 = globaltimer_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; i++)
      if (ld_acquire_sys_u64(flag) >= want) return true;
    if (globaltimer_ns() - t0 > kPeerTimeoutNs) {
      if (err) *err = 1;
      return false;
    }
  }
}

// Low-latency exchange of a double (the "LL" idea of NCCL): each half of the value travels with the 32-bit sequence
// number in a single 8-byte store, which is atomic, so the receiver needs no separate flag and the sender no fence.
__device__ __forceinline__ void ll_send(unsigned long long* slot2, double v, unsigned int seq32) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  st_volatile_u64(slot2, ((unsigned long long)seq32 << 32) | (bits & 0xffffffffull));
  st_volatile_u64(slot2 + 1, ((unsigned long long)seq32 << 32) | (bits >> 32));
}
__device__ __forceinline__ double ll_recv(const unsigned long long* slot2, unsigned int seq32, int* err) {
  unsigned long long a = ld_volatile_u64(slot2), b = ld_volatile_u64(slot2 + 1);
  if ((unsigned int)(a >> 32) != seq32 || (unsigned int)(b >> 32) != seq32) {
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
      a = ld_volatile_u64(slot2);
      b = ld_volatile_u64(slot2 + 1);
      if ((unsigned int)(a >> 32) == seq32 && (unsigned int)(b >> 32) == seq32) break;
      if (globaltimer_ns() - t0 > kPeerTimeoutNs) {
        if (err) *err = 1;
        break;
      }
    }
  }
  return __longlong_as_double((long long)((b << 32) | (a & 0xffffffffull)));
}

// Programmatic dependent launch: kernels of the solver loops are launched with the stream-serialization attribute,
// trigger their dependents at once and wait for their predecessor before touching memory - the launch latency of
// kernel i+1 hides behind the tail of kernel i.  Both instructions are no-ops in a plain launch.
// Order: wait, THEN trigger - when a kernel releases its dependents its own predecessor is complete, so at most
// two consecutive kernels overlap and a dependent may pre-load, before its own wait, anything its immediate
// predecessor does not write.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}

// Finishes a grid-wide dot product inside the producing kernel.  Called by ALL threads of EVERY CTA; t0 / t1 are
// the CTA's partial sums (valid in thread 0).  The last CTA to arrive adds the partials in a fixed order
// (deterministic) and, when row-sharded on the peer path, performs the one-shot all-reduce: it stores the local
// sums into slot [seq & 3][me] of every rank's control block (low-latency 8-byte {data, sequence} stores), polls
// the W slots of its own block and adds the W contributions in rank order - the same order on every rank, so all
// ranks hold bit-identical scalars and take identical convergence decisions.
// Returns true in every thread of the CTA that wrote out[] (after a CTA barrier, so out[] may be read at once by any of
// its threads - the solver loops hang their scalar recurrences on it), false elsewhere.
__device__ __forceinline__ bool grid_finish_reduce(const ReduceDesc& rd, double t0, double t1, int cta, int ncta,
                                                   unsigned int gen0 = 0u) {
  __shared__ int s_last;
  __shared__ double s_red[2][32];
  __shared__ double s_tot[2];
  __shared__ double s_contrib[kMaxPeers][2];
  __shared__ unsigned long long s_seq;
  const int tid = threadIdx.x;
  if (tid == 0) {
    rd.partials[cta] = t0;
    if (rd.nq > 1) rd.partials[rd.stride + cta] = t1;
    __threadfence();
    s_last = atomicAdd(rd.ticket, 1u) == (unsigned)(ncta - 1);
  }
  __syncthreads();
  if (!s_last) {
    if (rd.gen) {  // grid barrier: wait until the last CTA has published the result (gen0 was read at kernel entry)
      if (tid == 0) {
        // co-residency of the grid is checked on the host (occupancy query); should the SMs nevertheless be taken away
        // (another stream of the process, MPS), the launch fails loudly after the timeout instead of spinning forever
        const unsigned long long t_begin = globaltimer_ns();
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int*>(rd.gen) == gen0)
          if ((++spins & 0x3ffu) == 0 && globaltimer_ns() - t_begin > kPeerTimeoutNs) asm volatile("trap;");
      }
      __syncthreads();
      __threadfence();
    }
    return false;
  }
  __threadfence();
  for (int q = 0; q < rd.nq; q++) {
    double v = 0.0;
    for (int i = tid; i < ncta; i += blockDim.x) v += __ldcg(rd.partials + (size_t)q * rd.stride + i);
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((tid & 31) == 0) s_red[q][tid >> 5] = v;
  }
  __syncthreads();
  if (tid == 0) {
    for (int q = 0; q < rd.nq; q++) {
      double t = 0.0;
      for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) t += s_red[q][w];
      s_tot[q] = t;
    }
    *rd.ticket = 0;
    if (rd.ctrl) {
      s_seq = rd.ctrl->ar_seq + 1;
      rd.ctrl->ar_seq = s_seq;
    }
  }
  __syncthreads();
  if (rd.ctrl == nullptr) {
    if (tid < rd.nq) rd.out[tid] = s_tot[tid];
    if (rd.gen) {
      __syncthreads();
      if (tid == 0) { __threadfence(); atomicExch(rd.gen, gen0 + 1u); }
    }
    __syncthreads();
    return true;
  }
  const unsigned long long seq = s_seq;
  const int slot = (int)(seq & 3ull);
  const unsigned int seq32 = (unsigned int)seq;
  if (tid < rd.world) {
    PeerCtrl* pc = rd.peers[tid];
    for (int q = 0; q < rd.nq; q++) ll_send(&pc->ar_ll[slot][rd.me][q][0], s_tot[q], seq32);
  }
  if (rd.publish_only) return false;
  if (tid < rd.world)
    for (int q = 0; q < rd.nq; q++) s_contrib[tid][q] = ll_recv(&rd.ctrl->ar_ll[slot][tid][q][0], seq32, &rd.ctrl->error);
  __syncthreads();
  if (tid < rd.nq) {
    double t = 0.0;
    for (int r = 0; r < rd.world; r++) t += s_contrib[r][tid];
    rd.out[tid] = t;
  }
  if (rd.gen) {
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicExch(rd.gen, gen0 + 1u); }
  }
  __syncthreads();
  return true;
}

// In-situ timeline (profiling runs: CASK_B200_TRACE=<file>): per kernel, the earliest CTA entry, the earliest CTA
// past its dependency wait and the latest CTA exit, in globaltimer nanoseconds.  ncu cannot profile a multi-rank
// job and serialises kernels; this shows the real gaps between the kernels of a solver iteration on every rank.
__device__ __forceinline__ void trace_min(unsigned long long* slot) {
  if (slot && threadIdx.x == 0) atomicMin(slot, globaltimer_ns());
}
__device__ __forceinline__ void trace_max(unsigned long long* slot) {
  if (slot && threadIdx.x == 0) atomicMax(slot, globaltimer_ns());
}

// Consumer side of a publish-only reduction (first quantity): every CTA gathers the W contributions of the latest
// collective from its own control block and adds them in rank order.  Called by all threads; s_c is CTA scratch.
__device__ __forceinline__ double peer_gather_sum(const GatherDesc& g, double* s_c) {
  const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(&g.ctrl->ar_seq);
  const int slot = (int)(seq & 3ull);
  if ((int)threadIdx.x < g.world)
    s_c[threadIdx.x] = ll_recv(&g.ctrl->ar_ll[slot][threadIdx.x][0][0], (unsigned int)seq, &g.ctrl->error);
  __syncthreads();
  double t = 0.0;
  for (int r = 0; r < g.world; r++) t += s_c[r];
  __syncthreads();
  return t;
}

// true if some pushed range intersects local rows [lo, hi)
__device__ __forceinline__ bool push_overlaps(const PushDesc& d, int64_t lo, int64_t hi) {
  bool any = false;
#pragma unroll
  for (int s = 0; s < kMaxPush; s++)
    if (s < d.nsend) any |= (lo < d.hi[s]) & (hi > d.lo[s]);
  return any;
}

// true if local row i lies inside a pushed range
__device__ __forceinline__ bool push_contains(const PushDesc& d, int64_t i) {
  bool any = false;
#pragma unroll
  for (int s = 0; s < kMaxPush; s++)
    if (s < d.nsend) any |= (i >= d.lo[s]) & (i < d.hi[s]);
  return any;
}

// stores entry i (local row index) of the vector into every peer copy that stages it
__device__ __forceinline__ void push_store(const PushDesc& d, int64_t i, double v) {
#pragma unroll
  for (int s = 0; s < kMaxPush; s++)
    if (s < d.nsend && i >= d.lo[s] && i < d.hi[s]) d.dst[s][i] = v;
}

// Called by ALL threads of EVERY CTA of the pushing grid after their push_store calls: the last CTA to arrive
// bumps the channel's epoch and publishes it to the peers.
// The epoch advances on every rank of a peer-mode job, including ranks that have nothing to send, because a
// consumer waits for the epoch number of its OWN channel counter.
__device__ __forceinline__ void push_signal(const PushDesc& d) {
  if (d.ctrl == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's peer stores (ordered before by the barrier) are visible system-wide
    const unsigned int ticket = atomicAdd(&d.ctrl->push_ticket[d.channel], 1u);
    if (ticket == gridDim.x * gridDim.y - 1) {
      d.ctrl->push_ticket[d.channel] = 0;
      const unsigned long long seq = d.ctrl->push_seq[d.channel] + 1;
      d.ctrl->push_seq[d.channel] = seq;
      __threadfence_system();
#pragma unroll
      for (int s = 0; s < kMaxPush; s++)
        if (s < d.nsend) st_release_sys_u64(&d.peer_ctrl[s]->halo_flag[d.channel][d.me], seq);
    }
  }
}

// Consumer side: block until every peer in the mask has published the epoch this rank itself has pushed.
__device__ __forceinline__ void halo_wait(const HaloWait& w) {
  if (w.ack) {  // plain sharded SpMV: host-counted epochs
    for (int q = 0; q < kMaxPeers; q++)
      if (w.peer_mask & (1u << q)) peer_wait_ge(&w.ctrl->halo_flag_acked[w.channel][q], w.acked_want + 1, &w.ctrl_rw->error);
    asm volatile("fence.proxy.async;" ::: "memory");
    return;
  }
  const unsigned long long want = w.ctrl->push_seq[w.channel];
  for (int q = 0; q < kMaxPeers; q++)
    if (w.peer_mask & (1u << q)) peer_wait_ge(&w.ctrl->halo_flag[w.channel][q], want, &w.ctrl_rw->error);
  asm volatile("fence.proxy.async;" ::: "memory");  // the data is read next by TMA (async proxy)
}

// Plain sharded SpMV, called by ALL threads of the grid's last CTA before it turns to its slices (it owns one slice less
// than the first CTAs whenever the slices do not divide evenly): acknowledgement out, acknowledgements in, boundary
// rows of x into the neighbours' copies, epoch out.  The halo rows sit at fixed positions of the receiver's vector, so
// epoch k + 1 may only be written once the receiver has finished the SpMV that read epoch k: every rank tells ALL ranks
// how many SpMVs it has finished on the channel (stream order: the previous one is complete when this kernel runs),
// then waits for that count from the ranks that stage its rows.  Acknowledgements leave before any wait, so the ranks
// cannot deadlock; the waits carry the peer timeout.
__device__ __forceinline__ void acked_push(const HaloWait& w, const double* __restrict__ x_full) {
  const AckDesc& ad = *w.ack;
  const PushDesc& pd = ad.push;
  if (threadIdx.x == 0) {
    // plain (relaxed, system-scope) stores, one per lane-independent target: an acknowledgement orders nothing - the reads
    // it vouches for finished with the previous kernel - and a RELEASE store per peer would serialise seven NVLink round
    // trips in front of this CTA's work (measured: +40 us per step at 8 ranks, profiles/r2i_scaling_table.md)
    for (int q = 0; q < ad.world; q++)
      if (q != ad.me) st_volatile_u64(&ad.peers[q]->halo_ack[pd.channel][ad.me], w.acked_want);
    if (w.acked_want)
      for (int q = 0; q < ad.world; q++)
        if (ad.send_mask & (1u << q)) peer_wait_ge(&pd.ctrl->halo_ack[pd.channel][q], w.acked_want, &pd.ctrl->error);
  }
  __syncthreads();
  for (int s = 0; s < pd.nsend; s++) {
    double* dst = pd.dst[s];                 // rebased: dst[i] is the peer's entry for LOCAL row i
    const double* src = x_full + w.own_row0; // this rank's slice inside its full-layout vector
    for (int64_t i = pd.lo[s] + threadIdx.x; i < pd.hi[s]; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's row stores (ordered before by the barrier) are visible system-wide; then the epochs
    for (int s = 0; s < pd.nsend; s++) st_volatile_u64(&pd.peer_ctrl[s]->halo_flag_acked[pd.channel][pd.me], w.acked_want + 1);
  }
}

}  // namespace caskb200
