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

// true if some pushed range intersects local rows [lo, hi)
__device__ __forceinline__ bool push_overlaps(const PushDesc& d, int64_t lo, int64_t hi) {
  bool any = false;
#pragma unroll
  for (int s = 0; s < kMaxPush; s++)
    if (s < d.nsend) any |= (lo < d.hi[s]) & (hi > d.lo[s]);
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
  const unsigned long long want = w.ctrl->push_seq[w.channel];
  for (int q = 0; q < kMaxPeers; q++)
    if (w.peer_mask & (1u << q)) peer_wait_ge(&w.ctrl->halo_flag[w.channel][q], want, &w.ctrl_rw->error);
  asm volatile("fence.proxy.async;" ::: "memory");  // the data is read next by TMA (async proxy)
}

}  // namespace caskb200
