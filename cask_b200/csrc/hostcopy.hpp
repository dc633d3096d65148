// Host-side helpers of the host-buffer SpMV call when the caller's vectors are PAGEABLE memory (cask::Vector is a
// std::vector<double>: what every caller of the reference's Spmv::spmv(const Vector&) holds, src/runtime/Spmv.cpp:185).
// cudaMemcpyAsync on pageable memory is staged by the driver through its own bounce buffer by ONE thread (measured:
// 13 GB/s for both directions together on C2, 20 ms per SpMV against 3.3 ms with pinned buffers).  The library stages
// such vectors itself: a ring of pinned chunks per direction, filled / drained by several host threads at memory
// bandwidth while the DMA engines move the previous chunks (capi.cu: spmv_host_pipelined).
//
//   HostCopyPool   T worker threads; copy() splits one memcpy into pieces, the caller takes a piece too and returns
//                  when all pieces are done.  Safe to call from two threads at once (upload side and drain side).
//   DrainQueue     hand-over of downloaded chunks from the thread that issues the D2H copies to the thread that copies
//                  them out of the pinned ring into the caller's y.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace caskb200 {

class HostCopyPool {
 public:
  explicit HostCopyPool(int threads) {
    for (int t = 0; t < threads; t++) workers_.emplace_back([this] { run(); });
  }
  ~HostCopyPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  HostCopyPool(const HostCopyPool&) = delete;
  HostCopyPool& operator=(const HostCopyPool&) = delete;

  int threads() const { return (int)workers_.size(); }

  void copy(void* dst, const void* src, size_t bytes) {
    constexpr size_t kMinPiece = 256u << 10;
    const size_t parts = std::max<size_t>(1, std::min<size_t>(workers_.size() + 1, bytes / kMinPiece));
    if (parts == 1) {
      std::memcpy(dst, src, bytes);
      return;
    }
    Batch b;
    b.pending = (int)parts - 1;
    const size_t piece = ((bytes + parts - 1) / parts + 63) & ~size_t(63);
    {
      std::lock_guard<std::mutex> g(mu_);
      for (size_t i = 1; i < parts; i++) {
        const size_t off = std::min(bytes, i * piece), end = std::min(bytes, (i + 1) * piece);
        jobs_.push_back(Job{(char*)dst + off, (const char*)src + off, end - off, &b});
      }
    }
    cv_.notify_all();
    std::memcpy(dst, src, std::min(bytes, piece));  // the caller's own piece
    std::unique_lock<std::mutex> g(b.mu);
    b.cv.wait(g, [&] { return b.pending == 0; });
  }

 private:
  struct Batch {
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
  };
  struct Job {
    char* dst;
    const char* src;
    size_t bytes;
    Batch* batch;
  };
  void run() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return stop_ || !jobs_.empty(); });
        if (jobs_.empty()) return;  // stop requested and nothing left
        j = jobs_.front();
        jobs_.pop_front();
      }
      if (j.bytes) std::memcpy(j.dst, j.src, j.bytes);
      {
        std::lock_guard<std::mutex> g(j.batch->mu);
        j.batch->pending--;
        if (j.batch->pending == 0) j.batch->cv.notify_all();  // under the lock: the waiter cannot destroy the batch before
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Job> jobs_;
  bool stop_ = false;
};

// One downloaded chunk: bytes sit in pinned slot `src` once `ready` has completed; they belong at `dst`.
struct DrainItem {
  void* dst = nullptr;
  const void* src = nullptr;
  size_t bytes = 0;
  cudaEvent_t ready = nullptr;
};

class DrainQueue {
 public:
  // consumer thread: copies chunks out in order until close(); the first CUDA error stops it and is kept
  void run(HostCopyPool* pool) {
    for (;;) {
      DrainItem it;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return closed_ || !items_.empty(); });
        if (items_.empty()) return;
        it = items_.front();
        items_.pop_front();
      }
      if (error_.load() == cudaSuccess) {
        const cudaError_t e = cudaEventSynchronize(it.ready);
        if (e != cudaSuccess) error_.store(e);
        else pool->copy(it.dst, it.src, it.bytes);
      }
      {
        std::lock_guard<std::mutex> g(mu_);
        drained_++;
      }
      slot_cv_.notify_all();
    }
  }
  void push(const DrainItem& it) {
    {
      std::lock_guard<std::mutex> g(mu_);
      items_.push_back(it);
      pushed_++;
    }
    cv_.notify_all();
  }
  // producer: blocks until at most `in_flight` pushed chunks are still undrained (ring slot reuse)
  void wait_in_flight(int64_t in_flight) {
    std::unique_lock<std::mutex> g(mu_);
    slot_cv_.wait(g, [&] { return pushed_ - drained_ <= in_flight; });
  }
  void close() {
    {
      std::lock_guard<std::mutex> g(mu_);
      closed_ = true;
    }
    cv_.notify_all();
  }
  cudaError_t error() const { return error_.load(); }

 private:
  std::mutex mu_;
  std::condition_variable cv_, slot_cv_;
  std::deque<DrainItem> items_;
  int64_t pushed_ = 0, drained_ = 0;
  bool closed_ = false;
  std::atomic<cudaError_t> error_{cudaSuccess};
};

}  // namespace caskb200
