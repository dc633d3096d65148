// Device-wide inclusive scan (reduce-then-scan, three kernels per level) and CTA-level helpers.
// Hand-written so that the partitioner does not depend on CUB/Thrust.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace caskb200 {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Scan across a CTA of kScanThreads threads: returns this thread's inclusive value, its exclusive
// value through `excl` and the CTA total through `total`. `smem` holds kScanThreads/32 items.
// Ends with a barrier, so `smem` may be reused immediately.
template <typename T, typename Op>
__device__ __forceinline__ T cta_scan(T v, Op op, T identity, T* smem, T* excl, T* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v = op(o, v);
  }
  T up = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane == 31) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < (kScanThreads / 32) ? smem[lane] : identity;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w = op(o, w);
    }
    if (lane < (kScanThreads / 32)) smem[lane] = w;
  }
  __syncthreads();
  T prefix = warp ? smem[warp - 1] : identity;
  *total = smem[kScanThreads / 32 - 1];
  *excl = lane ? op(prefix, up) : prefix;
  __syncthreads();
  return op(prefix, v);
}

template <typename T, typename Op>
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const T* __restrict__ in, T* __restrict__ out, T* __restrict__ tile_totals, int64_t n,
                  Op op, T identity) {
  __shared__ T smem[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  T v[kScanItems];
  T run = identity;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = base + i < n ? in[base + i] : identity;
    run = op(run, v[i]);
    v[i] = run;
  }
  T total, excl;
  cta_scan(run, op, identity, smem, &excl, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; i++)
    if (base + i < n) out[base + i] = op(excl, v[i]);
  if (threadIdx.x == 0 && tile_totals) tile_totals[blockIdx.x] = total;
}

template <typename T, typename Op>
__global__ void __launch_bounds__(kScanThreads)
scan_add_offsets_kernel(T* __restrict__ out, const T* __restrict__ tile_prefix, int64_t n, Op op) {
  if (blockIdx.x == 0) return;
  const T off = tile_prefix[blockIdx.x - 1];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
    if (base + i < n) out[base + i] = op(off, out[base + i]);
}

// out may alias in. Returns a cudaError_t; counts launches into *launches if given.
template <typename T, typename Op>
cudaError_t device_inclusive_scan(const T* in, T* out, int64_t n, Op op, T identity, cudaStream_t s,
                                  int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kScanTile - 1) / kScanTile;
  T* totals = nullptr;
  cudaError_t e;
  if (tiles > 1) {
    e = cudaMalloc(&totals, sizeof(T) * tiles);
    if (e != cudaSuccess) return e;
  }
  scan_tiles_kernel<T, Op><<<(unsigned)tiles, kScanThreads, 0, s>>>(in, out, totals, n, op, identity);
  if (launches) ++*launches;
  if (tiles > 1) {
    e = device_inclusive_scan<T, Op>(totals, totals, tiles, op, identity, s, launches);
    if (e != cudaSuccess) { cudaFree(totals); return e; }
    scan_add_offsets_kernel<T, Op><<<(unsigned)tiles, kScanThreads, 0, s>>>(out, totals, n, op);
    if (launches) ++*launches;
    e = cudaStreamSynchronize(s);
    cudaFree(totals);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

struct OpAddI64 { __host__ __device__ int64_t operator()(int64_t a, int64_t b) const { return a + b; } };
struct OpAddI32 { __host__ __device__ int32_t operator()(int32_t a, int32_t b) const { return a + b; } };
struct OpMaxI64 { __host__ __device__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; } };

}  // namespace caskb200
