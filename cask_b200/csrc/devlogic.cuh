// Backend of the "device logic" translation units (ingest.cu, precond.cu): index-heavy data-preparation steps
// written as per-element functors plus a host-side orchestration that only talks to the small interface below.
//
// The product builds this header (CUDA): functors are __device__ code, for_each() is a grid-stride kernel launch,
// sorting is cub::DeviceRadixSort, scans are scan.cuh.  tests/emu/dev_host.hpp implements the SAME interface with
// plain loops so that the orchestration and the functor bodies - the parts where an off-by-one hides - can be run
// against the oracle on a machine without a GPU.  That emulation is test infrastructure: it is compiled only under
// tests/emu/, never into libcask_b200.so, and the library has no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>

#include "ctx.cuh"
#include "scan.cuh"

#define CB_DEV __device__ __forceinline__

namespace caskb200 {
namespace dev {

struct Exec {
  cudaStream_t stream = nullptr;
  int64_t* launches = nullptr;
  int sm_count = 148;
};

inline Exec exec_of(cask_b200_ctx* ctx) {
  Exec ex;
  ex.stream = ctx->stream;
  ex.launches = &ctx->launches;
  ex.sm_count = ctx->sm_count;
  return ex;
}

inline int alloc(void** p, size_t bytes) {
  *p = nullptr;
  CB_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  return CASK_B200_OK;
}
inline void release(void* p) {
  if (p) cudaFree(p);
}
// host -> device; returns after the copy has completed (the host buffer may be a temporary)
inline int upload(Exec& ex, void* d, const void* h, size_t bytes) {
  if (!bytes) return CASK_B200_OK;
  CB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ex.stream));
  CB_CUDA(cudaStreamSynchronize(ex.stream));
  return CASK_B200_OK;
}
inline int download(Exec& ex, void* h, const void* d, size_t bytes) {
  if (!bytes) return CASK_B200_OK;
  CB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ex.stream));
  CB_CUDA(cudaStreamSynchronize(ex.stream));
  return CASK_B200_OK;
}
inline int copy(Exec& ex, void* d, const void* s, size_t bytes) {
  if (bytes) CB_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, ex.stream));
  return CASK_B200_OK;
}
inline int zero(Exec& ex, void* d, size_t bytes) {
  if (bytes) CB_CUDA(cudaMemsetAsync(d, 0, bytes, ex.stream));
  return CASK_B200_OK;
}
inline int sync(Exec& ex) {
  CB_CUDA(cudaStreamSynchronize(ex.stream));
  return CASK_B200_OK;
}

template <class F>
__global__ void __launch_bounds__(256) for_each_kernel(int64_t n, F f) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

// f(i) for every i in [0, n): one thread per index, grid-stride over at most 8 CTAs of 256 threads per SM
template <class F>
int for_each(Exec& ex, int64_t n, const F& f) {
  if (n <= 0) return CASK_B200_OK;
  const int64_t ctas = (n + 255) / 256;
  const int grid = (int)(ctas < (int64_t)ex.sm_count * 8 ? ctas : (int64_t)ex.sm_count * 8);
  for_each_kernel<F><<<grid, 256, 0, ex.stream>>>(n, f);
  if (ex.launches) ++*ex.launches;
  CB_CUDA(cudaGetLastError());
  return CASK_B200_OK;
}

// stable ascending sort of (key, value) pairs; the result lands in keys_out / vals_out
inline int sort_pairs_u64_u32(Exec& ex, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                              uint32_t* vals_out, int64_t n, int end_bit) {
  if (n <= 0) return CASK_B200_OK;
  if (n > INT32_MAX) return fail(CASK_B200_ERR_UNSUPPORTED, "sort: more than 2^31-1 entries");
  size_t temp_bytes = 0;
  CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit,
                                          ex.stream));
  void* temp = nullptr;
  CB_CUDA(cudaMalloc(&temp, temp_bytes ? temp_bytes : 16));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit,
                                                  ex.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ex.stream);
  cudaFree(temp);
  if (ex.launches) *ex.launches += 8;  // onesweep: histogram + one pass per 8-bit digit (approximate, library kernels)
  if (e != cudaSuccess) return fail(CASK_B200_ERR_CUDA, std::string("radix sort: ") + cudaGetErrorString(e));
  return CASK_B200_OK;
}

// out[i] = in[0] + ... + in[i]; out may alias in
inline int inclusive_sum_i32(Exec& ex, const int32_t* in, int32_t* out, int64_t n) {
  const cudaError_t e = device_inclusive_scan<int32_t, OpAddI32>(in, out, n, OpAddI32(), 0, ex.stream, ex.launches);
  if (e != cudaSuccess) return fail(CASK_B200_ERR_CUDA, std::string("scan: ") + cudaGetErrorString(e));
  return CASK_B200_OK;
}

// IEEE multiply / subtract that the compiler may not contract into an FMA (the oracle is built with -ffp-contract=off)
CB_DEV double mul_rn(double a, double b) { return __dmul_rn(a, b); }
CB_DEV double sub_rn(double a, double b) { return __dsub_rn(a, b); }
// load of an entry another SM may have written earlier in the SAME kernel (persistent level loop of the triangular solves):
// served by L2, never by a stale L1 line
CB_DEV double ld_l2(const double* p) { return __ldcg(p); }

CB_DEV void atomic_or_i32(int32_t* p, int32_t v) { atomicOr(reinterpret_cast<int*>(p), (int)v); }
CB_DEV void atomic_add_i64(int64_t* p, int64_t v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
CB_DEV void atomic_min_i64(int64_t* p, int64_t v) { atomicMin(reinterpret_cast<long long*>(p), (long long)v); }

}  // namespace dev
}  // namespace caskb200
