// The reference's device boundary, kept verbatim in shape: three blocking C functions per design
//   run   (SLiC `Spmv`,           src/spmv/src/SpmvDeviceInterface.h:61-73; GeneratedImplSupport.hpp:38-42)
//   write (SLiC `Spmv_dramWrite`, src/spmv/src/SpmvDeviceInterface.h:21-24; GeneratedImplSupport.hpp:44-49)
//   read  (SLiC `Spmv_dramRead`,  src/spmv/src/SpmvDeviceInterface.h:34-37; GeneratedImplSupport.hpp:31-36)
// over a flat byte-addressed memory per "controller" — here a cudaMalloc arena per controller on one
// B200.  The UNMODIFIED reference Spmv::spmv (src/runtime/Spmv.cpp:185-328) lays its partitions out in
// that memory with dramWrite, calls run, and reads y back with dramRead; run executes the
// reference-format kernel of refformat.cu on whatever the host wrote.
#include <cstring>
#include <map>
#include <mutex>

#include "ctx.cuh"

namespace caskb200 {
namespace {

struct Arena {
  uint8_t* base = nullptr;
  int64_t bytes = 0;
};
std::map<int, Arena> g_arena;  // controller -> memory
std::mutex g_mu;
cudaStream_t g_stream = nullptr;
int64_t g_launches = 0;

int controller_of(const int64_t* sizes, int n_ctl) {
  for (int c = 0; c < n_ctl; c++)
    if (sizes[c] != 0) return c;
  return 0;
}

int ensure(int ctl, int64_t end) {
  Arena& a = g_arena[ctl];
  if (a.bytes >= end) return CASK_B200_OK;
  int64_t want = std::max<int64_t>(end, a.bytes * 2);
  want = (want + (1 << 20) - 1) & ~(int64_t)((1 << 20) - 1);
  uint8_t* p = nullptr;
  CB_CUDA(cudaMalloc(&p, want));
  CB_CUDA(cudaMemset(p, 0, want));
  if (a.base) {
    CB_CUDA(cudaMemcpy(p, a.base, a.bytes, cudaMemcpyDeviceToDevice));
    cudaFree(a.base);
  }
  a.base = p;
  a.bytes = want;
  return CASK_B200_OK;
}

}  // namespace
}  // namespace caskb200

using namespace caskb200;

extern "C" {

int cask_b200_legacy_write(int32_t num_controllers, int64_t size_bytes_cpu, const int64_t* size_bytes_memory_ctl,
                           const int64_t* start_bytes_memory_ctl, const uint8_t* instream_fromcpu) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (num_controllers < 1 || !size_bytes_memory_ctl || !start_bytes_memory_ctl || (!instream_fromcpu && size_bytes_cpu))
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "legacy_write: bad arguments");
  const int c = controller_of(size_bytes_memory_ctl, num_controllers);
  const int64_t addr = start_bytes_memory_ctl[c];
  if (addr < 0 || size_bytes_cpu < 0) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "legacy_write: negative address/size");
  CB_TRY(ensure(c, addr + size_bytes_cpu));
  if (size_bytes_cpu) CB_CUDA(cudaMemcpy(g_arena[c].base + addr, instream_fromcpu, size_bytes_cpu, cudaMemcpyHostToDevice));
  return CASK_B200_OK;
}

int cask_b200_legacy_read(int32_t num_controllers, int64_t size_bytes_cpu, const int64_t* size_bytes_memory_ctl,
                          const int64_t* start_bytes_memory_ctl, uint8_t* outstream_tocpu) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (num_controllers < 1 || !size_bytes_memory_ctl || !start_bytes_memory_ctl || (!outstream_tocpu && size_bytes_cpu))
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "legacy_read: bad arguments");
  const int c = controller_of(size_bytes_memory_ctl, num_controllers);
  const int64_t addr = start_bytes_memory_ctl[c];
  CB_TRY(ensure(c, addr + size_bytes_cpu));
  if (size_bytes_cpu) CB_CUDA(cudaMemcpy(outstream_tocpu, g_arena[c].base + addr, size_bytes_cpu, cudaMemcpyDeviceToHost));
  return CASK_B200_OK;
}

// Every array has num_pipes entries (Spmv.cpp:271-284).  nPartitions is the number of column blocks,
// vectorLoadCycles the cache size in doubles; input_width is a build parameter of the design.
int cask_b200_legacy_run(int32_t num_pipes, int32_t num_controllers, int32_t input_width, int64_t nIterations,
                         int64_t nPartitions, int64_t vectorLoadCycles, const int64_t* colPtrStartAddresses,
                         const int32_t* colptrSizes, const int64_t* indptrValuesAddresses,
                         const int32_t* indptrValuesSizes, const int32_t* nrows, const int64_t* outStartAddresses,
                         const int32_t* reductionCycles, const int32_t* totalCycles, const int64_t* vStartAddresses) {
  std::lock_guard<std::mutex> lk(g_mu);
  (void)reductionCycles; (void)totalCycles; (void)indptrValuesSizes;
  if (num_pipes < 1 || num_controllers < 1 || num_pipes % num_controllers != 0 || input_width < 1)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "legacy_run: bad design parameters");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(CASK_B200_ERR_NO_DEVICE, "legacy_run: no CUDA device (no CPU fallback)");
  if (!g_stream) CB_CUDA(cudaStreamCreate(&g_stream));
  const int per_ctl = num_pipes / num_controllers;
  for (int64_t it = 0; it < std::max<int64_t>(nIterations, 1); it++) {
    for (int p = 0; p < num_pipes; p++) {
      const int c = p / per_ctl;
      if (!g_arena.count(c)) return fail(CASK_B200_ERR_RUNTIME, "legacy_run: nothing was written to controller " + std::to_string(c));
      uint8_t* base = g_arena[c].base;
      const int64_t out_bytes = (int64_t)((nrows[p] + 47) / 48 * 48) * 8;  // 384-byte bursts, Spmv.cpp:91-102
      CB_TRY(ensure(c, outStartAddresses[p] + out_bytes));
      base = g_arena[c].base;
      double* y = reinterpret_cast<double*>(base + outStartAddresses[p]);
      CB_CUDA(cudaMemsetAsync(y, 0, out_bytes, g_stream));
      const int64_t x_len = nPartitions * vectorLoadCycles;  // x was padded to a multiple of the cache size
      CB_TRY(refformat_stripe(g_stream, &g_launches, reinterpret_cast<const int32_t*>(base + colPtrStartAddresses[p]),
                              colptrSizes[p] / 4, base + indptrValuesAddresses[p], nrows[p], (int32_t)nPartitions,
                              (int32_t)vectorLoadCycles, input_width, x_len,
                              reinterpret_cast<const double*>(base + vStartAddresses[p]), y));
    }
  }
  CB_CUDA(cudaStreamSynchronize(g_stream));
  return CASK_B200_OK;
}

int cask_b200_legacy_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_arena) cudaFree(kv.second.base);
  g_arena.clear();
  return CASK_B200_OK;
}

int cask_b200_legacy_launch_count(int64_t* count) {
  if (!count) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "legacy_launch_count: null");
  *count = g_launches;
  return CASK_B200_OK;
}

}  // extern "C"
