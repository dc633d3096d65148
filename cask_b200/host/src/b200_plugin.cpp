// libSpmv_b200.so — the B200 counterpart of the reference's generated plugin libSpmv_<target>.so
// (src/frontend/cask.py:172-283, linked by src/cmake/Utils.cmake:29-43 or LD_PRELOADed by
// src/frontend/mockrunner:1).  It defines the one symbol such a plugin must define, the out-of-line
// constructor cask::runtime::SpmvImplementationLoader::SpmvImplementationLoader()
// (src/runtime/GeneratedImplSupport.hpp:104), and registers B200 "designs".  Each design carries the
// three device callbacks with the exact SLiC signatures (GeneratedImplSupport.hpp:31-49); they forward
// to the C ABI's flat-memory boundary (cask_b200_legacy_write/run/read), so the UNMODIFIED reference
// Spmv::spmv (src/runtime/Spmv.cpp:185-328) runs its partitions on the GPU.
//
// Compiles against either GeneratedImplSupport.hpp: the host mirror in ../include or the reference's own
// (that build lives in oracle/Makefile, target ref_harness).
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "../../../include/cask_b200.h"
#include "GeneratedImplSupport.hpp"

namespace {

struct DesignParams { int max_rows, num_pipes, cache_size, input_width, num_controllers; };
// smallest-first, like a swept set of builds (SURVEY.md section 5 "Config / flags")
const DesignParams kDesigns[] = {
    {20000, 1, 2048, 16, 1},
    {1 << 20, 4, 4096, 16, 2},
    {INT_MAX / 2, 8, 8192, 16, 1},
};

void die_on(int rc, const char* what) {
  if (rc == CASK_B200_OK) return;
  // the SLiC functions return void (src/spmv/src/SpmvDeviceInterface.h:21-73): a device failure cannot be
  // reported through them, and silently returning would hand zeros to the caller like the mock does.
  std::fprintf(stderr, "libSpmv_b200: %s failed: %s\n", what, cask_b200_last_error());
  std::abort();
}

template <int ID>
void run_fn(int64_t nIterations, int64_t nPartitions, int64_t vectorLoadCycles, const int64_t* colPtrStartAddresses,
            const int32_t* colptrSizes, const int64_t* indptrValuesAddresses, const int32_t* indptrValuesSizes,
            const int32_t* nrows, const int64_t* outStartAddresses, const int32_t* reductionCycles,
            const int32_t* totalCycles, const int64_t* vStartAddresses) {
  const DesignParams& d = kDesigns[ID];
  die_on(cask_b200_legacy_run(d.num_pipes, d.num_controllers, d.input_width, nIterations, nPartitions, vectorLoadCycles,
                              colPtrStartAddresses, colptrSizes, indptrValuesAddresses, indptrValuesSizes, nrows,
                              outStartAddresses, reductionCycles, totalCycles, vStartAddresses),
         "run");
}
template <int ID>
void write_fn(const int64_t size_bytes_cpu, const int64_t* size_bytes_memory_ctl, const int64_t* start_bytes_memory_ctl,
              const uint8_t* instream_fromcpu, const char*) {
  die_on(cask_b200_legacy_write(kDesigns[ID].num_controllers, size_bytes_cpu, size_bytes_memory_ctl,
                                start_bytes_memory_ctl, instream_fromcpu), "dramWrite");
}
template <int ID>
void read_fn(const int64_t size_bytes_cpu, const int64_t* size_bytes_memory_ctl, const int64_t* start_bytes_memory_ctl,
             uint8_t* outstream_tocpu, const char*) {
  die_on(cask_b200_legacy_read(kDesigns[ID].num_controllers, size_bytes_cpu, size_bytes_memory_ctl,
                               start_bytes_memory_ctl, outstream_tocpu), "dramRead");
}

template <int ID>
cask::runtime::GeneratedSpmvImplementation* make() {
  const DesignParams& d = kDesigns[ID];
  return new cask::runtime::GeneratedSpmvImplementation(ID, run_fn<ID>, write_fn<ID>, read_fn<ID>, d.max_rows, d.num_pipes,
                                                        d.cache_size, d.input_width, /*dram_reduction_enabled=*/0,
                                                        d.num_controllers);
}

}  // namespace

cask::runtime::SpmvImplementationLoader::SpmvImplementationLoader() {
  this->impls.push_back(make<0>());
  this->impls.push_back(make<1>());
  this->impls.push_back(make<2>());
}
