// The plugin boundary of the reference (src/runtime/GeneratedImplSupport.hpp:26-127), same names and
// members: a design = its architecture parameters + three device callbacks (run / write / read), and
// the loader whose out-of-line constructor a plugin library defines.  Here the plugin is
// libSpmv_b200.so (src/b200_plugin.cpp), which registers B200 designs whose callbacks drive the GPU
// through the C ABI (cask_b200_legacy_run/write/read).
#ifndef CASK_B200_HOST_GENERATEDIMPLSUPPORT_HPP
#define CASK_B200_HOST_GENERATEDIMPLSUPPORT_HPP
#include <cstdint>
#include <functional>
#include <vector>

namespace cask {
namespace runtime {

// No-op device functions: the reference's mock flow (GeneratedImplSupport.hpp:31-49). A design built on
// these computes nothing; cask::spmv::Spmv below never calls them — it owns a GPU context instead.
inline void spmvReadMock(const int64_t, const int64_t*, const int64_t*, uint8_t*, const char*) {}
inline void spmvRunMock(int64_t, int64_t, int64_t, const int64_t*, const int32_t*, const int64_t*, const int32_t*,
                        const int32_t*, const int64_t*, const int32_t*, const int32_t*, const int64_t*) {}
inline void spmvWriteMock(const int64_t, const int64_t*, const int64_t*, const uint8_t*, const char*) {}

class GeneratedSpmvImplementation {
  using SpmvFunctionT = decltype(spmvRunMock);
  using SpmvDramWriteFunctionT = decltype(spmvWriteMock);
  using SpmvDramReadFunctionT = decltype(spmvReadMock);

 public:
  const int id, max_rows, num_pipes, cache_size, input_width, dram_reduction_enabled, num_controllers;
  std::function<SpmvFunctionT> Spmv;
  std::function<SpmvDramWriteFunctionT> write;
  std::function<SpmvDramReadFunctionT> read;

  GeneratedSpmvImplementation(int _id, SpmvFunctionT run, SpmvDramWriteFunctionT dramWrite,
                              SpmvDramReadFunctionT dramRead, int _max_rows, int _num_pipes, int _cache_size,
                              int _input_width, int _dram_reduction_enabled, int _num_controllers)
      : id(_id), max_rows(_max_rows), num_pipes(_num_pipes), cache_size(_cache_size), input_width(_input_width),
        dram_reduction_enabled(_dram_reduction_enabled), num_controllers(_num_controllers), Spmv(run),
        write(dramWrite), read(dramRead) {}

  bool operator==(const GeneratedSpmvImplementation& o) const {
    return max_rows == o.max_rows && num_pipes == o.num_pipes && cache_size == o.cache_size &&
           input_width == o.input_width && dram_reduction_enabled == o.dram_reduction_enabled &&
           num_controllers == o.num_controllers;
  }
};

class SpmvImplementationLoader {
  std::vector<GeneratedSpmvImplementation*> impls;

 public:
  SpmvImplementationLoader();  // defined by the plugin library (cask.py:255-283 in the reference)

  // smallest design whose max_rows covers the matrix; nullptr if none does (:110-118)
  GeneratedSpmvImplementation* architectureWithParams(int maxRows) {
    GeneratedSpmvImplementation* best = nullptr;
    for (auto* a : impls)
      if (a->max_rows >= maxRows && (!best || a->max_rows < best->max_rows)) best = a;
    return best;
  }
  GeneratedSpmvImplementation* architectureWithId(int id) { return impls.at(id); }  // std::out_of_range like the reference
};

}  // namespace runtime
}  // namespace cask
#endif
