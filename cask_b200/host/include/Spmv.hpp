// cask::spmv::Spmv — the reference's SpMV architecture interface (src/runtime/Spmv.hpp:49-202,
// src/runtime/Spmv.cpp), same public members, with the device side on a B200:
//   preprocess(const CsrMatrix&)  -> cask_b200_preprocess   (GPU partitioner)
//   spmv(const Vector&)           -> cask_b200_spmv         (staged-ELL / CSR kernels)
//   do_blocking(...)              -> cask_b200_partition_*  (reference-format arrays, produced on the GPU)
// Errors come back from the C ABI as codes and are rethrown as the reference's exception types with the
// reference's messages.  The north_star's older spellings are kept as aliases at the bottom.
#ifndef CASK_B200_HOST_SPMV_HPP
#define CASK_B200_HOST_SPMV_HPP
#include <algorithm>
#include <chrono>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>

#include "../../../include/cask_b200.h"
#include "GeneratedImplSupport.hpp"
#include "SparseMatrix.hpp"
#include "Utils.hpp"

namespace cask {
namespace spmv {

#pragma pack(push, 1)
struct indptr_value {  // Spmv.hpp:13-20 — 12 bytes, no padding
  double value;
  int indptr;
  indptr_value(double v, int i) : value(v), indptr(i) {}
  indptr_value() : value(0), indptr(0) {}
};
#pragma pack(pop)
static_assert(sizeof(indptr_value) == 12, "indptr_value must be packed to 12 bytes");

struct Partition {  // Spmv.hpp:25-44
  int nBlocks, n, paddingCycles, totalCycles, vector_load_cycles, outSize;
  int reductionCycles, emptyCycles;
  int m_colptr_unpaddedLength;
  int m_indptr_values_unpaddedLength;
  std::vector<int> m_colptr;
  std::vector<indptr_value> m_indptr_values;
  std::string to_string() const {
    std::stringstream s;
    s << "Vector load cycles " << vector_load_cycles << std::endl << "Padding cycles = " << paddingCycles << std::endl
      << "Total cycles = " << totalCycles << std::endl << "Nrows = " << n << std::endl << "Partitions = " << nBlocks
      << std::endl << "Reduction cycles = " << reductionCycles << std::endl << "Empty cycles = " << emptyCycles << std::endl;
    return s.str();
  }
};

namespace detail {
inline void throw_on(int rc) {
  if (rc == CASK_B200_OK) return;
  const std::string msg = cask_b200_last_error();
  if (rc == CASK_B200_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);  // CASK_B200_ERR_RUNTIME and every device-side failure
}
struct CtxDeleter {
  void operator()(cask_b200_ctx* c) const { cask_b200_destroy(c); }
};
inline std::shared_ptr<cask_b200_ctx> make_ctx(int device = 0) {
  cask_b200_ctx* c = nullptr;
  throw_on(cask_b200_create(&c, device));
  return std::shared_ptr<cask_b200_ctx>(c, CtxDeleter());
}
}  // namespace detail

class Spmv {
 protected:
  CsrMatrix mat;
  std::shared_ptr<cask_b200_ctx> ctx;  // shared so that Spmv stays copyable like the reference's (Cask.hpp returns by value)
  bool preprocessed = false;

  virtual int arch() const { return CASK_B200_ARCH_SIMPLE; }
  cask_b200_design design() const {
    cask_b200_design d;
    d.num_pipes = impl.num_pipes; d.cache_size = impl.cache_size; d.input_width = impl.input_width;
    d.max_rows = impl.max_rows; d.num_controllers = impl.num_controllers;
    d.dram_reduction_enabled = impl.dram_reduction_enabled; d.arch = arch();
    return d;
  }
  void need_ctx() { if (!ctx) ctx = detail::make_ctx(); }
  static Partition fetch_partition(cask_b200_ctx* c, int p) {
    cask_b200_partition_info info;
    detail::throw_on(cask_b200_partition_get_info(c, p, &info));
    Partition q;
    q.nBlocks = info.nBlocks; q.n = info.n; q.paddingCycles = info.paddingCycles; q.totalCycles = info.totalCycles;
    q.vector_load_cycles = info.vector_load_cycles; q.outSize = info.outSize; q.reductionCycles = info.reductionCycles;
    q.emptyCycles = info.emptyCycles; q.m_colptr_unpaddedLength = info.m_colptr_unpaddedLength;
    q.m_indptr_values_unpaddedLength = info.m_indptr_values_unpaddedLength;
    q.m_colptr.resize(info.len_colptr);
    q.m_indptr_values.resize(info.len_pairs);
    detail::throw_on(cask_b200_partition_export(c, p, q.m_colptr.data(), q.m_indptr_values.data()));
    return q;
  }

 public:
  runtime::GeneratedSpmvImplementation impl;

  // Spmv.hpp:56-66: parameters only (the reference wires its mock callbacks here; this one computes)
  Spmv(int _cacheSize, int _inputWidth, int _numPipes, int _maxRows, int _numControllers)
      : impl(-1, runtime::spmvRunMock, runtime::spmvWriteMock, runtime::spmvReadMock, _maxRows, _numPipes, _cacheSize,
             _inputWidth, false, _numControllers) {}
  // Spmv.hpp:71: from a loaded design
  Spmv(runtime::GeneratedSpmvImplementation _impl) : impl(_impl) {}
  virtual ~Spmv() {}

  virtual std::string get_name() { return "Simple"; }
  bool operator==(const Spmv& o) const { return impl == o.impl; }
  double getFrequency() { return 200.0 * 1E6; }  // the FPGA model clock, kept for the cycle figures (Spmv.hpp:91-93)
  virtual double getGFlopsCount() { return 2 * this->mat.nnzs / 1E9; }
  virtual bool isValid() { return impl.num_pipes >= impl.num_controllers; }

  // Spmv::preprocess, Spmv.cpp:329-365
  void preprocess(const CsrMatrix& m) {
    need_ctx();
    mat = m;
    const cask_b200_design d = design();
    detail::throw_on(cask_b200_preprocess(ctx.get(), &d, m.n, m.m, (int64_t)m.values.size(), m.row_ptr.data(),
                                          m.col_ind.data(), m.values.data()));
    preprocessed = true;
  }

  // Spmv::spmv, Spmv.cpp:185-328 (argument checks happen behind the ABI with the same messages).  Prints the complete
  // set of "Result  <key>=" lines of Spmv.cpp:266-301 in the reference's order, so the frontend's scrapers
  // (src/frontend/cask.py:360-368) keep working: the three per-partition cycle lists come from the reference-format
  // partitions the GPU partitioner emits (empty lists where that format cannot exist, rows x blocks > INT32_MAX), the
  // estimates follow the reference's formulas on its 200 MHz model clock, "Iterations" is 1 (the reference's device call
  // runs the product twice and halves the time), "Took (ms)" and "Gflops (actual)" are measured around the ABI call.
  Vector spmv(const Vector& x) {
    if (!preprocessed) throw std::runtime_error("numPipes should equal numPartitions");  // Spmv.cpp:222-224: nothing preprocessed
    if (x.size() < mat.m) throw std::invalid_argument("spmv: x has fewer entries than the matrix has columns");
    Vector y(mat.n);
    std::vector<int> totalCycles, paddingCycles, reductionCycles;
    if (log_results) {
      for (int p = 0; p < impl.num_pipes; p++) {
        cask_b200_partition_info info;
        if (cask_b200_partition_get_info(ctx.get(), p, &info) != CASK_B200_OK) {
          totalCycles.clear(); paddingCycles.clear(); reductionCycles.clear();
          break;
        }
        totalCycles.push_back(info.totalCycles);
        paddingCycles.push_back(info.paddingCycles);
        reductionCycles.push_back(info.reductionCycles);
      }
      std::cout << "Running on B200" << std::endl;  // Spmv.cpp:262 prints "Running on DFE" here (not scraped)
      utils::logResult("Total cycles", totalCycles);
      utils::logResult("Padding cycles", paddingCycles);
      utils::logResult("Reduction cycles", reductionCycles);
    }
    const int nIterations = 1;
    const auto t0 = std::chrono::high_resolution_clock::now();
    detail::throw_on(cask_b200_spmv(ctx.get(), x.data.data(), y.data.data()));
    const double took = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count() / nIterations;
    if (log_results) {
      const double maxCycles = totalCycles.empty() ? 0.0 : (double)*std::max_element(totalCycles.begin(), totalCycles.end());
      const double bwidthEst = impl.num_pipes * impl.input_width * getFrequency() * 12 / 1E9;  // Spmv.cpp:285-289
      const double scalingFactor = std::max(bwidthEst / 65.0, 1.0);
      const double est = maxCycles / getFrequency();
      const double gflopsEst = est > 0 ? (2.0 * (double)mat.nnzs / (est * scalingFactor)) / 1E9 : 0.0;
      utils::logResult("Input width ", impl.input_width);
      utils::logResult("Pipes ", impl.num_pipes);
      utils::logResult("Iterations", nIterations);
      utils::logResult("Took (ms)", took * 1e3);
      utils::logResult("Est (ms)", est);
      utils::logResult("Gflops (est)", gflopsEst);
      utils::logResult("Gflops (actual)", 2.0 * (double)mat.nnzs / took / 1e9);
      utils::logResult("BWidth (est)", bwidthEst);
    }
    return y;
  }
  bool log_results = true;  // solver loops that call spmv() every iteration switch the Result lines off

  // Spmv::do_blocking, Spmv.cpp:42-107: the reference-format partition of a (stripe) matrix
  Partition do_blocking(const CsrMatrix& m, int blockSize, int inputWidth) {
    auto tmp = detail::make_ctx();
    cask_b200_design d = design();
    d.num_pipes = 1; d.cache_size = blockSize; d.input_width = inputWidth;
    detail::throw_on(cask_b200_preprocess(tmp.get(), &d, m.n, m.m, (int64_t)m.values.size(), m.row_ptr.data(),
                                          m.col_ind.data(), m.values.data()));
    return fetch_partition(tmp.get(), 0);
  }
  // the Partition list the reference keeps privately (Spmv.hpp:51)
  std::vector<Partition> getPartitions() {
    if (!preprocessed) throw std::runtime_error("Matrix not defined - run preprocess on the matrix");
    std::vector<Partition> r;
    for (int p = 0; p < impl.num_pipes; p++) r.push_back(fetch_partition(ctx.get(), p));
    return r;
  }
  // max totalCycles over the partitions (Spmv.hpp:103-109); the FPGA resource model of Model.hpp is out of scope
  virtual double getEstimatedClockCycles() {
    if (!preprocessed) throw std::runtime_error("Matrix not defined - run preprocess on the matrix");
    int best = 0;
    for (int p = 0; p < impl.num_pipes; p++) {
      cask_b200_partition_info info;
      detail::throw_on(cask_b200_partition_get_info(ctx.get(), p, &info));
      best = std::max(best, (int)info.totalCycles);
    }
    return best;
  }
  cask_b200_plan_stats getPlanStats() {
    cask_b200_plan_stats s;
    detail::throw_on(cask_b200_plan_get_stats(ctx.get(), &s));
    return s;
  }
  cask_b200_ctx* handle() { need_ctx(); return ctx.get(); }

  // north_star spelling of the same call
  Vector dfespmv(const Vector& x) { return spmv(x); }
};

// Spmv.hpp:211-259: middle column blocks run-length encode their empty rows
class SkipEmptyRowsSpmv : public Spmv {
 protected:
  int arch() const override { return CASK_B200_ARCH_SKIPEMPTY; }

 public:
  SkipEmptyRowsSpmv(int _cacheSize, int _inputWidth, int _numPipes, int _maxRows, int _numControllers)
      : Spmv(_cacheSize, _inputWidth, _numPipes, _maxRows, _numControllers) {}
  SkipEmptyRowsSpmv(runtime::GeneratedSpmvImplementation _impl) : Spmv(_impl) {}
  std::string get_name() override { return "SkipEmpty"; }
};

// names used by BASELINE.json's north_star (an older CASK revision)
using SpmvArchitecture = Spmv;
using SimpleSpmvArchitecture = Spmv;
using SkipEmptyRowsSpmvArchitecture = SkipEmptyRowsSpmv;

}  // namespace spmv
}  // namespace cask
#endif
