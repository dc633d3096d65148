// cask::CaskContext — the client API (include/Cask.hpp:14-37), source-compatible and corrected: the
// shipped header passes a GeneratedSpmvImplementation* where a value is expected (l.25/30) and getCg has
// no return (l.33-35) — SURVEY.md section 0.5.
#ifndef CASK_B200_HOST_CASK_HPP
#define CASK_B200_HOST_CASK_HPP
#include "Cg.hpp"
#include "GeneratedImplSupport.hpp"
#include "SparseMatrix.hpp"
#include "Spmv.hpp"

namespace cask {

class CaskContext {
  cask::runtime::SpmvImplementationLoader spmvManager;

  cask::spmv::Spmv pick(int rows) {
    auto* impl = spmvManager.architectureWithParams(rows);
    if (!impl) throw std::invalid_argument("No loaded SpMV design supports " + std::to_string(rows) + " rows");
    return spmv::Spmv(*impl);
  }

 public:
  void preprocess(const SymCsrMatrix&) {}
  cask::spmv::Spmv getSpmv(SymCsrMatrix& matrix) { return pick(matrix.n); }
  cask::spmv::Spmv getSpmv(CsrMatrix& matrix) { return pick(matrix.n); }
  cask::solvers::Cg getCg(SymCsrMatrix& matrix) {
    cask::solvers::Cg cg;
    cg.preprocess(matrix);
    return cg;
  }
};

}  // namespace cask
#endif
