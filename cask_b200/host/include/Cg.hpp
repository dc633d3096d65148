// cask::solvers::Cg — the client-API solver object the reference leaves as a stub
// (src/runtime/Cg.hpp:4-14: solve() returns its argument). Here it solves A x = v on the GPU.
#ifndef CASK_B200_HOST_CG_HPP
#define CASK_B200_HOST_CG_HPP
#include "SparseLinearSolvers.hpp"

namespace cask {
namespace solvers {

class Cg {
  CsrMatrix lower_;
  bool have_ = false;

 public:
  int iterations = 0;
  bool converged = false;
  void preprocess(cask::SymCsrMatrix& a) { lower_ = a.matrix; have_ = true; }
  Vector solve(Vector& v) {
    if (!have_) throw std::runtime_error("Cg: preprocess a matrix first");
    Vector rhs(v), x(lower_.n);
    converged = sparse_linear_solvers::pcg<>(lower_, rhs.data.data(), x.data.data(), iterations);
    return x;
  }
};

}  // namespace solvers
}  // namespace cask
#endif
