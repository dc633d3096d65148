// cask::sparse_linear_solvers — the reference's pcg<> entry point (src/runtime/SparseLinearSolvers.hpp:
// 162-239) with its preconditioners (:64-156) and BiCGStab (Eigen::BiCGSTAB via solveBICG, src/runtime/SparseLinearSolvers.cpp:18-26),
// executed on the GPU through cask_b200_cg / cask_b200_bicgstab.
#ifndef CASK_B200_HOST_SPARSELINEARSOLVERS_HPP
#define CASK_B200_HOST_SPARSELINEARSOLVERS_HPP
#include <memory>
#include <type_traits>
#include <vector>

#include "SparseMatrix.hpp"
#include "Spmv.hpp"
#include "Utils.hpp"

namespace cask {
namespace sparse_linear_solvers {

namespace detail {
inline cask_b200_design solver_design() {
  cask_b200_design d;
  d.num_pipes = 1; d.cache_size = 8192; d.input_width = 16; d.max_rows = 0; d.num_controllers = 1;
  d.dram_reduction_enabled = 0; d.arch = CASK_B200_ARCH_SIMPLE;
  return d;
}
struct CsrDeleter {
  void operator()(cask_b200_csr* c) const { cask_b200_csr_free(c); }
};
// the arrays of `a` as a device-resident matrix (the preconditioner is built from exactly what the caller passed)
inline std::shared_ptr<cask_b200_csr> device_copy(cask_b200_ctx* ctx, const CsrMatrix& a) {
  std::vector<int32_t> rows(a.values.size());
  for (int i = 0; i < a.n; i++)
    for (int k = a.row_ptr[i]; k < a.row_ptr[i + 1]; k++) rows[k] = i;
  cask_b200_csr* c = nullptr;
  spmv::detail::throw_on(cask_b200_ingest_coo(ctx, a.n, a.m, (int64_t)a.values.size(), rows.data(), a.col_ind.data(),
                                              a.values.data(), 0, &c));
  return std::shared_ptr<cask_b200_csr>(c, CsrDeleter());
}
}  // namespace detail

class IdentityPreconditioner {  // SparseLinearSolvers.hpp:64-73
 public:
  static constexpr int kind = CASK_B200_PRECON_IDENTITY;
  IdentityPreconditioner(const CsrMatrix&) {}
  virtual std::vector<double> apply(const std::vector<double>& x) { return x; }
  virtual ~IdentityPreconditioner() {}
};

// z = r / a_ii (1 where a_ii is absent or 0).  Not in the reference; SURVEY.md 8(f) rank 4.
class JacobiPreconditioner {
  std::vector<double> invd;

 public:
  static constexpr int kind = CASK_B200_PRECON_JACOBI;
  JacobiPreconditioner(const CsrMatrix& a) : invd(a.n, 1.0) {
    for (int i = 0; i < a.n; i++)
      for (int k = a.row_ptr[i]; k < a.row_ptr[i + 1]; k++)
        if (a.col_ind[k] == i && a.values[k] != 0) invd[i] = 1.0 / a.values[k];
  }
  virtual std::vector<double> apply(const std::vector<double>& x) {
    std::vector<double> z(x.size());
    for (size_t i = 0; i < x.size(); i++) z[i] = invd[i] * x[i];
    return z;
  }
  virtual ~JacobiPreconditioner() {}
};

// ILUPreconditioner, SparseLinearSolvers.hpp:77-156, with the factorisation and the triangular solves on the GPU
// (level-scheduled; arithmetic and order inside a row as in the reference).  `pc` is filled like the reference's
// member so that test/LinearSolvers.cpp:79-123 reads the same.  kUnitLower = true is the textbook application of the
// same factors (unit lower solve) - the reference's own apply() divides the lower solve by U's diagonal.
template <bool kUnitLower>
class IluPreconditionerT {
  std::shared_ptr<cask_b200_ctx> ctx;
  std::shared_ptr<cask_b200_csr> dev;

 public:
  static constexpr int kind = kUnitLower ? CASK_B200_PRECON_ILU_UNIT : CASK_B200_PRECON_ILU;
  DokMatrix pc;
  int levels_lower = 0, levels_upper = 0;
  IluPreconditionerT(const CsrMatrix& a) : ctx(spmv::detail::make_ctx()), pc(a.n, a.m, a.nnzs) {
    if (!a.isSymmetric()) throw std::invalid_argument("ILUPreconditioner only supports symmetric CSR matrices");  // :90-91
    const cask_b200_design d = detail::solver_design();
    spmv::detail::throw_on(cask_b200_preprocess(ctx.get(), &d, a.n, a.m, (int64_t)a.values.size(), a.row_ptr.data(),
                                                a.col_ind.data(), a.values.data()));
    std::vector<double> f(a.values.size() ? a.values.size() : 1);
    int32_t ll = 0, lu = 0;
    spmv::detail::throw_on(cask_b200_ilu_factor(ctx.get(), f.data(), &ll, &lu));
    levels_lower = ll; levels_upper = lu;
    for (int i = 0; i < a.n; i++)
      for (int k = a.row_ptr[i]; k < a.row_ptr[i + 1]; k++) pc.dok[i][a.col_ind[k]] = f[k];
  }
  virtual std::vector<double> apply(const std::vector<double>& x) {  // :142-150
    std::vector<double> z(x.size());
    int32_t zero_pivot = 0;
    spmv::detail::throw_on(cask_b200_ilu_apply(ctx.get(), kUnitLower ? 1 : 0, x.data(), z.data(), &zero_pivot));
    return z;
  }
  void pretty_print() { pc.pretty_print(); }
  virtual ~IluPreconditionerT() {}
};
using ILUPreconditioner = IluPreconditionerT<false>;
using IluUnitPreconditioner = IluPreconditionerT<true>;

namespace detail {
// mkl_dcsrsymv('l') consumes the lower triangle of a symmetric matrix; the GPU kernels take the full
// matrix, so mirror the strict lower triangle once (DokMatrix::explicitSymmetric semantics: the diagonal
// is not doubled; entries above the diagonal in the input are ignored, as MKL ignores them).
inline CsrMatrix expand_lower(const CsrMatrix& a) {
  const int n = a.n;
  std::vector<int> cnt(n + 1, 0);
  for (int i = 0; i < n; i++)
    for (int k = a.row_ptr[i]; k < a.row_ptr[i + 1]; k++) {
      const int j = a.col_ind[k];
      if (j > i) continue;
      cnt[i + 1]++;
      if (j != i) cnt[j + 1]++;
    }
  for (int i = 0; i < n; i++) cnt[i + 1] += cnt[i];
  std::vector<int> rp(cnt), cur(cnt.begin(), cnt.end() - 1), ci(cnt[n]);
  std::vector<double> va(cnt[n]);
  // row i receives its own lower entries (columns < = i) first, then mirrored entries (columns > i) in
  // increasing column order because the source rows are visited in increasing order
  for (int i = 0; i < n; i++)
    for (int k = a.row_ptr[i]; k < a.row_ptr[i + 1]; k++)
      if (a.col_ind[k] <= i) { ci[cur[i]] = a.col_ind[k]; va[cur[i]++] = a.values[k]; }
  for (int i = 0; i < n; i++)
    for (int k = a.row_ptr[i]; k < a.row_ptr[i + 1]; k++) {
      const int j = a.col_ind[k];
      if (j < i) { ci[cur[j]] = i; va[cur[j]++] = a.values[k]; }
    }
  return CsrMatrix(n, a.m, cnt[n], va, ci, rp);
}
}  // namespace detail

// pcg: `a` is the LOWER TRIANGLE of the symmetric system, 0-based CSR, exactly what readSymMatrix
// yields (.matrix); x holds the initial guess on entry and the solution on return; `iterations` keeps the
// reference's meaning (SparseLinearSolvers.hpp:231). maxiters = 2000, tol = 1e-5 as in :166-167.
template <typename T = double, typename Precon = IdentityPreconditioner>
bool pcg(const CsrMatrix& a, double* rhs, double* x, int& iterations, bool verbose = false,
         cask::utils::Timer* t = nullptr) {
  static_assert(std::is_same<T, double>::value, "pcg runs in fp64 like the reference");
  if (t) t->tic("cg:setup");
  auto ctx = spmv::detail::make_ctx();
  const CsrMatrix full = detail::expand_lower(a);
  const cask_b200_design d = detail::solver_design();
  spmv::detail::throw_on(cask_b200_preprocess(ctx.get(), &d, full.n, full.m, (int64_t)full.values.size(),
                                              full.row_ptr.data(), full.col_ind.data(), full.values.data()));
  if (t) { t->toc("cg:setup"); t->tic("cg:solve"); }
  int32_t it = iterations, converged = 0;
  double rs = 0;
  if (Precon::kind == CASK_B200_PRECON_IDENTITY) {
    spmv::detail::throw_on(cask_b200_cg(ctx.get(), rhs, x, 2000, 1E-5, &it, &converged, &rs));
  } else {
    // Precon precon{a} (:171): the reference builds the preconditioner from the arrays it was HANDED - the stored lower
    // triangle - while the product runs on the implied symmetric matrix
    std::shared_ptr<cask_b200_csr> pm = detail::device_copy(ctx.get(), a);
    spmv::detail::throw_on(cask_b200_precond_set_matrix(ctx.get(), pm.get()));
    const int rc = cask_b200_pcg(ctx.get(), rhs, x, 2000, 1E-5, Precon::kind, &it, &converged, &rs);
    cask_b200_precond_set_matrix(ctx.get(), nullptr);
    spmv::detail::throw_on(rc);
  }
  iterations = it;
  if (verbose) std::cout << " rsnew " << rs << " iterations " << iterations << "\n";
  if (t) t->toc("cg:solve");
  return converged != 0;
}

// BiCGStab on a general CSR matrix with Eigen's defaults (Jacobi preconditioner, tol = DBL_EPSILON,
// maxIt = 2 n, x0 = 0); returns the solution, reports iterations / relative residual.
inline Vector bicgstab(const CsrMatrix& a, const Vector& b, int* iterations = nullptr, double* error = nullptr,
                       double tol = 0.0, int maxit = 0) {
  auto ctx = spmv::detail::make_ctx();
  const cask_b200_design d = detail::solver_design();
  spmv::detail::throw_on(cask_b200_preprocess(ctx.get(), &d, a.n, a.m, (int64_t)a.values.size(), a.row_ptr.data(),
                                              a.col_ind.data(), a.values.data()));
  Vector x(a.n);
  int32_t it = maxit;
  double te = tol;
  spmv::detail::throw_on(cask_b200_bicgstab(ctx.get(), b.data.data(), x.data.data(), &it, &te));
  if (iterations) *iterations = it;
  if (error) *error = te;
  return x;
}

}  // namespace sparse_linear_solvers
}  // namespace cask
#endif
