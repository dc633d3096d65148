// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// C-callable driver around the reference's OWN solver code, compiled where it lies: pcg<T, Precon>,
// IdentityPreconditioner and ILUPreconditioner (src/runtime/SparseLinearSolvers.hpp:62-239) with -DUSEMKL and
// the six MKL routines they call supplied by ref_shim/mkl.h (MKL itself is absent).  Part of
// oracle/_ref/libcaskref.so (oracle/Makefile, target `ref`).  No reference source is copied into the repository.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "SparseLinearSolvers.hpp"

using cask::CsrMatrix;
namespace sls = cask::sparse_linear_solvers;

namespace {
thread_local std::string g_serr;
CsrMatrix make(int n, int m, int nnz, const int* rp, const int* ci, const double* va) {
  return CsrMatrix(n, m, nnz, std::vector<double>(va, va + nnz), std::vector<int>(ci, ci + nnz),
                   std::vector<int>(rp, rp + n + 1));
}
}  // namespace

extern "C" {

const char* ref_solvers_last_error() { return g_serr.c_str(); }

// pcg<double, Precon>(a, rhs, x, iterations): precon 0 = IdentityPreconditioner, 1 = ILUPreconditioner.
// `a` is handed over exactly as given (the reference's tests pass the stored lower triangle, readSymMatrix().matrix).
// Returns 1 converged / 0 not / -1 exception.
int ref_pcg(int n, int nnz, const int* rp, const int* ci, const double* va, const double* rhs, double* x,
            int* iterations, int precon) {
  try {
    CsrMatrix a = make(n, n, nnz, rp, ci, va);
    std::vector<double> b(rhs, rhs + n);
    const bool ok = precon == 1 ? sls::pcg<double, sls::ILUPreconditioner>(a, b.data(), x, *iterations)
                                : sls::pcg<double, sls::IdentityPreconditioner>(a, b.data(), x, *iterations);
    return ok ? 1 : 0;
  } catch (std::exception& e) { g_serr = e.what(); return -1; }
}

// ILUPreconditioner{a}: pc_out receives pc in the pattern of `a` (entries of `a`'s pattern, row-major ascending
// column; the constructor never creates or drops entries), nnzs_out its nnzs field.
int ref_ilu(int n, int nnz, const int* rp, const int* ci, const double* va, double* pc_out, int* nnzs_out) {
  try {
    sls::ILUPreconditioner ilu{make(n, n, nnz, rp, ci, va)};
    for (int i = 0; i < n; i++)
      for (int k = rp[i]; k < rp[i + 1]; k++) pc_out[k] = ilu.pc.dok.at(i).at(ci[k]);
    if (nnzs_out) *nnzs_out = ilu.pc.nnzs;
    return 0;
  } catch (std::exception& e) { g_serr = e.what(); return -1; }
}

// ILUPreconditioner{a}.apply(x)
int ref_ilu_apply(int n, int nnz, const int* rp, const int* ci, const double* va, const double* x, double* z) {
  try {
    sls::ILUPreconditioner ilu{make(n, n, nnz, rp, ci, va)};
    std::vector<double> r = ilu.apply(std::vector<double>(x, x + n));
    std::copy(r.begin(), r.end(), z);
    return 0;
  } catch (std::exception& e) { g_serr = e.what(); return -1; }
}

}  // extern "C"
