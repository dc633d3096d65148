// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// C-callable driver around the reference's OWN solver code, compiled where it lies: pcg<T, Precon>,
// IdentityPreconditioner and ILUPreconditioner (src/runtime/SparseLinearSolvers.hpp:62-239) with -DUSEMKL and
// the six MKL routines they call supplied by ref_shim/mkl.h (sequential stand-in; part of oracle/_ref/libcaskref.so,
// oracle/Makefile target `ref`) or by ref_shim_mkl/mkl.h (adapters over the real oneMKL that libtorch_cpu.so exports;
// oracle/_ref/libcaskref_mkl.so, target `ref_mkl`).  No reference source is copied into the repository.
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
// Returns 1 converged / 0 not / -1 exception.  *solve_seconds (optional) receives the reference's own "cg:solve"
// timer (cask::utils::Timer, Utils.hpp:15-50): the loop without the set-up copies, as CgTest / sparse-bench report it.
int ref_pcg_timed(int n, int nnz, const int* rp, const int* ci, const double* va, const double* rhs, double* x,
                  int* iterations, int precon, double* solve_seconds) {
  try {
    CsrMatrix a = make(n, n, nnz, rp, ci, va);
    std::vector<double> b(rhs, rhs + n);
    cask::utils::Timer t;
    const bool ok = precon == 1 ? sls::pcg<double, sls::ILUPreconditioner>(a, b.data(), x, *iterations, false, &t)
                                : sls::pcg<double, sls::IdentityPreconditioner>(a, b.data(), x, *iterations, false, &t);
    if (solve_seconds) *solve_seconds = t.get("cg:solve").count();
    return ok ? 1 : 0;
  } catch (std::exception& e) { g_serr = e.what(); return -1; }
}

int ref_pcg(int n, int nnz, const int* rp, const int* ci, const double* va, const double* rhs, double* x,
            int* iterations, int precon) {
  return ref_pcg_timed(n, nnz, rp, ci, va, rhs, x, iterations, precon, nullptr);
}

// y = A x exactly as the reference's pcg<> computes it (SparseLinearSolvers.hpp:175-190, 206): one-based copies of the
// CSR handed over, then mkl_dcsrsymv('l', ...) - A symmetric, only its stored lower triangle read.  `reps` timed calls
// after one warm-up; *seconds receives the total of the timed calls.
int ref_symv_timed(int n, int nnz, const int* rp, const int* ci, const double* va, const double* x, double* y, int reps,
                   double* seconds) {
  try {
    CsrMatrix a = make(n, n, nnz, rp, ci, va);
    char tr = 'l';
    auto values = a.values;
    auto row_ptr = a.getRowPtrWithOneBasedIndex();
    auto col_ind = a.getColIndWithOneBasedIndex();
    mkl_dcsrsymv(&tr, &n, values.data(), row_ptr.data(), col_ind.data(), x, y);
    cask::utils::Timer t;
    t.tic("symv");
    for (int r = 0; r < reps; r++) mkl_dcsrsymv(&tr, &n, values.data(), row_ptr.data(), col_ind.data(), x, y);
    if (seconds) *seconds = t.toc("symv").count();
    return 0;
  } catch (std::exception& e) { g_serr = e.what(); return -1; }
}

// 1 when this build runs on Intel MKL's arithmetic (ref_shim_mkl/mkl.h), 0 with the sequential stand-in (ref_shim/mkl.h)
int ref_solvers_real_mkl() {
#ifdef CASK_REF_REAL_MKL
  return 1;
#else
  return 0;
#endif
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
