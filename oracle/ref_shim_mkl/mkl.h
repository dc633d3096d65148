// <mkl.h> for the reference's solver code, backed by the REAL Intel MKL found in this image.
// TEST INFRASTRUCTURE ONLY (oracle/Makefile, target `ref_mkl` -> oracle/_ref/libcaskref_mkl.so).
//
// The image has no MKL development package, but PyTorch's libtorch_cpu.so links oneMKL 2024.2 statically and
// EXPORTS part of it: the inspector-executor sparse BLAS (mkl_sparse_d_create_csr / _mv / _trsv / _destroy),
// cblas_daxpy and the Fortran-interface ddot_.  The reference (src/runtime/SparseLinearSolvers.hpp:162-239,
// MklLayer.hpp:61-84) calls the older NIST-style entry points mkl_dcsrsymv / mkl_dcsrtrsv, which that export list
// does not contain.  The adapters below give those two names the same documented contract (1-based CSR, `uplo`
// selects the triangle that is read, diag = 'N' takes the diagonal from the matrix) on top of the routines that ARE
// exported, so the reference's pcg<> / ILUPreconditioner run with MKL's arithmetic - its summation order, its
// threading - instead of the sequential stand-in of ref_shim/mkl.h:
//     mkl_dcsrsymv(uplo, ...)            -> mkl_sparse_d_mv  (SPARSE_MATRIX_TYPE_SYMMETRIC,  fill = uplo, non-unit)
//     mkl_dcsrtrsv(uplo, 'N', diag, ...) -> mkl_sparse_d_trsv(SPARSE_MATRIX_TYPE_TRIANGULAR, fill = uplo, diag)
//     cblas_ddot                         -> ddot_   (MKL)
//     cblas_daxpy                        -> cblas_daxpy (MKL)
//     cblas_daxpby                       -> not exported: y = alpha x + beta y as a plain loop (-ffp-contract=off)
// The MKL symbols stay undefined in libcaskref_mkl.so; oracle/mklbind.py loads libtorch_cpu.so with RTLD_GLOBAL
// first.  Enum values are those of mkl_spblas.h (oneMKL 2024); tests/test_oracle_mkl.py checks them by behaviour.
#pragma once
#include <cstddef>
#include <stdexcept>
typedef int MKL_INT;

extern "C" {
struct cask_mkl_matrix_descr { int type, mode, diag; };  // struct matrix_descr, three enums, passed by value
int mkl_sparse_d_create_csr(void** A, int indexing, MKL_INT rows, MKL_INT cols, MKL_INT* rows_start, MKL_INT* rows_end,
                            MKL_INT* col_indx, double* values);
int mkl_sparse_d_mv(int operation, double alpha, void* A, cask_mkl_matrix_descr descr, const double* x, double beta,
                    double* y);
int mkl_sparse_d_trsv(int operation, double alpha, void* A, cask_mkl_matrix_descr descr, const double* x, double* y);
int mkl_sparse_destroy(void* A);
void cblas_daxpy(MKL_INT n, double alpha, const double* x, MKL_INT incx, double* y, MKL_INT incy);
double ddot_(const MKL_INT* n, const double* x, const MKL_INT* incx, const double* y, const MKL_INT* incy);
}

namespace cask_mkl_adapter {
enum { INDEX_ONE = 1, OP_N = 10, TYPE_SYMMETRIC = 21, TYPE_TRIANGULAR = 23, FILL_LOWER = 40, FILL_UPPER = 41,
       DIAG_NON_UNIT = 50, DIAG_UNIT = 51 };
struct Handle {  // a view of the caller's arrays: create_csr does not copy them
  void* h = nullptr;
  Handle(const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja) {
    const int rc = mkl_sparse_d_create_csr(&h, INDEX_ONE, *m, *m, const_cast<MKL_INT*>(ia), const_cast<MKL_INT*>(ia) + 1,
                                           const_cast<MKL_INT*>(ja), const_cast<double*>(a));
    if (rc != 0) throw std::runtime_error("mkl_sparse_d_create_csr failed");
  }
  ~Handle() { if (h) mkl_sparse_destroy(h); }
};
inline bool is(const char* c, char lo) { return *c == lo || *c == lo - 32; }
}  // namespace cask_mkl_adapter

inline void mkl_dcsrsymv(const char* uplo, const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja,
                         const double* x, double* y) {
  using namespace cask_mkl_adapter;
  Handle A(m, a, ia, ja);
  const cask_mkl_matrix_descr d = {TYPE_SYMMETRIC, is(uplo, 'l') ? FILL_LOWER : FILL_UPPER, DIAG_NON_UNIT};
  if (mkl_sparse_d_mv(OP_N, 1.0, A.h, d, x, 0.0, y) != 0) throw std::runtime_error("mkl_sparse_d_mv failed");
}

inline void mkl_dcsrtrsv(const char* uplo, const char* transa, const char* diag, const MKL_INT* m, const double* a,
                         const MKL_INT* ia, const MKL_INT* ja, const double* x, double* y) {
  using namespace cask_mkl_adapter;
  if (!is(transa, 'n')) throw std::runtime_error("mkl_dcsrtrsv adapter: transa must be 'N'");
  Handle A(m, a, ia, ja);
  const cask_mkl_matrix_descr d = {TYPE_TRIANGULAR, is(uplo, 'l') ? FILL_LOWER : FILL_UPPER,
                                   is(diag, 'u') ? DIAG_UNIT : DIAG_NON_UNIT};
  if (mkl_sparse_d_trsv(OP_N, 1.0, A.h, d, x, y) != 0) throw std::runtime_error("mkl_sparse_d_trsv failed");
}

inline double cblas_ddot(MKL_INT n, const double* x, MKL_INT incx, const double* y, MKL_INT incy) {
  return ddot_(&n, x, &incx, y, &incy);
}
inline void cblas_daxpby(MKL_INT n, double alpha, const double* x, MKL_INT incx, double beta, double* y, MKL_INT incy) {
  for (MKL_INT i = 0; i < n; i++) y[(size_t)i * incy] = alpha * x[(size_t)i * incx] + beta * y[(size_t)i * incy];
}
inline void mkl_free_buffers() {}
