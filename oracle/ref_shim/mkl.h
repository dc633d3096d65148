// Stand-in for Intel MKL's <mkl.h> (absent from this image; the reference finds it through $MKLROOT,
// src/cmake/FindMKL.cmake:1-35).  TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the reference's
// OWN pcg<> loop and ILUPreconditioner (src/runtime/SparseLinearSolvers.hpp:75-239, MklLayer.hpp) where
// they lie, with -DUSEMKL, so that the reference code - not a restatement of it - produces the CG / ILU
// fixtures.  Only the six routines those files call exist.  Each follows MKL's documented contract
// (1-based CSR, `uplo` selects the triangle that is read, diag = 'N' takes the diagonal from the matrix);
// MKL's internal summation order is not documented and cannot be reproduced: sums run in storage order.
#pragma once
#include <cstddef>
typedef int MKL_INT;

// y = A x, A symmetric, only the `uplo` triangle of the 1-based CSR is referenced
inline void mkl_dcsrsymv(const char* uplo, const MKL_INT* m, const double* a, const MKL_INT* ia, const MKL_INT* ja,
                         const double* x, double* y) {
  const bool lower = *uplo == 'l' || *uplo == 'L';
  for (MKL_INT i = 0; i < *m; i++) y[i] = 0.0;
  for (MKL_INT i = 0; i < *m; i++)
    for (MKL_INT k = ia[i] - 1; k < ia[i + 1] - 1; k++) {
      const MKL_INT j = ja[k] - 1;
      if (lower ? j > i : j < i) continue;
      y[i] += a[k] * x[j];
      if (j != i) y[j] += a[k] * x[i];
    }
}

// solve T y = x with T the `uplo` triangle of the 1-based CSR (transa = 'N' only; diag 'N': non-unit, 'U': unit)
inline void mkl_dcsrtrsv(const char* uplo, const char* transa, const char* diag, const MKL_INT* m, const double* a,
                         const MKL_INT* ia, const MKL_INT* ja, const double* x, double* y) {
  (void)transa;
  const bool lower = *uplo == 'l' || *uplo == 'L';
  const bool unit = *diag == 'u' || *diag == 'U';
  const MKL_INT n = *m;
  for (MKL_INT s = 0; s < n; s++) {
    const MKL_INT i = lower ? s : n - 1 - s;
    double acc = x[i], d = 1.0;
    for (MKL_INT k = ia[i] - 1; k < ia[i + 1] - 1; k++) {
      const MKL_INT j = ja[k] - 1;
      if (j == i) { if (!unit) d = a[k]; }
      else if (lower ? j < i : j > i) acc -= a[k] * y[j];
    }
    y[i] = acc / d;
  }
}

inline double cblas_ddot(MKL_INT n, const double* x, MKL_INT incx, const double* y, MKL_INT incy) {
  double s = 0.0;
  for (MKL_INT i = 0; i < n; i++) s += x[(size_t)i * incx] * y[(size_t)i * incy];
  return s;
}
inline void cblas_daxpy(MKL_INT n, double alpha, const double* x, MKL_INT incx, double* y, MKL_INT incy) {
  for (MKL_INT i = 0; i < n; i++) y[(size_t)i * incy] += alpha * x[(size_t)i * incx];
}
inline void cblas_daxpby(MKL_INT n, double alpha, const double* x, MKL_INT incx, double beta, double* y, MKL_INT incy) {
  for (MKL_INT i = 0; i < n; i++) y[(size_t)i * incy] = alpha * x[(size_t)i * incx] + beta * y[(size_t)i * incy];
}
inline void mkl_free_buffers() {}
