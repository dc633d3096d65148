/* CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the CASK hot path (caskorg/cask @ 9e561d7): the row-stripe /
 * column-block partitioner, the reference y = A*x, the `pcg` loop and Eigen's BiCGSTAB loop.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (libcask_b200.so) never links,
 * loads or calls anything in oracle/.
 *
 * Pinning: every function below is checked in tests/test_oracle_*.py against
 *   (1) the compiled reference itself (oracle/_ref/libcaskref.so, built in place from
 *       /root/reference/src/runtime/Spmv.cpp) where /root/reference exists, and
 *   (2) the committed golden fixtures in tests/golden/ that were generated from (1)
 *       by tests/golden/make_golden.py.
 * BiCGSTAB is the exception: its arithmetic lives in Eigen 3.3.1 (CMakeLists.txt:24), which is
 * neither in this image nor vendored in the reference, and no reference test pins it
 * ("parity unpinned" for oracle_bicgstab — cross-checked against scipy and a direct solve only).
 */
#ifndef CASK_ORACLE_H
#define CASK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#pragma pack(push, 1)
typedef struct { double value; int32_t indptr; } oracle_pair; /* Spmv.hpp:13-20, 12 bytes */
#pragma pack(pop)

typedef struct {
  /* Spmv.hpp:25-31 */
  int32_t nBlocks, n, paddingCycles, totalCycles, vector_load_cycles, outSize;
  int32_t reductionCycles, emptyCycles;
  int32_t m_colptr_unpaddedLength, m_indptr_values_unpaddedLength;
  int64_t len_colptr, len_pairs;
  int32_t* m_colptr;
  oracle_pair* m_indptr_values;
} oracle_partition;

enum { ORACLE_ARCH_SIMPLE = 0, ORACLE_ARCH_SKIPEMPTY = 1 };

/* Spmv::countComputeCycles, Spmv.cpp:25-40 */
int32_t oracle_count_compute_cycles(const int32_t* end_offsets, int32_t size, int32_t input_width);

/* SkipEmptyRowsSpmv::encodeEmptyRows, Spmv.hpp:213-238.  out must hold `size` entries; returns length. */
int32_t oracle_encode_empty_rows(const int32_t* end_offsets, int32_t size, int32_t* out);

/* CsrMatrix::sliceRows (SparseMatrix.hpp:426-443) + Spmv::do_blocking (Spmv.cpp:42-107,
 * via CsrMatrix::sliceColumns SparseMatrix.hpp:459-482) on rows [start, start+nrows). */
int oracle_do_blocking(int32_t m, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
                       int32_t start, int32_t nrows, int32_t block_size, int32_t input_width, int arch,
                       oracle_partition* out);

/* Spmv::preprocess, Spmv.cpp:329-365.  parts must hold num_pipes entries. */
int oracle_preprocess(int32_t n, int32_t m, const int32_t* row_ptr, const int32_t* col_ind,
                      const double* values, int arch, int32_t num_pipes, int32_t cache_size,
                      int32_t input_width, oracle_partition* parts);
void oracle_free_partition(oracle_partition* p);

/* y = A x as DokMatrix::dot computes it (SparseMatrix.hpp:255-264 through :376-384,422-424):
 * per row, ascending column, result[row] += x[col] * value, separate multiply and add. */
void oracle_csr_dot(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
                    const double* x, double* y);

/* What the device computes from the partition arrays: per stripe, blocks in
 * ascending column order, per row accumulate value * xcache[idx]; bit-31 entries skip empty rows. */
int oracle_partition_spmv_w(const oracle_partition* parts, int32_t nparts, int32_t cache_size,
                            int32_t input_width, const double* x, int32_t m, double* y,
                            int32_t n_total);

/* pcg<double, IdentityPreconditioner>, SparseLinearSolvers.hpp:162-239; `a` is the LOWER triangle
 * (0-based CSR) exactly as readSymMatrix hands it over; mkl_dcsrsymv('l') is restated as a
 * symmetric product from the stored triangle.  Returns 1 if converged.  *iterations keeps the
 * reference's quirk (assigned only at the end of a non-converged iteration). */
int oracle_pcg(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
               const double* rhs, double* x, int32_t* iterations, int32_t maxiters, double tol);

/* Same loop on a general (full) CSR matrix: what the GPU solver runs after symmetric expansion. */
int oracle_pcg_full(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
                    const double* rhs, double* x, int32_t* iterations, int32_t maxiters, double tol,
                    double* rs_final);

/* ---- preconditioners ----
 * ILUPreconditioner's constructor (SparseLinearSolvers.hpp:89-140): ILU(0), IKJ order, on the pattern of the
 * CSR it is given (rows in ascending column order); pc has rp[n] entries in that pattern. */
int oracle_ilu0(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* values, double* pc);
/* ILUPreconditioner::apply (:142-150): L y = x, U z = y, both NON-unit (MklLayer.hpp:66-84). Returns 1 if a
 * diagonal entry is missing or zero. */
int oracle_ilu_apply(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* pc,
                     const double* x, double* z);
/* unit_lower != 0: the textbook variant (unit diagonal in the lower solve) - not the reference's arithmetic */
int oracle_ilu_apply_mode(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* pc,
                          const double* x, double* z, int unit_lower);
/* pcg<double, Precon> (:162-239). precon: 0 identity, 1 ILU (built from the same arrays), 2 Jacobi (not in the
 * reference), 3 ILU(0) with a unit lower solve (not in the reference).  lower != 0: stored lower triangle + symmetric product (the reference's call); 0: full matrix.
 * Returns 1 converged, 0 not, -1 if the ILU solve met a zero pivot. */
int oracle_pcg_precond(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
                       const double* rhs, double* x, int32_t* iterations, int32_t maxiters, double tol,
                       int precon, int lower, double* rs_final);

/* Eigen 3.3.1 bicgstab() with DiagonalPreconditioner, as called at SparseLinearSolvers.cpp:18-26.
 * In: *iters = max iterations, *tol_error = tolerance.  Out: iterations done, relative residual. */
int oracle_bicgstab(int32_t n, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
                    const double* b, double* x, int32_t* iters, double* tol_error);

/* Synthetic matrices of BASELINE.json.  Call with NULL arrays to get nnz. */
int64_t oracle_gen_poisson2d(int32_t N, int32_t* row_ptr, int32_t* col_ind, double* values);
int64_t oracle_gen_poisson3d27(int32_t N, int32_t* row_ptr, int32_t* col_ind, double* values);
int64_t oracle_gen_convdiff3d7(int32_t N, int32_t* row_ptr, int32_t* col_ind, double* values);
/* R-MAT: counter-based (splitmix64) edges, (a,b,c,d)=(0.57,0.19,0.19,0.05); duplicates summed,
 * rows sorted by column.  Two-call protocol like the others (row_ptr NULL -> returns nnz). */
int64_t oracle_gen_rmat(int32_t scale, int32_t edge_factor, uint64_t seed, int32_t* row_ptr,
                        int32_t* col_ind, double* values);

/* CPU baseline port: OpenMP static row loop (what Eigen's row-major product / mkl_dcsrgemv do).
 * Returns the number of threads used. 64-bit row pointers so BASELINE sizes fit. */
int oracle_csr_spmv_omp(int64_t n, const int64_t* row_ptr, const int32_t* col_ind,
                        const double* values, const double* x, double* y);
int oracle_csr_spmv_omp32(int32_t n, const int32_t* row_ptr, const int32_t* col_ind,
                          const double* values, const double* x, double* y);
int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
