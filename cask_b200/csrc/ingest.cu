// Matrix Market ingest on the GPU: coordinate entries -> CSR with the reference's dictionary-of-keys semantics
// (io::readMatrix / readSymMatrix, src/runtime/IO.hpp:124-176; DokMatrix::explicitSymmetric and CsrMatrix(DokMatrix),
// src/runtime/SparseMatrix.hpp:156-189, 289-305).  The algorithm is in ingest_logic.inl; this file is its CUDA
// instantiation and the C ABI around it.
#include <cstring>

#include "ctx.cuh"
#include "devlogic.cuh"
#include "mmio.hpp"

#include "ingest_logic.inl"

struct cask_b200_csr {
  int device = 0;
  caskb200::ingest::CsrArrays a;
};

using namespace caskb200;

namespace {

int finish(cask_b200_ctx* ctx, const ingest::CsrArrays& arrays, int32_t err, int64_t first_bad, cask_b200_csr** out) {
  if (err & ingest::kErrBadIndex)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ingest: entry " + std::to_string(first_bad + 1) + " has an index outside the matrix");
  if (err & ingest::kErrNotSymmetric) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Matrix is not symmetric");  // SparseMatrix.hpp:173
  cask_b200_csr* h = new cask_b200_csr();
  h->device = ctx->device;
  h->a = arrays;
  *out = h;
  return CASK_B200_OK;
}

}  // namespace

extern "C" {

int cask_b200_ingest_coo_device(cask_b200_ctx* ctx, int64_t n, int64_t m, int64_t count, const int32_t* d_rows,
                                const int32_t* d_cols, const double* d_vals, int32_t flags, cask_b200_csr** out) {
  if (!ctx || !out) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ingest: null");
  *out = nullptr;
  if (count > 0 && (!d_rows || !d_cols || !d_vals)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ingest: null arrays");
  CB_TRY(ensure_device(ctx));
  dev::Exec ex = dev::exec_of(ctx);
  ingest::CsrArrays arrays;
  int32_t err = 0;
  int64_t first_bad = -1;
  CB_TRY(ingest::coo_to_csr(ex, n, m, count, d_rows, d_cols, d_vals, flags, &arrays, &err, &first_bad));
  return finish(ctx, arrays, err, first_bad, out);
}

int cask_b200_ingest_coo(cask_b200_ctx* ctx, int64_t n, int64_t m, int64_t count, const int32_t* rows, const int32_t* cols,
                         const double* vals, int32_t flags, cask_b200_csr** out) {
  if (!ctx || !out) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ingest: null");
  *out = nullptr;
  if (count < 0 || (count > 0 && (!rows || !cols || !vals))) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ingest: null arrays");
  CB_TRY(ensure_device(ctx));
  dev::Exec ex = dev::exec_of(ctx);
  ingest::Scratch in;
  CB_TRY(dev::alloc(&in.p[0], sizeof(int32_t) * (size_t)count));
  CB_TRY(dev::alloc(&in.p[1], sizeof(int32_t) * (size_t)count));
  CB_TRY(dev::alloc(&in.p[2], sizeof(double) * (size_t)count));
  CB_TRY(dev::upload(ex, in.p[0], rows, sizeof(int32_t) * (size_t)count));
  CB_TRY(dev::upload(ex, in.p[1], cols, sizeof(int32_t) * (size_t)count));
  CB_TRY(dev::upload(ex, in.p[2], vals, sizeof(double) * (size_t)count));
  return cask_b200_ingest_coo_device(ctx, n, m, count, (const int32_t*)in.p[0], (const int32_t*)in.p[1], (const double*)in.p[2],
                                     flags, out);
}

int cask_b200_read_matrix(cask_b200_ctx* ctx, const char* path, int32_t mode, cask_b200_csr** out) {
  if (!ctx || !path || !out) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "read_matrix: null");
  *out = nullptr;
  if (mode != 0 && mode != 1) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "read_matrix: mode must be 0 (readMatrix) or 1 (readSymMatrix)");
  try {  // no C++ exception crosses the C ABI (host allocations of the tokeniser, std::thread)
  mm::File f;
  CB_TRY(mm::read_header(path, &f));
  const std::string p(path);
  if (!f.matrix()) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Error! Expecting MatrixMarket matrix in " + p);  // IO.hpp:153-155
  if (mode == 1 && !f.symmetric())                                                                                // IO.hpp:171-174
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Error! Matrix found in " + p +
                                                    " is not symmetric. To read unsymmetric matrix use cask::io::readSymMatrix()");
  CB_TRY(mm::read_coo(path, &f));
  const int32_t flags = CASK_B200_INGEST_ONE_BASED | (mode == 0 && f.symmetric() ? CASK_B200_INGEST_SYMMETRIC : 0);
  return cask_b200_ingest_coo(ctx, f.n, f.m, f.l, f.rows.data(), f.cols.data(), f.vals.data(), flags, out);
  } catch (const std::exception& e) {
    return fail(CASK_B200_ERR_RUNTIME, std::string("read_matrix: ") + e.what());
  } catch (...) {
    return fail(CASK_B200_ERR_RUNTIME, "read_matrix: unknown exception");
  }
}

int cask_b200_csr_get_info(const cask_b200_csr* csr, int64_t* n, int64_t* m, int64_t* nnz, int64_t* nnzs_field) {
  if (!csr) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "csr_get_info: null");
  if (n) *n = csr->a.n;
  if (m) *m = csr->a.m;
  if (nnz) *nnz = csr->a.nnz;
  if (nnzs_field) *nnzs_field = csr->a.nnzs_field;
  return CASK_B200_OK;
}

int cask_b200_csr_export(cask_b200_ctx* ctx, const cask_b200_csr* csr, int32_t* row_ptr, int32_t* col_ind, double* values) {
  if (!ctx || !csr) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "csr_export: null");
  if (ctx->device != csr->device) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "csr_export: matrix lives on another device");
  CB_TRY(ensure_device(ctx));
  dev::Exec ex = dev::exec_of(ctx);
  if (row_ptr) CB_TRY(dev::download(ex, row_ptr, csr->a.row_ptr, sizeof(int32_t) * (size_t)(csr->a.n + 1)));
  if (col_ind) CB_TRY(dev::download(ex, col_ind, csr->a.col, sizeof(int32_t) * (size_t)csr->a.nnz));
  if (values) CB_TRY(dev::download(ex, values, csr->a.val, sizeof(double) * (size_t)csr->a.nnz));
  return CASK_B200_OK;
}

int cask_b200_csr_device_arrays(const cask_b200_csr* csr, const int32_t** d_row_ptr, const int32_t** d_col_ind,
                                const double** d_values) {
  if (!csr) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "csr_device_arrays: null");
  if (d_row_ptr) *d_row_ptr = csr->a.row_ptr;
  if (d_col_ind) *d_col_ind = csr->a.col;
  if (d_values) *d_values = csr->a.val;
  return CASK_B200_OK;
}

int cask_b200_csr_free(cask_b200_csr* csr) {
  if (!csr) return CASK_B200_OK;
  cudaSetDevice(csr->device);
  dev::release(csr->a.row_ptr);
  dev::release(csr->a.col);
  dev::release(csr->a.val);
  delete csr;
  return CASK_B200_OK;
}

int cask_b200_preprocess_csr(cask_b200_ctx* ctx, const cask_b200_design* design, const cask_b200_csr* csr) {
  if (!ctx || !csr) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess_csr: null");
  if (ctx->device != csr->device) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "preprocess_csr: matrix lives on another device");
  return cask_b200_preprocess_device(ctx, design, csr->a.n, csr->a.m, csr->a.nnz, csr->a.row_ptr, csr->a.col, csr->a.val);
}

}  // extern "C"
