// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// C-callable driver around the UNMODIFIED reference sources.  It is compiled together with
// /root/reference/src/runtime/Spmv.cpp *where that file lies* (see oracle/Makefile, target
// `ref`) into oracle/_ref/libcaskref.so.  No reference source is copied into this repository.
//
// What it exposes (all results come from the reference's own code):
//   * io::readMatrix / io::readSymMatrix / io::readVector        (src/runtime/IO.hpp:73-176)
//   * Spmv::preprocess / SkipEmptyRowsSpmv::preprocess            (src/runtime/Spmv.cpp:329-365)
//     and the resulting std::vector<Partition>                    (src/runtime/Spmv.hpp:25-44)
//   * CsrMatrix::dot                                              (src/runtime/SparseMatrix.hpp:422-424)
//   * CsrMatrix::sliceRows / sliceColumns                         (src/runtime/SparseMatrix.hpp:426-482)
//
// `partitions` is a private member of cask::spmv::Spmv (Spmv.hpp:51); this translation unit
// re-declares access for itself only so the real preprocess() can be executed and inspected.
#include <algorithm>
#include <cassert>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <map>
#include <regex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <Eigen/Sparse>
#include <boost/algorithm/string.hpp>
#define private public
#define protected public
#define class struct  // Spmv's first members are private by `class` default (Spmv.hpp:49-53)
#include "Spmv.hpp"
#undef class
#undef private
#undef protected
#include "IO.hpp"

#include <chrono>
#include <cstring>
#include <memory>
#include <string>

using cask::CsrMatrix;
using cask::spmv::Partition;

namespace {
struct RefHandle {
  CsrMatrix mat;
  std::unique_ptr<cask::spmv::Spmv> arch;
  std::string err;
};
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_sizeof_pair() { return (int)sizeof(cask::spmv::indptr_value); }

// ---- matrices -------------------------------------------------------------------------
// kind: 0 = io::readMatrix (symmetric files expanded), 1 = io::readSymMatrix (lower triangle kept)
void* ref_read_matrix(const char* path, int kind) {
  try {
    auto* h = new RefHandle();
    if (kind == 0) h->mat = cask::io::readMatrix(path);
    else h->mat = cask::io::readSymMatrix(path).matrix;
    return h;
  } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

void* ref_from_csr(int n, int m, int nnz, const int* row_ptr, const int* col_ind, const double* values) {
  auto* h = new RefHandle();
  h->mat = CsrMatrix(n, m, nnz,
                     std::vector<double>(values, values + nnz),
                     std::vector<int>(col_ind, col_ind + nnz),
                     std::vector<int>(row_ptr, row_ptr + n + 1));
  return h;
}

void ref_free(void* hv) { delete (RefHandle*)hv; }

void ref_dims(void* hv, int* n, int* m, int* nnz) {
  auto* h = (RefHandle*)hv;
  *n = h->mat.n; *m = h->mat.m; *nnz = (int)h->mat.values.size();
}

void ref_get_csr(void* hv, int* row_ptr, int* col_ind, double* values) {
  auto* h = (RefHandle*)hv;
  std::memcpy(row_ptr, h->mat.row_ptr.data(), sizeof(int) * h->mat.row_ptr.size());
  std::memcpy(col_ind, h->mat.col_ind.data(), sizeof(int) * h->mat.col_ind.size());
  std::memcpy(values, h->mat.values.data(), sizeof(double) * h->mat.values.size());
}

int ref_read_vector(const char* path, double* out, int cap) {
  try {
    cask::Vector v = cask::io::readVector(path);
    if (out) for (int i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    return v.size();
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// ---- y = A x via the reference's own CsrMatrix::dot ------------------------------------
// NB DokMatrix::dot sizes its result by x (SparseMatrix.hpp:256); callers pass len(x) >= n.
double ref_dot(void* hv, const double* x, int xlen, double* y, int ylen) {
  auto* h = (RefHandle*)hv;
  cask::Vector xv(std::vector<double>(x, x + xlen));
  auto t0 = std::chrono::high_resolution_clock::now();
  cask::Vector r = h->mat.dot(xv);
  double s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  for (int i = 0; i < ylen && i < r.size(); i++) y[i] = r[i];
  return s;
}

// ---- preprocess -------------------------------------------------------------------------
// arch: 0 = Spmv ("Simple"), 1 = SkipEmptyRowsSpmv ("SkipEmpty").  Returns seconds, <0 on error.
double ref_preprocess(void* hv, int arch, int num_pipes, int cache_size, int input_width,
                      int max_rows, int num_controllers) {
  auto* h = (RefHandle*)hv;
  try {
    if (arch == 0)
      h->arch.reset(new cask::spmv::Spmv(cache_size, input_width, num_pipes, max_rows, num_controllers));
    else
      h->arch.reset(new cask::spmv::SkipEmptyRowsSpmv(cache_size, input_width, num_pipes, max_rows,
                                                     num_controllers));
    auto t0 = std::chrono::high_resolution_clock::now();
    h->arch->preprocess(h->mat);
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  } catch (std::exception& e) { g_err = e.what(); return -1.0; }
}

int ref_num_partitions(void* hv) { return (int)((RefHandle*)hv)->arch->partitions.size(); }

// scalars[12]: nBlocks, n, paddingCycles, totalCycles, vector_load_cycles, outSize,
//              reductionCycles, emptyCycles, m_colptr_unpaddedLength,
//              m_indptr_values_unpaddedLength, len(m_colptr), len(m_indptr_values)
void ref_partition_scalars(void* hv, int p, int64_t* s) {
  const Partition& q = ((RefHandle*)hv)->arch->partitions[p];
  s[0] = q.nBlocks; s[1] = q.n; s[2] = q.paddingCycles; s[3] = q.totalCycles;
  s[4] = q.vector_load_cycles; s[5] = q.outSize; s[6] = q.reductionCycles; s[7] = q.emptyCycles;
  s[8] = q.m_colptr_unpaddedLength; s[9] = q.m_indptr_values_unpaddedLength;
  s[10] = (int64_t)q.m_colptr.size(); s[11] = (int64_t)q.m_indptr_values.size();
}

void ref_partition_arrays(void* hv, int p, int* colptr, void* pairs) {
  const Partition& q = ((RefHandle*)hv)->arch->partitions[p];
  if (colptr) std::memcpy(colptr, q.m_colptr.data(), sizeof(int) * q.m_colptr.size());
  if (pairs)
    std::memcpy(pairs, q.m_indptr_values.data(),
                sizeof(cask::spmv::indptr_value) * q.m_indptr_values.size());
}

double ref_estimated_clock_cycles(void* hv) { return ((RefHandle*)hv)->arch->getEstimatedClockCycles(); }

// Runs the reference's Spmv::spmv with its mock device callbacks (returns zeros; exercises the
// argument checks and exception messages of Spmv.cpp:189-232).  rc: 0 ok, 1 invalid_argument,
// 2 runtime_error; message in ref_last_error().
int ref_spmv_mock(void* hv, const double* x, int xlen, double* y, int ylen) {
  auto* h = (RefHandle*)hv;
  try {
    cask::Vector r = h->arch->spmv(cask::Vector(std::vector<double>(x, x + xlen)));
    for (int i = 0; i < ylen && i < r.size(); i++) y[i] = r[i];
    return 0;
  } catch (std::invalid_argument& e) { g_err = e.what(); return 1; }
  catch (std::runtime_error& e) { g_err = e.what(); return 2; }
}

// ---- slices (unit-level pins for the restatement) ---------------------------------------
// sliceRows: returns nnz of slice; arrays sized by caller (row_ptr: nRows+1).
int ref_slice_rows(void* hv, int start, int nrows, int* row_ptr, int* col_ind, double* values) {
  auto s = ((RefHandle*)hv)->mat.sliceRows(start, nrows);
  if (row_ptr) std::memcpy(row_ptr, s.row_ptr.data(), sizeof(int) * s.row_ptr.size());
  if (col_ind) std::memcpy(col_ind, s.col_ind.data(), sizeof(int) * s.col_ind.size());
  if (values) std::memcpy(values, s.values.data(), sizeof(double) * s.values.size());
  return (int)s.values.size();
}

// sliceColumns: block b of sliceColumns(blockSize); returns nnz in that block, -1 if b out of range.
int ref_slice_columns(void* hv, int blockSize, int b, int* row_ptr, int* col_ind, double* values) {
  auto v = ((RefHandle*)hv)->mat.sliceColumns(blockSize);
  if (b < 0 || b >= (int)v.size()) return -1;
  auto& s = v[b];
  if (row_ptr) std::memcpy(row_ptr, s.row_ptr.data(), sizeof(int) * s.row_ptr.size());
  if (col_ind) std::memcpy(col_ind, s.col_ind.data(), sizeof(int) * s.col_ind.size());
  if (values) std::memcpy(values, s.values.data(), sizeof(double) * s.values.size());
  return (int)s.values.size();
}

}  // extern "C"
