// Host containers with the reference's public surface (src/runtime/SparseMatrix.hpp): cask::Vector,
// DokMatrix, CsrMatrix, SymCsrMatrix, sparse::SparkCooMatrix.  These are plain host data holders that
// feed the GPU path; the arithmetic members kept here (dot) exist for API compatibility and tests only.
#ifndef CASK_B200_HOST_SPARSEMATRIX_HPP
#define CASK_B200_HOST_SPARSEMATRIX_HPP
#include <algorithm>
#include <cassert>
#include <cmath>
#include <fstream>
#include <initializer_list>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace cask {
namespace sparse {
template <typename value_type>
class SparkCooMatrix {  // SparseMatrix.hpp:23-33
 public:
  using CoordType = std::tuple<int, int, value_type>;
  int n, m;
  std::vector<CoordType> data;
  SparkCooMatrix(int rows, int cols) : n(rows), m(cols) {}
};
}  // namespace sparse

class Vector {  // SparseMatrix.hpp:37-99
 public:
  std::vector<double> data;
  Vector(int n) : data(n, 0.0) {}
  Vector(std::initializer_list<double> l) : data(l) {}
  Vector(const std::vector<double>& v) : data(v) {}

  int size() const { return (int)data.size(); }
  double& operator[](int i) { return data[i]; }
  const double& operator[](int i) const { return data[i]; }
  bool operator==(const Vector& o) const { return data == o.data; }
  Vector operator-(const Vector& o) const {
    if (o.size() != size())
      throw std::invalid_argument("Attempt to subtract vectors of different lengths: " + std::to_string(o.size()) +
                                  " != " + std::to_string(size()));
    Vector r(size());
    for (int i = 0; i < size(); i++) r[i] = data[i] - o[i];
    return r;
  }
  // NB the reference reads an uninitialised accumulator here (SparseMatrix.hpp:82); this one starts at 0.
  double norm() const {
    double s = 0.0;
    for (double d : data) s += d * d;
    return std::sqrt(s);
  }
  double distance(const Vector& o) const { return (*this - o).norm(); }
  void print(const std::string& label = "") const {
    std::cout << label;
    for (double d : data) std::cout << d << " ";
    std::cout << std::endl;
  }
  void writeToFile(const std::string& path) const {
    std::ofstream f(path);
    if (!f) throw std::invalid_argument("Could not open file for writing");
    for (double d : data) f << d << std::endl;
  }
};

class DokMatrix {  // SparseMatrix.hpp:118-266
 public:
  int n, m, nnzs;
  std::unordered_map<int, std::map<int, double>> dok;  // std::map keeps each row sorted by column

  DokMatrix() : n(0), m(0), nnzs(0) {}
  DokMatrix(int rows, int cols) : n(rows), m(cols), nnzs(0) {}
  DokMatrix(int rows, int cols, int nz) : n(rows), m(cols), nnzs(nz) {}
  DokMatrix(const std::initializer_list<double>& dense)
      : DokMatrix((int)std::floor(std::sqrt((double)dense.size())), dense) {}
  DokMatrix(int rows, const std::initializer_list<double>& dense) : n(rows), nnzs(0) {
    m = (int)dense.size() / rows;
    auto it = dense.begin();
    for (int i = 0; i < n; i++)
      for (int j = 0; j < m; j++, ++it)
        if (*it != 0) { dok[i][j] = *it; nnzs++; }
  }

  double at(int i, int j) const {
    auto r = dok.find(i);
    if (r == dok.end()) return 0;
    auto c = r->second.find(j);
    return c == r->second.end() ? 0 : c->second;
  }
  void set(int i, int j, double v) { dok[i][j] = v; nnzs++; }
  bool isNnz(int i, int j) const { return at(i, j) != 0; }
  bool operator==(const DokMatrix& o) const { return n == o.n && m == o.m && nnzs == o.nnzs && dok == o.dok; }

  // mirror every off-diagonal entry; throws if a different transpose value is already stored
  DokMatrix explicitSymmetric() const {  // SparseMatrix.hpp:156-189
    DokMatrix r(n, m);
    for (const auto& row : dok)
      for (const auto& e : row.second) {
        const int i = row.first, j = e.first;
        r.dok[i][j] = e.second;
        r.nnzs++;
        if (i == j) continue;
        auto t = dok.find(j);
        if (t != dok.end()) {
          auto tv = t->second.find(i);
          if (tv != t->second.end()) {
            if (tv->second != e.second) throw std::invalid_argument("Matrix is not symmetric");
            std::cout << "Warning! Matrix already contains transpose entry for " << i << " " << j << std::endl;
          }
        }
        r.dok[j][i] = e.second;
        r.nnzs++;
      }
    return r;
  }
  DokMatrix getLowerTriangular() const { return triangle(true); }
  DokMatrix getUpperTriangular() const { return triangle(false); }

  Vector dot(const Vector& b) const {  // SparseMatrix.hpp:255-264
    Vector r(b.size());
    for (const auto& row : dok)
      for (const auto& e : row.second) r[row.first] += b[e.first] * e.second;
    return r;
  }
  void pretty_print() const {
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < m; j++) std::cout << at(i, j) << " ";
      std::cout << "\n";
    }
  }

 private:
  DokMatrix triangle(bool lower) const {
    DokMatrix r(n, m);
    for (const auto& row : dok)
      for (const auto& e : row.second)
        if (lower ? e.first <= row.first : row.first <= e.first) r.set(row.first, e.first, e.second);
    return r;
  }
};

class CsrMatrix {  // SparseMatrix.hpp:272-484; 0-based, row_ptr has n+1 entries
 public:
  int n, m, nnzs;
  std::vector<double> values;
  std::vector<int> col_ind;
  std::vector<int> row_ptr;

  CsrMatrix() : n(0), m(0), nnzs(0) {}
  CsrMatrix(std::initializer_list<double> dense) : CsrMatrix(DokMatrix(dense)) {}
  CsrMatrix(int rows, std::initializer_list<double> dense) : CsrMatrix(DokMatrix(rows, dense)) {}
  CsrMatrix(const DokMatrix& d) : n(d.n), m(d.m), nnzs(d.nnzs) {
    for (int i = 0; i < n; i++) {
      row_ptr.push_back((int)values.size());
      auto r = d.dok.find(i);
      if (r == d.dok.end()) continue;
      for (const auto& e : r->second) { col_ind.push_back(e.first); values.push_back(e.second); }
    }
    row_ptr.push_back(nnzs);
  }
  CsrMatrix(int rows, int cols, int nz, double* v, int* c, int* r)
      : n(rows), m(cols), nnzs(nz), values(v, v + nz), col_ind(c, c + nz), row_ptr(r, r + rows + 1) {}
  CsrMatrix(int rows, int cols, int nz, const std::vector<double>& v, const std::vector<int>& c,
            const std::vector<int>& r)
      : n(rows), m(cols), nnzs(nz), values(v), col_ind(c), row_ptr(r) {}

  bool operator==(const CsrMatrix& o) const {
    return n == o.n && m == o.m && nnzs == o.nnzs && values == o.values && row_ptr == o.row_ptr && col_ind == o.col_ind;
  }
  double& get(int i, int j) {
    for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++)
      if (col_ind[k] == j) return values[k];
    throw std::invalid_argument("No nonzero at row col:" + std::to_string(i) + " " + std::to_string(j));
  }
  bool isNnz(int i, int j) {
    for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++)
      if (col_ind[k] == j) return true;
    return false;
  }
  bool isSymmetric() const { return true; }
  DokMatrix toDok() const {
    DokMatrix d(n, m, nnzs);
    for (int i = 0; i < n; i++)
      for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++) d.dok[i][col_ind[k]] = values[k];
    return d;
  }
  std::vector<int> getRowPtrWithOneBasedIndex() const { return plus_one(row_ptr); }
  std::vector<int> getColIndWithOneBasedIndex() const { return plus_one(col_ind); }
  CsrMatrix getLowerTriangular() const { return CsrMatrix(toDok().getLowerTriangular()); }
  CsrMatrix getUpperTriangular() const { return CsrMatrix(toDok().getUpperTriangular()); }
  Vector dot(const Vector& b) const { return toDok().dot(b); }  // SparseMatrix.hpp:422-424

  // contiguous stripe of rows: row_ptr rebased to 0, columns untouched (SparseMatrix.hpp:426-443)
  CsrMatrix sliceRows(int startRow, int nRows) const {
    const int b = row_ptr[startRow], e = row_ptr[startRow + nRows];
    std::vector<int> rp(nRows + 1);
    for (int i = 0; i <= nRows; i++) rp[i] = row_ptr[startRow + i] - b;
    return CsrMatrix(nRows, m, e - b, std::vector<double>(values.begin() + b, values.begin() + e),
                     std::vector<int>(col_ind.begin() + b, col_ind.begin() + e), rp);
  }
  // column blocks of width blockSize: per block n cumulative END offsets (no leading 0), block-local
  // column indices, original entry order (SparseMatrix.hpp:459-482)
  std::vector<CsrMatrix> sliceColumns(int blockSize) const {
    const int nBlocks = m / blockSize + (m % blockSize ? 1 : 0);
    std::vector<CsrMatrix> blocks(nBlocks);
    for (auto& b : blocks) b.row_ptr.assign(n, 0);
    for (int i = 0; i < n; i++)
      for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
        CsrMatrix& b = blocks[col_ind[k] / blockSize];
        b.col_ind.push_back(col_ind[k] % blockSize);
        b.values.push_back(values[k]);
        b.row_ptr[i]++;
      }
    for (auto& b : blocks)
      for (int i = 1; i < n; i++) b.row_ptr[i] += b.row_ptr[i - 1];
    return blocks;
  }
  void pretty_print() const { toDok().pretty_print(); }

 private:
  static std::vector<int> plus_one(const std::vector<int>& v) {
    std::vector<int> r(v);
    for (int& x : r) x++;
    return r;
  }
};

// symmetric matrix, lower triangle stored (SparseMatrix.hpp:487-517)
class SymCsrMatrix {
 public:
  int n, m, nnzs;
  CsrMatrix matrix;
  explicit SymCsrMatrix(const DokMatrix& lower) : n(lower.n), m(lower.m), matrix(lower) {
    int diag = 0;
    for (int i = 0; i < lower.n; i++)
      if (lower.at(i, i) != 0) diag++;
    nnzs = 2 * (lower.nnzs - diag) + diag;
  }
  Vector dot(const Vector& b) const { return matrix.toDok().explicitSymmetric().dot(b); }
  void pretty_print() {
    std::cout << "Stored matrix: " << std::endl;
    matrix.pretty_print();
    std::cout << "Implicit values: " << std::endl;
    matrix.toDok().explicitSymmetric().pretty_print();
  }
};

}  // namespace cask
#endif
