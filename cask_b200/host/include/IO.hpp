// Matrix Market input with the reference's entry points (src/runtime/IO.hpp): readHeader, readVector,
// readDokMatrix, readMatrix, readSymMatrix, MmReader<T> on the host, and io::gpu::readMatrix / readSymMatrix through the
// GPU ingest path (cask_b200_read_matrix).
#ifndef CASK_B200_HOST_IO_HPP
#define CASK_B200_HOST_IO_HPP
#include <fstream>
#include <sstream>
#include <string>

#include "../../../include/cask_b200.h"
#include "SparseMatrix.hpp"

namespace cask {
namespace io {

struct MmInfo {  // IO.hpp:39-58
  std::string type, format, dataType, symmetry;
  bool isMatrix() const { return type == "matrix"; }
  bool isSymmetric() const { return symmetry == "symmetric"; }
  bool isCoordinate() const { return format == "coordinate"; }
};

// "%%MatrixMarket (matrix|array) (coordinate|array) (real|integer) (symmetric|general)" — IO.hpp:60-71
inline MmInfo readHeader(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::invalid_argument("File not found " + path);
  std::string line, banner;
  std::getline(f, line);
  std::istringstream ss(line);
  MmInfo i;
  ss >> banner >> i.type >> i.format >> i.dataType >> i.symmetry;
  const bool ok = banner == "%%MatrixMarket" && (i.type == "matrix" || i.type == "array") &&
                  (i.format == "coordinate" || i.format == "array") && (i.dataType == "real" || i.dataType == "integer") &&
                  (i.symmetry == "symmetric" || i.symmetry == "general");
  if (!ok) throw std::invalid_argument("Not a valid MatrixMarket file in " + path);
  return i;
}

namespace detail {
inline std::string first_data_line(std::ifstream& f) {
  std::string line;
  while (std::getline(f, line))
    if (line.empty() || line[0] != '%') break;
  return line;
}
}  // namespace detail

inline cask::Vector readVector(const std::string& path) {  // IO.hpp:73-115
  const MmInfo info = readHeader(path);
  std::ifstream f(path);
  std::istringstream dims(detail::first_data_line(f));
  int n = 0, m = 0;
  dims >> n >> m;
  Vector v(n);
  if (info.format == "coordinate") {
    int l = 0;
    dims >> l;
    for (int k = 0; k < l; k++) {
      int a, b;
      double val;
      f >> a >> b >> val;
      v[a] = val;  // sic: the reference does not rebase coordinate vectors
    }
    return v;
  }
  for (int i = 0; i < n; i++) f >> v[i];
  return v;
}

// symmetric files keep only the stored triangle; callers expand (IO.hpp:117-148)
inline DokMatrix readDokMatrix(const std::string& path, const MmInfo& info) {
  if (!info.isCoordinate()) throw std::invalid_argument("Expecting a coordinate MatrixMarket file in " + path);
  std::ifstream f(path);
  std::istringstream dims(detail::first_data_line(f));
  int n = 0, m = 0, l = 0;
  dims >> n >> m >> l;
  DokMatrix mat(n, m);
  for (int k = 0; k < l; k++) {
    int i, j;
    double val;
    f >> i >> j >> val;
    mat.set(i - 1, j - 1, val);
  }
  return mat;
}

inline cask::CsrMatrix readMatrix(const std::string& path) {  // IO.hpp:151-163
  const MmInfo info = readHeader(path);
  if (!info.isMatrix()) throw std::invalid_argument("Error! Expecting MatrixMarket matrix in " + path);
  if (info.isSymmetric()) return CsrMatrix(readDokMatrix(path, info).explicitSymmetric());
  return CsrMatrix(readDokMatrix(path, info));
}

inline cask::SymCsrMatrix readSymMatrix(const std::string& path) {  // IO.hpp:165-176
  const MmInfo info = readHeader(path);
  if (!info.isMatrix()) throw std::invalid_argument("Error! Expecting MatrixMarket matrix in " + path);
  if (!info.isSymmetric())
    throw std::invalid_argument("Error! Matrix found in " + path + " is not symmetric. To read unsymmetric matrix use cask::io::readSymMatrix()");
  return SymCsrMatrix(readDokMatrix(path, info));
}

// The same two readers with the tokenising on all host cores and the DokMatrix build on the GPU
// (cask_b200_read_matrix: radix sort + data-parallel passes instead of one hash-map insertion per entry; identical
// results, including "last value wins" for repeated keys and the "Matrix is not symmetric" check).  Opt-in for now:
// io::readMatrix / io::readSymMatrix above stay the host path until this one has its GPU parity run on record.
namespace gpu {
namespace detail {
inline CsrMatrix fetch(const std::string& path, int mode) {
  cask_b200_ctx* raw = nullptr;
  auto check = [](int rc) {
    if (rc == CASK_B200_OK) return;
    const std::string msg = cask_b200_last_error();
    if (rc == CASK_B200_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  };
  check(cask_b200_create(&raw, 0));
  struct Guard {
    cask_b200_ctx* c; cask_b200_csr* m;
    ~Guard() { cask_b200_csr_free(m); cask_b200_destroy(c); }
  } g{raw, nullptr};
  check(cask_b200_read_matrix(g.c, path.c_str(), mode, &g.m));
  int64_t n = 0, m = 0, nnz = 0, field = 0;
  check(cask_b200_csr_get_info(g.m, &n, &m, &nnz, &field));
  std::vector<int> rp((size_t)n + 1), ci((size_t)nnz);
  std::vector<double> va((size_t)nnz);
  check(cask_b200_csr_export(g.c, g.m, rp.data(), ci.data(), va.data()));
  rp[(size_t)n] = (int)field;  // CsrMatrix(const DokMatrix&) ends row_ptr with the nnzs FIELD (SparseMatrix.hpp:304)
  return CsrMatrix((int)n, (int)m, (int)field, va, ci, rp);
}
}  // namespace detail
inline cask::CsrMatrix readMatrix(const std::string& path) { return detail::fetch(path, 0); }  // IO.hpp:151-163
inline cask::SymCsrMatrix readSymMatrix(const std::string& path) {                             // IO.hpp:165-176
  return SymCsrMatrix(detail::fetch(path, 1).toDok());
}
}  // namespace gpu

// COO reader: 0-based, symmetric entries mirrored, sorted by (row, column) — IO.hpp:178-330
template <typename value_type>
class MmReader {
  std::string path_;
  std::ifstream f_;
  bool sparse_ = true, symmetric_ = false;
  int nrows_ = 0, ncols_ = 0, nnzs_ = -1;

  void parseHeader() {
    std::string line;
    if (!std::getline(f_, line)) throw std::invalid_argument("File " + path_ + " is empty");
    if (line.find("coordinate") != std::string::npos) sparse_ = true;
    else if (line.find("array") != std::string::npos) sparse_ = false;
    else throw std::invalid_argument("Cannot parse header, requires either 'coordinate' or 'matrix' type");
    if (line.compare(0, 2, "%%") == 0) {
      if (line.find("matrix") == std::string::npos) throw std::invalid_argument("Unsupported file type: " + line);
      symmetric_ = line.find("symmetric") != std::string::npos;
    }
    while (std::getline(f_, line) && !line.empty() && line[0] == '%') continue;
    std::istringstream dims(line);
    dims >> nrows_ >> ncols_;
    if (sparse_) dims >> nnzs_;
  }

 public:
  explicit MmReader(const std::string& path) : path_(path), f_(path) {
    if (!f_.is_open()) throw std::invalid_argument("Could not open file path " + path);
  }
  virtual ~MmReader() {}

  std::vector<double> readVector() {
    parseHeader();
    if (ncols_ > 1) throw std::invalid_argument("Object has > 1 columns ==> Use readMatrix");
    if (sparse_) throw std::invalid_argument("Sparse vectors not supported");
    std::vector<double> r;
    double v;
    while (f_ >> v) r.push_back(v);
    return r;
  }

  cask::sparse::SparkCooMatrix<value_type> mmreadMatrix(const std::string&) {
    parseHeader();
    if (ncols_ <= 1) throw std::invalid_argument("Matrix has only one column ==> Use readVector");
    cask::sparse::SparkCooMatrix<value_type> coo(nrows_, ncols_);
    std::string line;
    for (int k = 0; k < nnzs_; k++) {
      if (!std::getline(f_, line)) throw std::invalid_argument("File has less than given nonzeros!");
      std::istringstream ss(line);
      int i, j;
      value_type v;
      ss >> i >> j >> v;
      coo.data.emplace_back(i - 1, j - 1, v);
      if (symmetric_ && i != j) coo.data.emplace_back(j - 1, i - 1, v);
    }
    std::sort(coo.data.begin(), coo.data.end(), [](const std::tuple<int, int, value_type>& a, const std::tuple<int, int, value_type>& b) {
      return std::get<0>(a) != std::get<0>(b) ? std::get<0>(a) < std::get<0>(b) : std::get<1>(a) < std::get<1>(b);
    });
    return coo;
  }
};

}  // namespace io
}  // namespace cask
#endif
