// Matrix Market tokeniser (mmio.cpp): host-only, shared with ingest.cu.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/cask_b200.h"

namespace caskb200 {

int fail(int code, const std::string& msg);  // capi.cu

#ifndef CB_TRY
#define CB_TRY(expr)                     \
  do {                                   \
    int _rc = (expr);                    \
    if (_rc != CASK_B200_OK) return _rc; \
  } while (0)
#endif

namespace mm {

struct File {
  std::string type, format, data_type, symmetry;  // the four header words, IO.hpp:39-58
  int64_t n = 0, m = 0, l = 0;                     // the size line
  std::vector<int32_t> rows, cols;                 // as in the file: 1-based
  std::vector<double> vals;
  bool symmetric() const { return symmetry == "symmetric"; }
  bool matrix() const { return type == "matrix"; }
};

int read_header(const std::string& path, File* out);             // io::readHeader + the size line
int read_coo(const std::string& path, File* out);                // the token stream of io::readDokMatrix
int read_vector(const std::string& path, std::vector<double>* v);  // io::readVector

}  // namespace mm
}  // namespace caskb200
