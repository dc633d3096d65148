// test_spmv <matrix.mtx> [implId] — the reference's integration harness (test/test_spmv.cpp:19-83) against
// the B200 path: same flow, same output lines ("Param MatrixPath", "Test passed!", "All tests passed!"),
// same exit status. Eigen is not available, so the expected value `*eigenMatrix * ex` (test_spmv.cpp:45-47)
// is a plain row-major product of the sorted COO triplets (duplicates summed, as setFromTriplets does).
#include <cmath>
#include <iomanip>
#include <iostream>
#include <string>

#include "GeneratedImplSupport.hpp"
#include "IO.hpp"
#include "Spmv.hpp"

namespace {

// dfesnippets::numeric_utils::almost_equal(got, exp, 1E-8, 1E-11) at test/test_utils.hpp:36 (submodule
// absent; relative then absolute tolerance, meaning inferred from the magnitudes)
bool almost_equal(double a, double b, double rel, double abs_tol) {
  const double diff = std::fabs(a - b);
  return diff <= abs_tol || diff <= rel * std::fmax(std::fabs(a), std::fabs(b));
}

int test(const std::string& path, int implId) {
  std::cout << "File: " << path << std::endl;
  cask::io::MmReader<double> reader(path);
  auto coo = reader.mmreadMatrix(path);
  const int cols = coo.m;
  std::cout << "Param MatrixPath " << path << std::endl;

  cask::Vector x(cols);
  for (int i = 0; i < cols; i++) x[i] = (double)i * 0.25;

  cask::runtime::SpmvImplementationLoader implLoader;
  cask::runtime::GeneratedSpmvImplementation* deviceImpl =
      implId == -1 ? implLoader.architectureWithParams(coo.n) : implLoader.architectureWithId(implId);
  if (!deviceImpl) {
    std::cerr << "no design supports " << coo.n << " rows" << std::endl;
    return 1;
  }
  cask::spmv::Spmv a(*deviceImpl);
  auto csrMatrix = cask::io::readMatrix(path);
  a.preprocess(csrMatrix);
  cask::Vector got = a.spmv(x);

  std::vector<double> exp(coo.n, 0.0);
  for (const auto& t : coo.data) exp[std::get<0>(t)] += std::get<2>(t) * x[std::get<1>(t)];

  int mismatches = 0;
  for (int i = 0; i < coo.n; i++)
    if (!almost_equal(got[i], exp[i], 1E-8, 1E-11)) {
      if (mismatches++ < 20)
        std::cerr << std::fixed << std::setprecision(10) << "At " << i << " got: " << got[i] << " exp: " << exp[i] << std::endl;
    }
  if (!mismatches) {
    std::cout << "Test passed!" << std::endl;
    return 0;
  }
  std::cout << "Test failed: " << mismatches << " mismatches " << std::endl;
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  std::cout << "Program arguments:" << std::endl;
  for (int i = 0; i < argc; i++) std::cout << "   " << argv[i] << std::endl;
  if (argc < 2 || argc > 3) return 1;
  int status = -1;
  try {
    status = test(argv[1], argc == 3 ? std::stoi(argv[2]) : -1);
  } catch (std::exception& e) {
    std::cerr << "exception: " << e.what() << std::endl;
    status = 2;
  }
  std::cout << (status == 0 ? "All tests passed!" : "Tests failed!") << std::endl;
  return status;
}
