// The reference's DSE driver (src/main.cpp) for a B200: for every matrix, preprocess each candidate design on the GPU,
// score it with the traffic model of Dse.hpp, print the reference's table and write dse_out.json.
//   usage: cask_dse [--out dse_out.json] [--cache-min N] [--cache-max N] [--hbm-gbs X] <matrix.mtx> [...]
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "Dse.hpp"

int main(int argc, char** argv) {
  cask::dse::Benchmark bench;
  cask::dse::DseParameters params;
  cask::model::B200Model dm;
  std::string out = "dse_out.json";
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { std::cerr << a << " needs a value" << std::endl; std::exit(2); } return argv[++i]; };
    if (a == "--out") out = next();
    else if (a == "--cache-min") params.cacheSize.start = std::atoi(next());
    else if (a == "--cache-max") params.cacheSize.end = std::atoi(next());
    else if (a == "--hbm-gbs") dm.hbmGBs = std::atof(next());
    else if (a == "--gflops-only") params.gflopsOnly = true;
    else bench.add_matrix_path(a);
  }
  if (bench.get_benchmark_size() == 0) {
    std::cerr << "usage: cask_dse [--out file] [--cache-min N] [--cache-max N] [--hbm-gbs X] <matrix.mtx> [...]" << std::endl;
    return 2;
  }
  try {
    const auto start = std::chrono::high_resolution_clock::now();
    cask::dse::SparkDse dse;
    auto results = dse.run(bench, params, dm);
    const double took = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
    cask::dse::write_dse_results(results, took, dm, out);
    std::cout << "Wrote " << out << " (" << results.size() << " best architecture(s), took " << took << " s)" << std::endl;
  } catch (std::exception& e) {
    std::cerr << "cask_dse: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
