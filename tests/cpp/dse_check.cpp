// CPU check of the selector's traffic model and of the dse_out.json writer (cask_b200/host/include/Dse.hpp): the
// plan statistics are fabricated (no GPU), the architectures are constructed but never preprocessed.
#include <iostream>

#include "Dse.hpp"

int main(int argc, char** argv) {
  using namespace cask;
  model::B200Model dm;
  // C2 as the GPU partitioner lays it out: every slice staged ELL of width 5
  cask_b200_plan_stats c2{};
  c2.n = c2.m = 16777216; c2.nnz = 83869696; c2.slice_rows = 1024; c2.num_slices = 16384; c2.slices_staged_ell = 16384;
  c2.ell_padded_entries = 5LL * 16777216; c2.ell_nnz = c2.nnz; c2.device_bytes = 10LL * c2.ell_padded_entries;
  const dse::Estimate e2 = dse::estimate(c2, dm);
  // an R-MAT-like plan: everything gather CSR, x (268 MB) larger than L2
  cask_b200_plan_stats c3{};
  c3.n = c3.m = 33554432; c3.nnz = 500000000; c3.slice_rows = 1024; c3.num_slices = 32768; c3.slices_gather_csr = 32768;
  c3.device_bytes = 12LL * c3.nnz;
  const dse::Estimate e3 = dse::estimate(c3, dm);
  std::cout.precision(17);
  std::cout << e2.bytes << " " << e2.seconds << " " << e2.gflops << " " << e2.ellFill << " " << e3.bytes << " " << e3.seconds << " "
            << e3.gflops << std::endl;
  dse::Candidate a, b;
  a.arch = std::make_shared<spmv::SkipEmptyRowsSpmv>(8192, 16, 1, 16777216, 1); a.stats = c2; a.est = e2;
  b.arch = std::make_shared<spmv::SkipEmptyRowsSpmv>(2048, 16, 1, 33554432, 1); b.stats = c3; b.est = e3;
  std::cout << (dse::better(a, b) ? "a" : "b") << std::endl << a.to_string() << std::endl;
  dse::write_dse_results({dse::DseResult{a, {"/m/poisson.mtx", "/m/other.mtx"}}, dse::DseResult{b, {"/m/rmat.mtx"}}}, 1.5, dm,
                         argc > 1 ? argv[1] : "dse_out.json");
  dse::DseParameters p;
  for (int v : p.cacheSize.values()) std::cout << v << " ";
  std::cout << std::endl;
  return 0;
}
