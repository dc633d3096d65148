// cask::dse — the architecture selector, B200 edition (src/runtime/Dse.hpp, Dse.cpp:32-74, main.cpp:60-117).
// The reference sweeps (numPipes, inputWidth, cacheSize, numControllers), runs Spmv::preprocess for each candidate,
// scores it with a cycle / resource model of the FPGA and prints one line per candidate plus a dse_out.json.  On a
// B200 the build parameters that matter are the x-cache budget of a slice (cache_size) and the ELL fill threshold;
// the score is a traffic model of the plan the GPU partitioner actually produced for the candidate:
//
//   bytes   = 10 * ell_padded_entries                       staged-ELL slices: 8 B value + 2 B cache index per stored entry
//           + 12 * nnz_csr + 4 * rows_csr                   gather-CSR slices: value + 32-bit column, row pointers
//           + 8 * m + 8 * n                                 x read once, y written once
//           + 32 * nnz_csr * max(0, 1 - L2 / (8 * m))       gathers of x that miss the 126 MB L2 cost one 32 B sector each
//   seconds = bytes / HBM bandwidth;  GFLOP/s = 2 nnz / seconds;  "clock cycles" = seconds * SM clock
//
// Same entry points and output shape as the reference (Benchmark, DseParameters, SparkDse::run, DseResult, the table
// header of Dse.cpp:116 with B200 columns, write_dse_results -> dse_out.json with every value written as a string, as
// boost::property_tree does).
#ifndef CASK_B200_HOST_DSE_HPP
#define CASK_B200_HOST_DSE_HPP
#include <algorithm>
#include <chrono>
#include <ctime>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "IO.hpp"
#include "Spmv.hpp"
#include "Utils.hpp"

namespace cask {
namespace model {
struct B200Model {  // stands where DeviceModel / Max4Model stand in the reference (src/runtime/Model.hpp)
  double hbmGBs = 6553.0;      // measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json); nominal 8000
  double l2Bytes = 126.0e6;
  double smClockHz = 1.965e9;
  int sms = 148;
  std::string getId() const { return "B200"; }
};
}  // namespace model

namespace dse {

class Benchmark {  // Dse.hpp:15-34
  std::vector<std::string> paths;

 public:
  std::string get_matrix_path(int id) const {
    if (id >= 0 && id < (int)paths.size()) return paths[id];
    std::stringstream ss;
    ss << "Benchmark::Index out of range " << id;
    throw std::invalid_argument(ss.str());
  }
  void add_matrix_path(std::string path) { paths.push_back(path); }
  int get_benchmark_size() const { return (int)paths.size(); }
};

struct Range {  // utils::Parameter<int> (Utils.hpp): start, end, step - here multiplicative for cacheSize
  int start, end, step;
  bool geometric;
  std::vector<int> values() const {
    std::vector<int> v;
    for (int x = start; x <= end; x = geometric ? x * step : x + step) { v.push_back(x); if (step <= (geometric ? 1 : 0)) break; }
    return v;
  }
};

class DseParameters {  // Dse.hpp:45-52
 public:
  bool gflopsOnly = false;
  Range numPipes{1, 1, 1, false};
  Range inputWidth{16, 16, 1, false};
  Range cacheSize{2048, 16384, 2, true};
  Range numControllers{1, 1, 1, false};
};

struct Estimate {
  double bytes = 0, seconds = 0, gflops = 0, clockCycles = 0, ellFill = 0, memoryBandwidthGBs = 0;
};

// The model itself lives behind the ABI (cask_b200_plan_estimate): one definition for this selector, for bench-side
// validation (profiles/dse_validate.py) and for callers in other languages.
inline Estimate estimate(const cask_b200_plan_stats& s, const model::B200Model& dm) {
  Estimate e;
  if (cask_b200_plan_estimate(&s, dm.hbmGBs, dm.l2Bytes, dm.smClockHz, dm.sms, &e.bytes, &e.seconds) != CASK_B200_OK)
    throw std::runtime_error(cask_b200_last_error());
  e.gflops = e.seconds > 0 ? 2.0 * (double)s.nnz / e.seconds / 1e9 : 0.0;
  e.clockCycles = e.seconds * dm.smClockHz;
  e.ellFill = s.ell_padded_entries ? (double)s.ell_nnz / (double)s.ell_padded_entries : 0.0;
  e.memoryBandwidthGBs = dm.hbmGBs;  // the denominator of the byte term (stencil plans are bound by it)
  return e;
}

struct Candidate {
  std::shared_ptr<spmv::Spmv> arch;
  cask_b200_plan_stats stats;
  Estimate est;
  std::string to_string() const {  // Spmv::to_string (Spmv.hpp:176-186) + the B200 columns
    std::stringstream s;
    s << arch->get_name() << " " << arch->impl.cache_size << " " << arch->impl.input_width << " " << arch->impl.num_pipes << " "
      << arch->impl.num_controllers << " " << est.clockCycles << " " << est.gflops << " " << stats.slices_staged_ell << " "
      << stats.slices_gather_csr << " " << est.ellFill << " " << stats.device_bytes << " " << est.memoryBandwidthGBs;
    return s.str();
  }
};

struct DseResult {  // Dse.hpp: best architecture + the matrices it is best for
  Candidate best;
  std::vector<std::string> matrices;
};

inline bool better(const Candidate& a, const Candidate& b) {  // Dse.cpp:12-30: higher GFLOP/s, then the smaller footprint
  return a.est.gflops > b.est.gflops || (a.est.gflops == b.est.gflops && a.stats.device_bytes < b.stats.device_bytes);
}

class SparkDse {
 public:
  std::vector<DseResult> run(const Benchmark& benchmark, const DseParameters& params, const model::B200Model& dm) {
    std::vector<DseResult> results;
    for (int i = 0; i < benchmark.get_benchmark_size(); i++) {
      const std::string path = benchmark.get_matrix_path(i);
      const std::size_t pos = path.find_last_of("/");
      const std::string basename = pos == std::string::npos ? path : path.substr(pos);
      std::cout << basename << std::endl;
      const auto start = std::chrono::high_resolution_clock::now();
      CsrMatrix matrix = io::readMatrix(path);
      std::cout << "Reading took: " << std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count()
                << std::endl;
      std::cout << "File Architecture CacheSize InputWidth NumPipes NumControllers EstClockCycles EstGflops StagedSlices "
                   "GatherSlices EllFill DeviceBytes MemBandwidth Observation" << std::endl;
      bool have = false;
      Candidate best;
      for (int pipes : params.numPipes.values())
        for (int width : params.inputWidth.values())
          for (int cache : params.cacheSize.values())
            for (int ctl : params.numControllers.values()) {
              Candidate c;
              c.arch = std::make_shared<spmv::SkipEmptyRowsSpmv>(cache, width, pipes, matrix.n, ctl);  // Dse.cpp:41-45
              if (!c.arch->isValid()) continue;
              c.arch->preprocess(matrix);
              c.stats = c.arch->getPlanStats();
              c.est = estimate(c.stats, dm);
              std::cout << basename << " " << c.to_string() << std::endl;
              if (!have || better(c, best)) { best = c; have = true; }
            }
      if (!have) continue;
      std::cout << basename << " ";
      if (params.gflopsOnly) std::cout << best.est.gflops << best.est.clockCycles;
      else std::cout << best.to_string();
      std::cout << " Best " << std::endl;
      // the frontend's scrapers (cask.py:360-368)
      utils::logResult("Estimated gflops", best.est.gflops);
      utils::logResult("Cache size", best.arch->impl.cache_size);
      bool merged = false;
      for (auto& r : results)
        if (r.best.arch->impl == best.arch->impl && r.best.arch->get_name() == best.arch->get_name()) {
          r.matrices.push_back(path);  // main.cpp groups matrices by best architecture
          merged = true;
        }
      if (!merged) results.push_back(DseResult{best, {path}});
    }
    return results;
  }
};

// main.cpp:81-117: dse_out.json.  boost::property_tree writes every leaf as a string; kept.
inline void write_dse_results(const std::vector<DseResult>& results, double took, const model::B200Model& dm,
                              const std::string& path = "dse_out.json") {
  auto q = [](const std::string& s) { return "\"" + s + "\""; };
  auto num = [&](auto v) { std::stringstream ss; ss.precision(15); ss << v; return q(ss.str()); };  // ints stay ints
  std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
  std::string date = std::ctime(&now);
  while (!date.empty() && date.back() == '\n') date.pop_back();
  std::ofstream f(path);
  f << "{\n    \"date\": " << q(date + "\\n") << ",\n    \"took\": " << num(took) << ",\n    \"device\": " << q(dm.getId())
    << ",\n    \"best_architectures\": [\n";
  for (size_t i = 0; i < results.size(); i++) {
    const Candidate& c = results[i].best;
    f << "        {\n            \"name\": " << q(c.arch->get_name()) << ",\n            \"estimated_gflops\": " << num(c.est.gflops)
      << ",\n            \"estimated_clock_cycles\": " << num(c.est.clockCycles) << ",\n            \"architecture_params\": {\n"
      << "                \"num_pipes\": " << num(c.arch->impl.num_pipes) << ",\n                \"cache_size\": " << num(c.arch->impl.cache_size)
      << ",\n                \"input_width\": " << num(c.arch->impl.input_width) << ",\n                \"max_rows\": " << num(c.arch->impl.max_rows)
      << ",\n                \"num_controllers\": " << num(c.arch->impl.num_controllers) << "\n            },\n"
      << "            \"estimated_impl_params\": {\n                \"memory_bandwidth\": " << num(c.est.memoryBandwidthGBs)
      << ",\n                \"device_bytes\": " << num(c.stats.device_bytes) << ",\n                \"slices_staged_ell\": "
      << num(c.stats.slices_staged_ell) << ",\n                \"slices_gather_csr\": " << num(c.stats.slices_gather_csr)
      << ",\n                \"ell_fill\": " << num(c.est.ellFill) << ",\n                \"estimated_bytes_per_spmv\": " << num(c.est.bytes)
      << "\n            },\n            \"matrices\": [\n";
    for (size_t k = 0; k < results[i].matrices.size(); k++)
      f << "                " << q(results[i].matrices[k]) << (k + 1 < results[i].matrices.size() ? ",\n" : "\n");
    f << "            ]\n        }" << (i + 1 < results.size() ? ",\n" : "\n");
  }
  f << "    ]\n}\n";
}

}  // namespace dse
}  // namespace cask
#endif
