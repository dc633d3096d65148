// cask::benchmark::printSummary — the JSON-ish record CgTest writes (src/runtime/Benchmark.hpp:54-71).
#ifndef CASK_B200_HOST_BENCHMARK_HPP
#define CASK_B200_HOST_BENCHMARK_HPP
#include <iostream>
#include <sstream>
#include <string>

namespace cask {
namespace benchmark {

template <typename T>
std::string json(const std::string& key, T value, bool comma = true) {
  std::stringstream ss;
  ss << "\"" << key << "\":\"" << value << "\"" << (comma ? "," : "");
  return ss.str();
}

inline void printSummary(double setupSeconds, int iterations, double solveSeconds, double estimatedError,
                         double solutionVersusExpNorm, double benchmarkRepetitions, std::ostream& os = std::cout) {
  os << "{" << json("setup took", setupSeconds) << json("iterations", iterations) << json("solve took", solveSeconds)
     << json("estimated error", estimatedError) << json("error", solutionVersusExpNorm)
     << json("bench repetitions", benchmarkRepetitions, false) << "}";
}

}  // namespace benchmark
}  // namespace cask
#endif
