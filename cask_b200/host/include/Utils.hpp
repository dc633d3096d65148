// cask::utils — host helpers with the reference's names (src/runtime/Utils.hpp:15-112): Timer,
// align, size_bytes, ceilDivide, logResult.  The DSE parameter ranges of that header are out of scope.
#ifndef CASK_B200_HOST_UTILS_HPP
#define CASK_B200_HOST_UTILS_HPP
#include <chrono>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace cask {
namespace utils {

// tic(key) ... toc(key): named wall-clock intervals (Utils.hpp:15-50)
class Timer {
  using clock = std::chrono::high_resolution_clock;
  std::map<std::string, clock::time_point> running_;
  std::map<std::string, std::chrono::duration<double>> done_;

 public:
  void tic(const std::string& name) { running_[name] = clock::now(); }
  std::chrono::duration<double> toc(const std::string& name) {
    auto it = running_.find(name);
    if (it == running_.end()) throw std::invalid_argument("No previous tic() with " + name);
    done_[name] = clock::now() - it->second;
    running_.erase(it);
    return done_[name];
  }
  std::chrono::duration<double> get(const std::string& name) {
    auto it = done_.find(name);
    if (it == done_.end()) throw std::invalid_argument("No previous tic()/toc() with " + name);
    return it->second;
  }
};

// zero-pad v until its byte size is a multiple of widthInBytes; gives up after widthInBytes/sizeof(T)
// pushes, like the reference (Utils.hpp:61-68)
template <typename T>
void align(std::vector<T>& v, int widthInBytes) {
  for (int left = widthInBytes / (int)sizeof(T); left > 0 && (v.size() * sizeof(T)) % widthInBytes != 0; --left)
    v.push_back(T{});
}
inline int align(int bytes, int to) { return bytes % to ? (bytes / to + 1) * to : bytes; }
template <typename T>
long size_bytes(const std::vector<T>& v) { return (long)(sizeof(T) * v.size()); }
inline int ceilDivide(int a, int b) {
  if (a < 0 || b < 0) throw std::invalid_argument("ceilDivide: arguments must be positive");
  return a / b + (a % b != 0);
}

// "Result  <key>=v1,v2,..." lines scraped by the reference's frontend (cask.py:360-368, Utils.hpp:88-112)
template <typename U>
void logResult(const std::string& key, const std::vector<U>& vals) {
  std::cout << "Result  " << key << "=";
  for (const auto& v : vals) std::cout << v << ",";
  std::cout << std::endl;
}
template <typename U>
void logResult(const std::string& key, U val) {
  std::cout << "Result  " << key << "=" << val << "," << std::endl;
}

}  // namespace utils
}  // namespace cask
#endif
