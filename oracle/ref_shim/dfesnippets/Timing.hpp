// Stand-in for lib/dfe-snippets Timing.hpp (submodule is empty in /root/reference).
// Call sites: Spmv.cpp:285 (clock_diff), IO.hpp:324 (print_clock_diff).
#pragma once
#include <chrono>
#include <iostream>
#include <string>
namespace dfesnippets { namespace timing {
template <typename TP> inline double clock_diff(TP start) {
  return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
}
template <typename TP> inline void print_clock_diff(std::string msg, TP start) {
  (void)msg; (void)start;
}
}}
