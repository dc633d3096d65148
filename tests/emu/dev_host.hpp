// TEST INFRASTRUCTURE ONLY.  Host emulation of the `dev::` interface of cask_b200/csrc/devlogic.cuh, so that the
// device-logic sources (cask_b200/csrc/*_logic.inl: per-element functors + their orchestration) can be executed and
// checked against the oracle on a machine without a GPU.  Compiled only into tests/emu/libcask_emu.so by the test
// suite; never part of libcask_b200.so, which has no CPU path of any kind.
//
// for_each() visits the indices in a scrambled order (a fixed odd stride modulo n, several "threads" interleaved) -
// a functor that silently depends on ascending execution order fails here instead of on the GPU.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/cask_b200.h"

#define CB_DEV inline
#define CB_TRY(expr)                     \
  do {                                   \
    int _rc = (expr);                    \
    if (_rc != CASK_B200_OK) return _rc; \
  } while (0)

namespace caskb200 {

inline std::string& emu_last_error() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const std::string& msg) {
  emu_last_error() = msg;
  return code;
}

namespace dev {

struct Exec {
  int stream = 0;
  int64_t* launches = nullptr;
  int sm_count = 148;
  int order = 1;  // 0 ascending, 1 scrambled, 2 descending
};

inline int64_t& live_allocations() {
  static int64_t n = 0;
  return n;
}
inline int alloc(void** p, size_t bytes) {
  *p = std::malloc(bytes ? bytes : 16);
  if (!*p) return fail(CASK_B200_ERR_CUDA, "emu: out of memory");
  std::memset(*p, 0xA5, bytes ? bytes : 16);  // device memory is not zeroed either
  ++live_allocations();
  return CASK_B200_OK;
}
inline void release(void* p) {
  if (p) { std::free(p); --live_allocations(); }
}
inline int upload(Exec&, void* d, const void* h, size_t bytes) { if (bytes) std::memcpy(d, h, bytes); return CASK_B200_OK; }
inline int download(Exec&, void* h, const void* d, size_t bytes) { if (bytes) std::memcpy(h, d, bytes); return CASK_B200_OK; }
inline int copy(Exec&, void* d, const void* s, size_t bytes) { if (bytes) std::memmove(d, s, bytes); return CASK_B200_OK; }
inline int zero(Exec&, void* d, size_t bytes) { if (bytes) std::memset(d, 0, bytes); return CASK_B200_OK; }
inline int sync(Exec&) { return CASK_B200_OK; }

template <class F>
int for_each(Exec& ex, int64_t n, const F& f) {
  if (n <= 0) return CASK_B200_OK;
  if (ex.launches) ++*ex.launches;
  if (ex.order == 0) { for (int64_t i = 0; i < n; i++) f(i); return CASK_B200_OK; }
  if (ex.order == 2) { for (int64_t i = n - 1; i >= 0; i--) f(i); return CASK_B200_OK; }
  // a permutation of [0, n): i -> (start + i * stride) mod n with stride coprime to n
  int64_t stride = 7919 % n;
  while (stride < 1 || std::gcd(stride, n) != 1) stride = (stride + 1) % (n + 1) ? (stride + 1) : 1;
  int64_t pos = (n / 3) % n;
  for (int64_t i = 0; i < n; i++) {
    f(pos);
    pos += stride;
    if (pos >= n) pos -= n;
  }
  return CASK_B200_OK;
}

inline int sort_pairs_u64_u32(Exec& ex, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                              uint32_t* vals_out, int64_t n, int end_bit) {
  if (n <= 0) return CASK_B200_OK;
  if (n > INT32_MAX) return fail(CASK_B200_ERR_UNSUPPORTED, "sort: more than 2^31-1 entries");
  const uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
  std::vector<int64_t> perm((size_t)n);
  std::iota(perm.begin(), perm.end(), (int64_t)0);
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
  for (int64_t i = 0; i < n; i++) { keys_out[i] = keys_in[perm[i]]; vals_out[i] = vals_in[perm[i]]; }
  if (ex.launches) *ex.launches += 8;
  return CASK_B200_OK;
}

inline int inclusive_sum_i32(Exec& ex, const int32_t* in, int32_t* out, int64_t n) {
  int32_t run = 0;
  for (int64_t i = 0; i < n; i++) { run += in[i]; out[i] = run; }
  if (ex.launches) ++*ex.launches;
  return CASK_B200_OK;
}

inline double mul_rn(double a, double b) { return a * b; }  // built with -ffp-contract=off
inline double ld_l2(const double* p) { return *p; }
inline double sub_rn(double a, double b) { return a - b; }
inline void atomic_or_i32(int32_t* p, int32_t v) { *p |= v; }
inline void atomic_add_i64(int64_t* p, int64_t v) { *p += v; }
inline void atomic_min_i64(int64_t* p, int64_t v) { if (v < *p) *p = v; }

}  // namespace dev
}  // namespace caskb200
