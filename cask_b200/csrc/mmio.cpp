// Matrix Market text -> coordinate arrays, host side (no CUDA in this file).
// Replaces the tokenising half of io::readHeader / readDokMatrix / readVector, src/runtime/IO.hpp:60-148: the file is
// mapped, split at token boundaries and parsed by all host cores; what the reference then does with one hash-map
// insertion per entry is done on the GPU by ingest.cu.
//
// Accepted exactly as the reference accepts it:
//   * first line must match "%%MatrixMarket (matrix|array) (coordinate|array) (real|integer) (symmetric|general)"
//     with single spaces and nothing after it (std::regex_match at IO.hpp:66-70) - so pattern / complex / hermitian /
//     skew-symmetric files and CRLF line ends are rejected with the reference's message;
//   * the size line is the first line that does not start with '%' (IO.hpp:128-131);
//   * entries are whitespace-separated tokens, three per entry, NOT tied to lines (operator>> at IO.hpp:143);
//     tokens after the L-th entry are ignored.
// Stricter than the reference where the reference would read garbage: a token that is not a number, an index that does
// not fit 32 bits, or a file that ends before L entries is an error here (the reference carries on with whatever
// operator>> left in its variables).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mmio.hpp"

namespace caskb200 {
namespace mm {

namespace {

struct Mapped {
  const char* p = nullptr;
  size_t len = 0;
  int fd = -1;
  ~Mapped() {
    if (p && len) munmap(const_cast<char*>(p), len);
    if (fd >= 0) close(fd);
  }
  int open(const std::string& path) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "File not found " + path);  // IO.hpp:63-64
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "File not found " + path);
    len = (size_t)st.st_size;
    if (len) {
      void* q = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
      if (q == MAP_FAILED) { len = 0; return fail(CASK_B200_ERR_RUNTIME, "mmap failed for " + path); }
      p = static_cast<const char*>(q);
    }
    return CASK_B200_OK;
  }
};

inline bool is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

bool one_of(const std::string& w, std::initializer_list<const char*> set) {
  for (const char* s : set)
    if (w == s) return true;
  return false;
}

// [b, e) -> int32; the whole token must be consumed
bool parse_i32(const char* b, const char* e, int32_t* out) {
  if (b < e && *b == '+') b++;
  int32_t v = 0;
  const auto r = std::from_chars(b, e, v, 10);
  if (r.ec != std::errc() || r.ptr != e) return false;
  *out = v;
  return true;
}
bool parse_f64(const char* b, const char* e, double* out) {
  if (b < e && *b == '+') b++;
  double v = 0.0;
  const auto r = std::from_chars(b, e, v);
  if (r.ec != std::errc() || r.ptr != e) return false;
  *out = v;
  return true;
}

// positions: header line end, size line [s0, s1), data start
int locate(const Mapped& f, const std::string& path, File* out, size_t* data_begin) {
  const char* p = f.p;
  const size_t len = f.len;
  size_t e = 0;
  while (e < len && p[e] != '\n') e++;
  const std::string first(p, e);
  // "%%MatrixMarket w1 w2 w3 w4", single spaces, whole line
  std::vector<std::string> w;
  size_t a = 0;
  while (true) {
    const size_t sp = first.find(' ', a);
    w.push_back(first.substr(a, sp == std::string::npos ? std::string::npos : sp - a));
    if (sp == std::string::npos) break;
    a = sp + 1;
  }
  const bool ok = w.size() == 5 && w[0] == "%%MatrixMarket" && one_of(w[1], {"matrix", "array"}) &&
                  one_of(w[2], {"coordinate", "array"}) && one_of(w[3], {"real", "integer"}) &&
                  one_of(w[4], {"symmetric", "general"});
  if (!ok) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Not a valid MatrixMarket file in " + path);  // IO.hpp:71
  out->type = w[1]; out->format = w[2]; out->data_type = w[3]; out->symmetry = w[4];
  // the size line: first line (from the top of the file) that does not start with '%'
  size_t ls = 0;
  while (true) {
    if (ls >= len) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "No size line in MatrixMarket file " + path);
    size_t le = ls;
    while (le < len && p[le] != '\n') le++;
    if (le == ls || p[ls] != '%') {
      // parse "N M [L]"
      int64_t dims[3] = {0, 0, 0};
      int got = 0;
      size_t q = ls;
      while (q < le && got < 3) {
        while (q < le && is_space(p[q])) q++;
        size_t t = q;
        while (t < le && !is_space(p[t])) t++;
        if (t == q) break;
        int32_t v;
        if (!parse_i32(p + q, p + t, &v) || v < 0)
          return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Malformed size line in MatrixMarket file " + path);
        dims[got++] = v;
        q = t;
      }
      const int need = out->format == "coordinate" ? 3 : 2;
      if (got < need) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Malformed size line in MatrixMarket file " + path);
      out->n = dims[0]; out->m = dims[1]; out->l = need == 3 ? dims[2] : dims[0] * dims[1];
      *data_begin = le < len ? le + 1 : len;
      return CASK_B200_OK;
    }
    ls = le + 1;
  }
}

struct Chunk { size_t b, e; int64_t tokens, first_token; };

int64_t count_tokens(const char* p, size_t b, size_t e) {
  int64_t n = 0;
  bool in = false;
  for (size_t i = b; i < e; i++) {
    const bool sp = is_space(p[i]);
    if (!sp && !in) n++;
    in = !sp;
  }
  return n;
}

template <class Fn>
void run_parallel(int nthreads, Fn fn) {
  if (nthreads <= 1) { fn(0); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(fn, t);
  fn(0);
  for (auto& t : th) t.join();
}

// Splits [b, e) into chunks that begin and end between tokens, counts tokens per chunk (parallel) and numbers them.
std::vector<Chunk> make_chunks(const char* p, size_t b, size_t e, int* nthreads) {
  const size_t bytes = e - b;
  int T = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
  if (bytes < (1u << 20)) T = 1;
  std::vector<Chunk> ch((size_t)T);
  size_t start = b;
  for (int t = 0; t < T; t++) {
    size_t end = t == T - 1 ? e : std::max(start, b + bytes * (size_t)(t + 1) / (size_t)T);
    while (end < e && !is_space(p[end])) end++;  // never cut a token
    ch[t] = Chunk{start, end, 0, 0};
    start = end;
  }
  run_parallel(T, [&](int t) { ch[t].tokens = count_tokens(p, ch[t].b, ch[t].e); });
  int64_t run = 0;
  for (auto& c : ch) { c.first_token = run; run += c.tokens; }
  *nthreads = T;
  return ch;
}

}  // namespace

int read_header(const std::string& path, File* out) {
  Mapped f;
  CB_TRY(f.open(path));
  size_t data = 0;
  return locate(f, path, out, &data);
}

int read_coo(const std::string& path, File* out) {
  Mapped f;
  CB_TRY(f.open(path));
  size_t data = 0;
  CB_TRY(locate(f, path, out, &data));
  if (out->format != "coordinate")
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Expecting a coordinate MatrixMarket file in " + path);  // IO.hpp:133 (assert)
  const int64_t L = out->l;
  // the entry count of the size line is only a claim: the tokens are counted BEFORE anything is sized from it, so a
  // corrupt header cannot make a tiny file allocate gigabytes
  int T = 1;
  const std::vector<Chunk> ch = make_chunks(f.p, data, f.len, &T);
  const int64_t total = ch.empty() ? 0 : ch.back().first_token + ch.back().tokens;
  if (total < 3 * L)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "MatrixMarket file " + path + " ends after " + std::to_string(total / 3) +
                                                    " of " + std::to_string(L) + " entries");
  out->rows.assign((size_t)L, 0);
  out->cols.assign((size_t)L, 0);
  out->vals.assign((size_t)L, 0.0);
  std::atomic<int64_t> bad{INT64_MAX};
  const char* p = f.p;
  int32_t* rows = out->rows.data();
  int32_t* cols = out->cols.data();
  double* vals = out->vals.data();
  run_parallel(T, [&](int t) {
    int64_t g = ch[t].first_token;
    size_t i = ch[t].b;
    const size_t e = ch[t].e;
    while (i < e && g < 3 * L) {
      while (i < e && is_space(p[i])) i++;
      size_t j = i;
      while (j < e && !is_space(p[j])) j++;
      if (j == i) break;
      const int64_t entry = g / 3;
      const int role = (int)(g % 3);
      bool ok;
      if (role == 0) ok = parse_i32(p + i, p + j, rows + entry);
      else if (role == 1) ok = parse_i32(p + i, p + j, cols + entry);
      else ok = parse_f64(p + i, p + j, vals + entry);
      if (!ok) {
        int64_t cur = bad.load();
        while (entry < cur && !bad.compare_exchange_weak(cur, entry)) {}
      }
      g++;
      i = j;
    }
  });
  if (bad.load() != INT64_MAX)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Malformed entry " + std::to_string(bad.load() + 1) + " in MatrixMarket file " + path);
  return CASK_B200_OK;
}

// io::readVector, IO.hpp:73-115.  Coordinate vectors are NOT rebased by the reference (v[a] = val with the file's
// 1-based a, IO.hpp:99) - kept, with a bounds check where the reference would write past the end.
int read_vector(const std::string& path, std::vector<double>* v) {
  Mapped f;
  CB_TRY(f.open(path));
  File info;
  size_t data = 0;
  CB_TRY(locate(f, path, &info, &data));
  if (info.format != "coordinate" && info.n > (int64_t)(f.len - data))  // an array vector needs >= 2 bytes per entry
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "MatrixMarket file " + path + " is shorter than the " + std::to_string(info.n) +
                                                    " entries its size line claims");
  v->assign((size_t)info.n, 0.0);
  int T = 1;
  const char* p = f.p;
  // vectors are small: single pass
  size_t i = data;
  const size_t e = f.len;
  auto next = [&](const char** b, const char** en) {
    while (i < e && is_space(p[i])) i++;
    size_t j = i;
    while (j < e && !is_space(p[j])) j++;
    *b = p + i; *en = p + j;
    const bool got = j > i;
    i = j;
    return got;
  };
  (void)T;
  const char *b, *en;
  if (info.format == "coordinate") {
    for (int64_t k = 0; k < info.l; k++) {
      int32_t a = 0, c = 0;
      double val = 0;
      if (!next(&b, &en) || !parse_i32(b, en, &a) || !next(&b, &en) || !parse_i32(b, en, &c) || !next(&b, &en) ||
          !parse_f64(b, en, &val))
        return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Malformed entry " + std::to_string(k + 1) + " in MatrixMarket file " + path);
      if (a < 0 || a >= info.n)
        return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Vector index out of range in " + path + " (the reference does not rebase "
                                                        "coordinate vectors, IO.hpp:99)");
      (*v)[(size_t)a] = val;
    }
    return CASK_B200_OK;
  }
  for (int64_t k = 0; k < info.n; k++) {
    double val = 0;
    if (!next(&b, &en) || !parse_f64(b, en, &val))
      return fail(CASK_B200_ERR_INVALID_ARGUMENT, "Malformed entry " + std::to_string(k + 1) + " in MatrixMarket file " + path);
    (*v)[(size_t)k] = val;
  }
  return CASK_B200_OK;
}

}  // namespace mm
}  // namespace caskb200

using namespace caskb200;

namespace {
void fill_info(const mm::File& f, cask_b200_mm_info* info) {
  std::memset(info, 0, sizeof(*info));
  std::strncpy(info->type, f.type.c_str(), sizeof(info->type) - 1);
  std::strncpy(info->format, f.format.c_str(), sizeof(info->format) - 1);
  std::strncpy(info->data_type, f.data_type.c_str(), sizeof(info->data_type) - 1);
  std::strncpy(info->symmetry, f.symmetry.c_str(), sizeof(info->symmetry) - 1);
  info->n = f.n; info->m = f.m; info->entries = f.l;
}
}  // namespace

// No C++ exception may cross the C ABI (std::bad_alloc from a huge size line, std::system_error from std::thread):
// every entry point maps them to an error code.
#define CB_ABI_GUARD_BEGIN try {
#define CB_ABI_GUARD_END(name_)                                                                             \
  } catch (const std::bad_alloc&) {                                                                         \
    return fail(CASK_B200_ERR_RUNTIME, std::string(name_) + ": out of host memory");                          \
  } catch (const std::exception& e) {                                                                       \
    return fail(CASK_B200_ERR_RUNTIME, std::string(name_) + ": " + e.what());                                 \
  } catch (...) {                                                                                           \
    return fail(CASK_B200_ERR_RUNTIME, std::string(name_) + ": unknown exception");                           \
  }

extern "C" {

int cask_b200_mm_read_info(const char* path, cask_b200_mm_info* info) {
  if (!path || !info) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_info: null");
  CB_ABI_GUARD_BEGIN
  mm::File f;
  CB_TRY(mm::read_header(path, &f));
  fill_info(f, info);
  return CASK_B200_OK;
  CB_ABI_GUARD_END("mm_read_info")
}

int cask_b200_mm_read_coo(const char* path, int64_t capacity, int32_t* rows, int32_t* cols, double* vals, int64_t* count) {
  if (!path || !count) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_coo: null");
  CB_ABI_GUARD_BEGIN
  mm::File f;
  CB_TRY(mm::read_coo(path, &f));
  *count = f.l;
  if (f.l > capacity) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_coo: capacity too small");
  if (f.l && (!rows || !cols || !vals)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_coo: null arrays");
  if (f.l) {
    std::memcpy(rows, f.rows.data(), sizeof(int32_t) * (size_t)f.l);
    std::memcpy(cols, f.cols.data(), sizeof(int32_t) * (size_t)f.l);
    std::memcpy(vals, f.vals.data(), sizeof(double) * (size_t)f.l);
  }
  return CASK_B200_OK;
  CB_ABI_GUARD_END("mm_read_coo")
}

int cask_b200_mm_read_vector(const char* path, int64_t capacity, double* out, int64_t* n) {
  if (!path || !n) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_vector: null");
  CB_ABI_GUARD_BEGIN
  std::vector<double> v;
  CB_TRY(mm::read_vector(path, &v));
  *n = (int64_t)v.size();
  if ((int64_t)v.size() > capacity) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_vector: capacity too small");
  if (!v.empty() && !out) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "mm_read_vector: null array");
  if (!v.empty()) std::memcpy(out, v.data(), sizeof(double) * v.size());
  return CASK_B200_OK;
  CB_ABI_GUARD_END("mm_read_vector")
}

}  // extern "C"
