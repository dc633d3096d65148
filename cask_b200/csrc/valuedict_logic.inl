// Value dictionaries for staged-ELL slices ("coded ELL"): the values of a slice are replaced by 8-bit codes into a
// per-slice table of its distinct fp64 values, so a stored nonzero costs 3 bytes (code + 16-bit x-cache position)
// instead of 10.  The SpMV looks the value up in shared memory and multiplies the very same double, so y stays
// bit-identical to the uncoded path (and to CsrMatrix::dot, src/runtime/SparseMatrix.hpp:255-264, on staged slices).
// Constant-coefficient stencils (BASELINE configs[1], [3], [4]) have 2-3 distinct values per slice; a matrix with more
// than kMaxEntries distinct values in any slice keeps the uncoded format - the caller checks *overflow.
//
// Included by plan.cu (after devlogic.cuh) and by tests/emu/emu.cpp (after tests/emu/dev_host.hpp): only the `dev::`
// interface is used, so the functor body below is executed against a numpy model on a machine without a GPU
// (tests/test_valuedict_emu.py).
//
// One thread per slice: the slice's width * 1024 stored values (padding included, written by plan_fill_kernel) are
// scanned in storage order; the table lists the distinct BIT PATTERNS in order of first appearance (so +0.0 and -0.0,
// or two NaN payloads, are different entries) and is deterministic.  This is a preprocessing step; it reads the value
// array twice.
#include <vector>

namespace caskb200 {
namespace valuedict {

constexpr int kMaxEntries = 256;   // 8-bit codes
constexpr int kStride = 256;       // doubles reserved per slice in the table array (slot s of slice q: q * kStride + s)

struct BuildSlice {
  const int64_t* val_off;    // per listed slice: entry offset of the slice in the ELL arrays
  const int32_t* width;      // per listed slice: ELL width
  int32_t slice_rows;        // 1024
  const double* vals;        // ELL values
  double* table;             // kStride doubles per listed slice; unused slots are written as 0.0
  uint8_t* codes;            // same indexing as vals
  int32_t* ndict;            // per listed slice: number of table entries (0 if the slice overflowed)
  int32_t* overflow;         // set to 1 if some slice holds more than kMaxEntries distinct values
  CB_DEV void operator()(int64_t q) const {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(vals) + val_off[q];
    uint8_t* dst = codes + val_off[q];
    const int64_t total = (int64_t)width[q] * slice_rows;
    unsigned long long dict[kMaxEntries];
    int32_t n = 0;
    unsigned long long last = 0;
    int32_t last_code = -1;
    bool over = false;
    for (int64_t e = 0; e < total; e++) {
      const unsigned long long b = src[e];
      if (last_code < 0 || b != last) {
        int32_t c = 0;
        while (c < n && dict[c] != b) c++;
        if (c == n) {
          if (n == kMaxEntries) { over = true; break; }
          dict[n++] = b;
        }
        last = b;
        last_code = c;
      }
      dst[e] = (uint8_t)last_code;
    }
    if (over) {
      dev::atomic_or_i32(overflow, 1);
      n = 0;
    }
    ndict[q] = n;
    unsigned long long* out = reinterpret_cast<unsigned long long*>(table) + q * kStride;
    for (int32_t c = 0; c < kStride; c++) out[c] = c < n ? dict[c] : 0ull;
  }
};

// d_val_off / d_width: device arrays with one entry per listed slice.  On return *overflow tells whether the codes are
// usable (0) or the matrix must stay uncoded (1); *max_entries is the largest table.
inline int build(dev::Exec& ex, int64_t nlisted, const int64_t* d_val_off, const int32_t* d_width, int32_t slice_rows,
                 const double* d_vals, double* d_table, uint8_t* d_codes, int32_t* d_ndict, int32_t* overflow,
                 int32_t* max_entries) {
  *overflow = 0;
  *max_entries = 0;
  if (nlisted <= 0) return CASK_B200_OK;
  int32_t* d_over = nullptr;
  CB_TRY(dev::alloc((void**)&d_over, 16));
  int rc = dev::zero(ex, d_over, 16);
  BuildSlice f{d_val_off, d_width, slice_rows, d_vals, d_table, d_codes, d_ndict, d_over};
  if (rc == CASK_B200_OK) rc = dev::for_each(ex, nlisted, f);
  if (rc == CASK_B200_OK) rc = dev::download(ex, overflow, d_over, sizeof(int32_t));
  dev::release(d_over);
  CB_TRY(rc);
  std::vector<int32_t> nd((size_t)nlisted);
  CB_TRY(dev::download(ex, nd.data(), d_ndict, sizeof(int32_t) * (size_t)nlisted));
  for (int32_t v : nd) *max_entries = v > *max_entries ? v : *max_entries;
  return CASK_B200_OK;
}

// ---- pair dictionaries ("pair-coded ELL", option value_dict = 2) ----------------------------------------------
// One step further: the 8-bit code names a (value, x-cache displacement) PAIR, where the displacement of a stored entry
// is its x-cache position minus its row number inside the slice.  Along a diagonal of a banded matrix the position moves
// with the row, so the displacement is constant: a constant-coefficient stencil needs one pair per stencil point (2D
// 5-point: 5, 3D 27-point: 27) whatever the slice, the 16-bit index stream disappears and a stored nonzero costs ONE
// byte.  The kernel rebuilds position = displacement + row and multiplies the very same double by the very same x entry
// in the same order: y stays bit-identical.
// Code 0 is reserved for padding entries (written by plan_fill_kernel as value +0.0 at index 0; a real entry never has
// index 0 or 1): its pair is (+0.0, -slice_rows), i.e. position = row - slice_rows - the kernel keeps slice_rows zeros in
// front of every x cache, so padding needs no special case in the inner loop.  Real pairs take codes 1..255 in order of
// first appearance.  A slice with more than 255 distinct pairs raises *overflow and the caller falls back to the value
// dictionaries above.  Table entries are 16-byte records {value bits, displacement in BYTES (x 8), 0}: one shared-memory
// address per code serves both lookups.
// Storage order inside a slice (plan.cu): entry e = (k * T + t) * 4 + j, row = j * T + t, T = slice_rows / 4.
constexpr int kMaxPairs = 255;
constexpr int kPairStride = 256;    // records reserved per slice (slot c of slice q: q * kPairStride + c)

struct PairEntry {
  unsigned long long value_bits;
  int32_t disp8;   // (x-cache position - row) * 8
  int32_t zero_;
};
static_assert(sizeof(PairEntry) == 16, "pair records are 16 bytes");

struct BuildPairs {
  const int64_t* val_off;
  const int32_t* width;
  int32_t slice_rows;
  const double* vals;
  const uint16_t* idx;       // ELL x-cache positions, same indexing as vals
  PairEntry* table;          // kPairStride records per listed slice; unused slots are zero
  uint8_t* codes;
  int32_t* npairs;           // per listed slice: table length including slot 0 (0 if the slice is empty or overflowed)
  int32_t* overflow;
  CB_DEV void operator()(int64_t q) const {
    const unsigned long long* sv = reinterpret_cast<const unsigned long long*>(vals) + val_off[q];
    const uint16_t* si = idx + val_off[q];
    uint8_t* dst = codes + val_off[q];
    const int64_t total = (int64_t)width[q] * slice_rows;
    const int32_t T = slice_rows / 4;
    unsigned long long dv[kMaxPairs + 1];
    int32_t dd[kMaxPairs + 1];
    dv[0] = 0ull;
    dd[0] = -slice_rows;
    int32_t n = total > 0 ? 1 : 0;
    unsigned long long last_v = 0;
    int32_t last_d = 0;
    int32_t last_code = -1;
    bool over = false;
    int32_t rem = 0;  // e % slice_rows
    for (int64_t e = 0; e < total; e++, rem = rem + 1 == slice_rows ? 0 : rem + 1) {
      const int32_t row = (rem & 3) * T + (rem >> 2);
      const int32_t pos = si[e];
      if (pos == 0) {  // padding
        dst[e] = 0;
        continue;
      }
      const unsigned long long b = sv[e];
      const int32_t d = pos - row;
      if (last_code < 0 || b != last_v || d != last_d) {
        int32_t c = 1;
        while (c < n && !(dv[c] == b && dd[c] == d)) c++;
        if (c == n) {
          if (n == kMaxPairs + 1) { over = true; break; }
          dv[n] = b;
          dd[n] = d;
          n++;
        }
        last_v = b;
        last_d = d;
        last_code = c;
      }
      dst[e] = (uint8_t)last_code;
    }
    if (over) {
      dev::atomic_or_i32(overflow, 1);
      n = 0;
    }
    npairs[q] = n;
    PairEntry* out = table + q * kPairStride;
    for (int32_t c = 0; c < kPairStride; c++) {
      PairEntry r;
      r.value_bits = c < n ? dv[c] : 0ull;
      r.disp8 = c < n ? dd[c] * 8 : 0;
      r.zero_ = 0;
      out[c] = r;
    }
  }
};

inline int build_pairs(dev::Exec& ex, int64_t nlisted, const int64_t* d_val_off, const int32_t* d_width, int32_t slice_rows,
                       const double* d_vals, const uint16_t* d_idx, PairEntry* d_table, uint8_t* d_codes,
                       int32_t* d_npairs, int32_t* overflow, int32_t* max_pairs) {
  *overflow = 0;
  *max_pairs = 0;
  if (nlisted <= 0) return CASK_B200_OK;
  if (slice_rows <= 0 || (slice_rows & 3)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "pair dictionaries: slice_rows must be a multiple of 4");
  int32_t* d_over = nullptr;
  CB_TRY(dev::alloc((void**)&d_over, 16));
  int rc = dev::zero(ex, d_over, 16);
  BuildPairs f{d_val_off, d_width, slice_rows, d_vals, d_idx, d_table, d_codes, d_npairs, d_over};
  if (rc == CASK_B200_OK) rc = dev::for_each(ex, nlisted, f);
  if (rc == CASK_B200_OK) rc = dev::download(ex, overflow, d_over, sizeof(int32_t));
  dev::release(d_over);
  CB_TRY(rc);
  std::vector<int32_t> nd((size_t)nlisted);
  CB_TRY(dev::download(ex, nd.data(), d_npairs, sizeof(int32_t) * (size_t)nlisted));
  for (int32_t v : nd) *max_pairs = v > *max_pairs ? v : *max_pairs;
  return CASK_B200_OK;
}

}  // namespace valuedict
}  // namespace caskb200
