// COO -> CSR with the reference's dictionary-of-keys semantics, as data-parallel steps.
// Included by ingest.cu (after devlogic.cuh) and by tests/emu/emu.cpp (after tests/emu/dev_host.hpp): the text below
// only uses the `dev::` interface, so the same orchestration and functor bodies run on the GPU and in the emulation.
//
// What it replaces (reference @ 9e561d7):
//   io::readDokMatrix            src/runtime/IO.hpp:124-148        one DokMatrix::set(i-1, j-1, v) per file entry; a key that
//                                                                  appears twice keeps the LAST value (std::map assignment,
//                                                                  SparseMatrix.hpp:204-208) while nnzs counts every call
//   DokMatrix::explicitSymmetric src/runtime/SparseMatrix.hpp:156-189  every stored (i, j, v) also yields (j, i, v); if (j, i)
//                                                                  is stored too the values must be equal ("Matrix is not
//                                                                  symmetric" otherwise) and nnzs counts both visits twice
//   CsrMatrix(const DokMatrix&)  src/runtime/SparseMatrix.hpp:289-305  rows 0..n-1, columns ascending (std::map order),
//                                                                  row_ptr[n] = the nnzs FIELD (equal to the entry count
//                                                                  unless the file repeats a key)
// The hash-map build dominates the reference's end-to-end time; here it is one radix sort.
//
// Steps (E = L entries, or 2L with mirroring):
//   1. MakeKeys   key = row << 32 | col per entry, mirrors in the second half, dropped entries = all-ones sentinel
//   2. stable radix sort of (key, sequence number): a run of equal keys lists its direct entries in file order,
//      then its mirror entries in file order, so "last one wins" is "the run's last element"
//   3. MarkRuns   run ends, the reference's nnzs bookkeeping, and the symmetry check where a run has both kinds
//   4. inclusive sums -> output positions
//   5. Emit       columns, values and the row of every kept entry
//   6. RowPtr     row pointers from the (sorted) rows of the kept entries, empty rows included
namespace caskb200 {
namespace ingest {

constexpr int kOneBased = 1;    // CASK_B200_INGEST_ONE_BASED
constexpr int kSymmetric = 2;   // CASK_B200_INGEST_SYMMETRIC: explicitSymmetric
constexpr int kDropUpper = 4;   // CASK_B200_INGEST_DROP_UPPER: entries above the diagonal are ignored (mkl_dcsrsymv 'l')
constexpr int32_t kErrBadIndex = 1, kErrNotSymmetric = 2;
constexpr uint64_t kSentinel = ~0ull;

struct CsrArrays {       // device arrays owned by the caller of coo_to_csr once it has returned OK
  int64_t n = 0, m = 0;
  int64_t nnz = 0;         // entries stored
  int64_t nnzs_field = 0;  // what the reference's CsrMatrix::nnzs (and its row_ptr[n]) would hold
  int32_t* row_ptr = nullptr;
  int32_t* col = nullptr;
  double* val = nullptr;
};

struct MakeKeys {
  const int32_t* rows;
  const int32_t* cols;
  int64_t L, n, m;
  int32_t base, flags;
  uint64_t* keys;
  uint32_t* idx;
  int32_t* err;
  int64_t* first_bad;
  CB_DEV void operator()(int64_t e) const {
    const int64_t i = (int64_t)rows[e] - base, j = (int64_t)cols[e] - base;
    const bool sym = (flags & kSymmetric) != 0;
    bool ok = i >= 0 && i < n && j >= 0 && j < m;
    if (sym) ok = ok && j < n && i < m;  // the mirrored entry must exist in the matrix too
    idx[e] = (uint32_t)e;
    if (sym) idx[L + e] = (uint32_t)(L + e);
    if (!ok) {
      dev::atomic_or_i32(err, kErrBadIndex);
      dev::atomic_min_i64(first_bad, e);
      keys[e] = kSentinel;
      if (sym) keys[L + e] = kSentinel;
      return;
    }
    const bool keep = !((flags & kDropUpper) && j > i);
    keys[e] = keep ? ((uint64_t)i << 32) | (uint64_t)j : kSentinel;
    if (sym) keys[L + e] = keep && i != j ? ((uint64_t)j << 32) | (uint64_t)i : kSentinel;
  }
};

struct MarkRuns {
  const uint64_t* keys;   // sorted
  const uint32_t* idx;    // sequence numbers in sorted order: < L direct, >= L mirror
  const double* vals;     // file order
  int64_t E, L;
  int32_t sym;
  int32_t* flag;          // 1 at the last element of every run of a real key
  int32_t* cnt;           // the reference's nnzs contribution, at the last DIRECT element of every run
  int32_t* err;
  CB_DEV void operator()(int64_t p) const {
    const uint64_t key = keys[p];
    if (key == kSentinel) {
      flag[p] = 0;
      cnt[p] = 0;
      return;
    }
    const bool last = p + 1 == E || keys[p + 1] != key;
    const bool direct = (int64_t)idx[p] < L;
    const bool last_direct = direct && (last || (int64_t)idx[p + 1] >= L);
    const bool offdiag = (uint32_t)(key >> 32) != (uint32_t)key;
    cnt[p] = last_direct ? (sym && offdiag ? 2 : 1) : 0;
    flag[p] = last ? 1 : 0;
    if (last && !direct) {
      // the run ends in mirror entries; if it also holds direct ones the reference compares the stored (i, j) with the
      // stored (j, i) - i.e. the last direct value with the last mirror value (SparseMatrix.hpp:170-178)
      int64_t q = p - 1;
      while (q >= 0 && keys[q] == key && (int64_t)idx[q] >= L) q--;
      if (q >= 0 && keys[q] == key && vals[idx[q]] != vals[(int64_t)idx[p] - L]) dev::atomic_or_i32(err, kErrNotSymmetric);
    }
  }
};

struct Emit {
  const uint64_t* keys;
  const uint32_t* idx;
  const double* vals;
  const int32_t* flag;
  const int32_t* incl;   // inclusive sum of flag
  int64_t L;
  int32_t* col;
  double* val;
  int32_t* row_of;
  CB_DEV void operator()(int64_t p) const {
    if (!flag[p]) return;
    const int64_t o = (int64_t)incl[p] - 1;
    const uint64_t key = keys[p];
    const int64_t src = (int64_t)idx[p] < L ? (int64_t)idx[p] : (int64_t)idx[p] - L;
    col[o] = (int32_t)(uint32_t)key;
    row_of[o] = (int32_t)(key >> 32);
    val[o] = vals[src];
  }
};

struct RowPtr {
  const int32_t* row_of;  // ascending
  int64_t T, n;
  int32_t* row_ptr;       // n + 1 entries
  CB_DEV void operator()(int64_t o) const {
    const int64_t r = row_of[o];
    const int64_t prev = o ? (int64_t)row_of[o - 1] : -1;
    for (int64_t rr = prev + 1; rr <= r; rr++) row_ptr[rr] = (int32_t)o;   // first entry of row r; empty rows before it
    if (o == T - 1)
      for (int64_t rr = r + 1; rr <= n; rr++) row_ptr[rr] = (int32_t)T;    // empty rows after the last entry, and row_ptr[n]
  }
};

struct Scratch {  // freed on every exit path
  void* p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  ~Scratch() { for (void* q : p) dev::release(q); }
};

// d_rows / d_cols / d_vals: L entries in file order, resident on the device.  On CASK_B200_OK with *err_bits == 0, *out
// owns three fresh device arrays; with *err_bits != 0 nothing is returned (*first_bad = first entry with a bad index).
inline int coo_to_csr(dev::Exec& ex, int64_t n, int64_t m, int64_t L, const int32_t* d_rows, const int32_t* d_cols,
                      const double* d_vals, int flags, CsrArrays* out, int32_t* err_bits, int64_t* first_bad) {
  *out = CsrArrays();
  *err_bits = 0;
  *first_bad = -1;
  if (n < 0 || m < 0 || L < 0 || n > INT32_MAX - 1 || m > INT32_MAX - 1)
    return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ingest: bad dimensions");
  const bool sym = (flags & kSymmetric) != 0;
  const int64_t E = sym ? 2 * L : L;
  if (E > INT32_MAX) return fail(CASK_B200_ERR_UNSUPPORTED, "ingest: more than 2^31-1 entries after mirroring");
  out->n = n;
  out->m = m;
  int32_t* row_ptr = nullptr;
  CB_TRY(dev::alloc((void**)&row_ptr, sizeof(int32_t) * (size_t)(n + 1)));
  Scratch own_rp;  // released unless handed over
  own_rp.p[0] = row_ptr;
  CB_TRY(dev::zero(ex, row_ptr, sizeof(int32_t) * (size_t)(n + 1)));
  if (E == 0) {
    Scratch e;
    CB_TRY(dev::alloc(&e.p[0], 16));
    CB_TRY(dev::alloc(&e.p[1], 16));
    CB_TRY(dev::sync(ex));
    out->row_ptr = row_ptr;
    out->col = (int32_t*)e.p[0];
    out->val = (double*)e.p[1];
    own_rp.p[0] = e.p[0] = e.p[1] = nullptr;
    return CASK_B200_OK;
  }
  Scratch s;
  uint64_t *keys_a, *keys_b;
  uint32_t *idx_a, *idx_b;
  int32_t *flag, *cnt, *d_err;
  int64_t* d_first;
  CB_TRY(dev::alloc(&s.p[0], sizeof(uint64_t) * (size_t)E)); keys_a = (uint64_t*)s.p[0];
  CB_TRY(dev::alloc(&s.p[1], sizeof(uint64_t) * (size_t)E)); keys_b = (uint64_t*)s.p[1];
  CB_TRY(dev::alloc(&s.p[2], sizeof(uint32_t) * (size_t)E)); idx_a = (uint32_t*)s.p[2];
  CB_TRY(dev::alloc(&s.p[3], sizeof(uint32_t) * (size_t)E)); idx_b = (uint32_t*)s.p[3];
  CB_TRY(dev::alloc(&s.p[4], sizeof(int32_t) * (size_t)E)); flag = (int32_t*)s.p[4];
  CB_TRY(dev::alloc(&s.p[5], sizeof(int32_t) * (size_t)E)); cnt = (int32_t*)s.p[5];
  CB_TRY(dev::alloc(&s.p[6], 16)); d_err = (int32_t*)s.p[6];
  d_first = (int64_t*)((char*)s.p[6] + 8);
  const int64_t init_first = INT64_MAX;
  const int32_t init_err = 0;
  CB_TRY(dev::upload(ex, d_err, &init_err, sizeof(init_err)));
  CB_TRY(dev::upload(ex, d_first, &init_first, sizeof(init_first)));

  MakeKeys mk{d_rows, d_cols, L, n, m, (flags & kOneBased) ? 1 : 0, flags, keys_a, idx_a, d_err, d_first};
  CB_TRY(dev::for_each(ex, L, mk));
  CB_TRY(dev::sort_pairs_u64_u32(ex, keys_a, keys_b, idx_a, idx_b, E, 64));
  MarkRuns mr{keys_b, idx_b, d_vals, E, L, sym ? 1 : 0, flag, cnt, d_err};
  CB_TRY(dev::for_each(ex, E, mr));

  int32_t h_err = 0;
  int64_t h_first = 0;
  CB_TRY(dev::download(ex, &h_err, d_err, sizeof(h_err)));
  CB_TRY(dev::download(ex, &h_first, d_first, sizeof(h_first)));
  if (h_err) {
    *err_bits = h_err;
    *first_bad = (h_err & kErrBadIndex) ? h_first : -1;
    return CASK_B200_OK;
  }

  int32_t* incl = (int32_t*)keys_a;  // the unsorted keys are dead: reuse their storage for the two prefix sums
  int32_t* cnt_incl = incl + E;
  CB_TRY(dev::inclusive_sum_i32(ex, flag, incl, E));
  CB_TRY(dev::inclusive_sum_i32(ex, cnt, cnt_incl, E));
  int32_t T32 = 0, C32 = 0;
  CB_TRY(dev::download(ex, &T32, incl + (E - 1), sizeof(T32)));
  CB_TRY(dev::download(ex, &C32, cnt_incl + (E - 1), sizeof(C32)));
  const int64_t T = T32;
  out->nnz = T;
  // readDokMatrix counts one per set() call, duplicates included; explicitSymmetric recounts per stored entry
  out->nnzs_field = sym ? (int64_t)C32 : ((flags & kDropUpper) ? T : L);

  Scratch res;  // the result arrays: handed over only when everything has succeeded
  int32_t *row_of, *col;
  double* val;
  CB_TRY(dev::alloc(&res.p[0], sizeof(int32_t) * (size_t)(T ? T : 1))); col = (int32_t*)res.p[0];
  CB_TRY(dev::alloc(&res.p[1], sizeof(double) * (size_t)(T ? T : 1))); val = (double*)res.p[1];
  CB_TRY(dev::alloc(&s.p[7], sizeof(int32_t) * (size_t)(T ? T : 1))); row_of = (int32_t*)s.p[7];
  Emit em{keys_b, idx_b, d_vals, flag, incl, L, col, val, row_of};
  CB_TRY(dev::for_each(ex, E, em));
  RowPtr rp{row_of, T, n, row_ptr};
  CB_TRY(dev::for_each(ex, T, rp));
  CB_TRY(dev::sync(ex));
  out->row_ptr = row_ptr;
  out->col = col;
  out->val = val;
  own_rp.p[0] = nullptr;
  res.p[0] = res.p[1] = nullptr;
  return CASK_B200_OK;
}

}  // namespace ingest
}  // namespace caskb200
