// ILU(0) preconditioner of the reference's pcg<> as level-scheduled data-parallel steps.
// Included by precond.cu (after devlogic.cuh) and by tests/emu/emu.cpp (after tests/emu/dev_host.hpp).
//
// What it replaces (reference @ 9e561d7):
//   ILUPreconditioner::ILUPreconditioner  src/runtime/SparseLinearSolvers.hpp:89-140  ILU(0), IKJ order, on a DokMatrix
//   ILUPreconditioner::apply              src/runtime/SparseLinearSolvers.hpp:142-150 L y = r, U z = y through
//   cask::mkl::unittrsolve                src/runtime/MklLayer.hpp:61-84              mkl_dcsrtrsv(.., diag = 'N')
// Row i of the factorisation reads the finished rows k < i of its own pattern, and so does row i of the lower solve;
// the upper solve reads rows j > i.  Rows are therefore grouped into LEVELS (level(i) = 1 + max level of the rows it
// reads): all rows of a level are independent and run as one for_each, one thread per row, levels in sequence.
// Inside a row the arithmetic is the reference's, operation for operation (ascending k, ascending j, separate multiply
// and subtract), so the factors and the solves are bit-identical to the oracle's restatement.
//
// The reference's apply() solves with a NON-unit lower factor (diag = 'N' on getLowerTriangular(), which carries U's
// diagonal): M = (D + L)(D + U).  That M is not symmetric and the reference's own PCG stalls on SPD stencils
// (tests/test_oracle_precond.py); unit_lower = 1 applies the same factors the textbook way, M = (I + L)(D + U).
//
// Levels are computed on the host from a copy of the pattern (one O(nnz) pass; the reference builds the whole
// factorisation on the host through hash maps).  Requires rows in ascending column order without repeated columns -
// what CsrMatrix(const DokMatrix&) and the ingest path produce.
#include <vector>

namespace caskb200 {
namespace precond {

struct IluState {
  int64_t n = 0, nnz = 0;
  const int32_t* row_ptr = nullptr;  // device, borrowed
  const int32_t* col = nullptr;      // device, borrowed
  double* pc = nullptr;              // device, owned: the factors in the pattern of the matrix
  int32_t* diag_pos = nullptr;       // device, owned: position of (i, i) in row i, -1 if absent
  int32_t* order_l = nullptr;        // device, owned: rows sorted by lower level
  int32_t* order_u = nullptr;        // device, owned: rows sorted by upper level
  double* y = nullptr;               // device, owned: intermediate of apply()
  int32_t* flag = nullptr;           // device, owned: bit 0 = zero or missing pivot met in a solve
  std::vector<int64_t> ptr_l, ptr_u; // host: level l covers order_x[ptr_x[l] .. ptr_x[l+1])
  bool factored = false;
};

// ---- host: levels ---------------------------------------------------------------------------------------
// lower: level(i) = 1 + max level(j), j < i in row i (0 without such entries), rows visited ascending;
// upper: the mirror image over j > i, rows visited descending.  Returns false if a row is not strictly ascending.
inline bool compute_levels(int64_t n, const int32_t* rp, const int32_t* ci, bool lower, std::vector<int32_t>* order,
                           std::vector<int64_t>* ptr, std::vector<int32_t>* diag_pos) {
  std::vector<int32_t> lev((size_t)n, 0);
  int32_t nlev = n ? 1 : 0;
  for (int64_t s = 0; s < n; s++) {
    const int64_t i = lower ? s : n - 1 - s;
    int32_t l = 0;
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) {
      const int32_t j = ci[k];
      if (k > rp[i] && ci[k - 1] >= j) return false;
      if (j < 0 || j >= n) return false;
      if (lower ? j < i : j > i) l = lev[(size_t)j] + 1 > l ? lev[(size_t)j] + 1 : l;
      if (diag_pos && j == i) (*diag_pos)[(size_t)i] = k;
    }
    lev[(size_t)i] = l;
    if (l + 1 > nlev) nlev = l + 1;
  }
  ptr->assign((size_t)nlev + 1, 0);
  for (int64_t i = 0; i < n; i++) (*ptr)[(size_t)lev[(size_t)i] + 1]++;
  for (int32_t l = 0; l < nlev; l++) (*ptr)[(size_t)l + 1] += (*ptr)[(size_t)l];
  order->assign((size_t)n, 0);
  std::vector<int64_t> cur(ptr->begin(), ptr->end() - (nlev ? 1 : 0));
  for (int64_t i = 0; i < n; i++) (*order)[(size_t)cur[(size_t)lev[(size_t)i]]++] = (int32_t)i;  // ascending rows inside a level
  return true;
}

// ---- device: one thread per row of a level ------------------------------------------------------------------
#ifdef __CUDACC__
#define CB_UNROLL _Pragma("unroll")
#else
#define CB_UNROLL
#endif
constexpr int kRowBatch = 8;  // entries of a row whose operands are in flight together in the triangular solves
struct FactorRow {  // SparseLinearSolvers.hpp:93-117 for one row i
  const int32_t* row_ptr;
  const int32_t* col;
  const int32_t* diag_pos;
  const int32_t* order;  // already offset to the level
  double* pc;
  CB_DEV void operator()(int64_t t) const {
    const int32_t i = order[t];
    const int32_t rb = row_ptr[i], re = row_ptr[i + 1];
    for (int32_t a = rb; a < re; a++) {                        // :97  for (auto& p : pc.dok[i])
      const int32_t k = col[a];
      if (k >= i) break;                                       // :99-100
      const int32_t dk = diag_pos[k];
      if (dk < 0 || pc[dk] == 0.0) continue;                   // :101-102  !isNnz(k, k)
      const double beta = pc[a] / pc[dk];                      // :103-104
      pc[a] = beta;
      int32_t q = dk + 1;                                      // row k, columns > k (sorted: they follow the diagonal)
      const int32_t qe = row_ptr[k + 1];
      for (int32_t b = a + 1; b < re; b++) {                   // :106-109  columns j >= k + 1 of row i
        const int32_t j = col[b];
        while (q < qe && col[q] < j) q++;
        if (q < qe && col[q] == j && pc[q] != 0.0)             // :111  isNnz(k, j)
          pc[b] = dev::sub_rn(pc[b], dev::mul_rn(pc[q], beta));  // :112
      }
    }
  }
};

struct LowerRow {  // y_i = (x_i - sum_{j<i} L_ij y_j) / d,  d = pc_ii (the reference) or 1 (unit_lower)
  const int32_t* row_ptr;
  const int32_t* col;
  const int32_t* diag_pos;
  const int32_t* order;
  const double* pc;
  const double* x;
  double* y;
  int32_t unit_lower;
  int32_t* flag;
  CB_DEV void operator()(int64_t t) const {
    const int32_t i = order[t];
    double acc = x[i];
    // A level kernel is a handful of rows deep and latency bound: the operands of kRowBatch entries are loaded before the
    // first of them is used (two dependent memory round trips per batch instead of per entry).  The subtractions
    // still happen one by one in ascending column order - the reference's order, bit for bit.  Rows are sorted
    // (ilu_analyse refuses others), so the entries left of the diagonal are a prefix.
    const int32_t rb = row_ptr[i], re = row_ptr[i + 1];
    bool more = true;
    for (int32_t k0 = rb; k0 < re && more; k0 += kRowBatch) {
      int32_t j[kRowBatch];
      double a[kRowBatch], v[kRowBatch];
      CB_UNROLL
      for (int u = 0; u < kRowBatch; u++) {
        const bool in = k0 + u < re;
        j[u] = in ? col[k0 + u] : i;
        a[u] = in ? pc[k0 + u] : 0.0;
      }
      CB_UNROLL
      for (int u = 0; u < kRowBatch; u++) v[u] = j[u] < i ? dev::ld_l2(y + j[u]) : 0.0;
      CB_UNROLL
      for (int u = 0; u < kRowBatch; u++)
        if (j[u] < i) acc = dev::sub_rn(acc, dev::mul_rn(a[u], v[u]));
      more = j[kRowBatch - 1] < i;
    }
    double d = 1.0;
    if (!unit_lower) {
      const int32_t dp = diag_pos[i];
      d = dp >= 0 ? pc[dp] : 0.0;
      if (d == 0.0) dev::atomic_or_i32(flag, 1);
    }
    y[i] = acc / d;
  }
};

struct UpperRow {  // z_i = (y_i - sum_{j>i} U_ij z_j) / U_ii
  const int32_t* row_ptr;
  const int32_t* col;
  const int32_t* diag_pos;
  const int32_t* order;
  const double* pc;
  const double* y;
  double* z;
  int32_t* flag;
  CB_DEV void operator()(int64_t t) const {
    const int32_t i = order[t];
    const int32_t dp = diag_pos[i];
    double acc = dev::ld_l2(y + i);
    // without a stored diagonal the row still has to skip its lower part; operands in batches as in LowerRow
    const int32_t re = row_ptr[i + 1];
    for (int32_t k0 = dp >= 0 ? dp + 1 : row_ptr[i]; k0 < re; k0 += kRowBatch) {
      int32_t j[kRowBatch];
      double a[kRowBatch], v[kRowBatch];
      CB_UNROLL
      for (int u = 0; u < kRowBatch; u++) {
        const bool in = k0 + u < re;
        j[u] = in ? col[k0 + u] : i;
        a[u] = in ? pc[k0 + u] : 0.0;
      }
      CB_UNROLL
      for (int u = 0; u < kRowBatch; u++) v[u] = j[u] > i ? dev::ld_l2(z + j[u]) : 0.0;
      CB_UNROLL
      for (int u = 0; u < kRowBatch; u++)
        if (j[u] > i) acc = dev::sub_rn(acc, dev::mul_rn(a[u], v[u]));
    }
    const double d = dp >= 0 ? pc[dp] : 0.0;
    if (d == 0.0) dev::atomic_or_i32(flag, 1);
    z[i] = acc / d;
  }
};

struct InvDiag {  // Jacobi: 1 / a_ii, 1 where the diagonal is absent or zero (Eigen's DiagonalPreconditioner rule)
  const int32_t* row_ptr;
  const int32_t* col;
  const double* val;
  double* invd;
  CB_DEV void operator()(int64_t i) const {
    double d = 0.0;
    for (int32_t k = row_ptr[i]; k < row_ptr[i + 1]; k++)
      if (col[k] == (int32_t)i) d = val[k];
    invd[i] = d != 0.0 ? 1.0 / d : 1.0;
  }
};

inline void ilu_free(IluState* st) {
  dev::release(st->pc);
  dev::release(st->diag_pos);
  dev::release(st->order_l);
  dev::release(st->order_u);
  dev::release(st->y);
  dev::release(st->flag);
  *st = IluState();
}

// Pattern analysis: levels of both solves and the diagonal positions.  d_row_ptr / d_col stay borrowed by the state.
inline int ilu_analyse(dev::Exec& ex, int64_t n, int64_t nnz, const int32_t* d_row_ptr, const int32_t* d_col, IluState* st) {
  ilu_free(st);
  if (n < 0 || nnz < 0 || n > INT32_MAX - 1 || nnz > INT32_MAX) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ilu: bad dimensions");
  st->n = n;
  st->nnz = nnz;
  st->row_ptr = d_row_ptr;
  st->col = d_col;
  std::vector<int32_t> rp((size_t)n + 1, 0), ci((size_t)nnz), diag((size_t)n, -1), order_l, order_u;
  CB_TRY(dev::download(ex, rp.data(), d_row_ptr, sizeof(int32_t) * ((size_t)n + 1)));
  CB_TRY(dev::download(ex, ci.data(), d_col, sizeof(int32_t) * (size_t)nnz));
  if (n && (rp[0] != 0 || rp[(size_t)n] != nnz)) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ilu: row_ptr does not describe nnz entries");
  for (int64_t i = 0; i < n; i++)
    if (rp[(size_t)i + 1] < rp[(size_t)i]) return fail(CASK_B200_ERR_INVALID_ARGUMENT, "ilu: row_ptr is not monotone");
  const bool ok = compute_levels(n, rp.data(), ci.data(), true, &order_l, &st->ptr_l, &diag) &&
                  compute_levels(n, rp.data(), ci.data(), false, &order_u, &st->ptr_u, nullptr);
  if (!ok)
    return fail(CASK_B200_ERR_UNSUPPORTED, "ilu: rows must hold strictly ascending, in-range column indices (the layout of "
                                           "CsrMatrix(DokMatrix) and of the ingest path)");
  const size_t nb = sizeof(int32_t) * (size_t)(n ? n : 1);
  CB_TRY(dev::alloc((void**)&st->diag_pos, nb));
  CB_TRY(dev::alloc((void**)&st->order_l, nb));
  CB_TRY(dev::alloc((void**)&st->order_u, nb));
  CB_TRY(dev::alloc((void**)&st->pc, sizeof(double) * (size_t)(nnz ? nnz : 1)));
  CB_TRY(dev::alloc((void**)&st->y, sizeof(double) * (size_t)(n ? n : 1)));
  CB_TRY(dev::alloc((void**)&st->flag, 16));
  CB_TRY(dev::zero(ex, st->flag, 16));
  CB_TRY(dev::upload(ex, st->diag_pos, diag.data(), sizeof(int32_t) * (size_t)n));
  CB_TRY(dev::upload(ex, st->order_l, order_l.data(), sizeof(int32_t) * (size_t)n));
  CB_TRY(dev::upload(ex, st->order_u, order_u.data(), sizeof(int32_t) * (size_t)n));
  return CASK_B200_OK;
}

// Numeric factorisation of the values d_val (same pattern): one for_each per lower level.
inline int ilu_factor(dev::Exec& ex, const double* d_val, IluState* st) {
  CB_TRY(dev::copy(ex, st->pc, d_val, sizeof(double) * (size_t)st->nnz));
  for (size_t l = 0; l + 1 < st->ptr_l.size(); l++) {
    if (l == 0) continue;  // level 0 rows read no other row: their factors are their values (row 0 is never touched, :93)
    FactorRow f{st->row_ptr, st->col, st->diag_pos, st->order_l + st->ptr_l[l], st->pc};
    CB_TRY(dev::for_each(ex, st->ptr_l[l + 1] - st->ptr_l[l], f));
  }
  st->factored = true;
  return CASK_B200_OK;
}

// z = M^-1 x.  *zero_pivot (optional) = 1 if a solve met a missing or zero diagonal (costs a synchronisation).
inline int ilu_apply(dev::Exec& ex, IluState* st, int unit_lower, const double* d_x, double* d_z, int32_t* zero_pivot) {
  if (!st->factored) return fail(CASK_B200_ERR_RUNTIME, "ilu: apply before factor");
  for (size_t l = 0; l + 1 < st->ptr_l.size(); l++) {
    LowerRow f{st->row_ptr, st->col, st->diag_pos, st->order_l + st->ptr_l[l], st->pc, d_x, st->y, unit_lower, st->flag};
    CB_TRY(dev::for_each(ex, st->ptr_l[l + 1] - st->ptr_l[l], f));
  }
  for (size_t l = 0; l + 1 < st->ptr_u.size(); l++) {
    UpperRow f{st->row_ptr, st->col, st->diag_pos, st->order_u + st->ptr_u[l], st->pc, st->y, d_z, st->flag};
    CB_TRY(dev::for_each(ex, st->ptr_u[l + 1] - st->ptr_u[l], f));
  }
  if (zero_pivot) CB_TRY(dev::download(ex, zero_pivot, st->flag, sizeof(int32_t)));
  return CASK_B200_OK;
}

}  // namespace precond
}  // namespace caskb200
