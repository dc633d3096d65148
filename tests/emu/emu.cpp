// TEST INFRASTRUCTURE ONLY: the device-logic sources of libcask_b200.so compiled against the host emulation of
// their backend (dev_host.hpp) and exposed to pytest through a C interface.  See tests/emu/__init__.py.
#include "dev_host.hpp"
#include "../../cask_b200/csrc/ingest_logic.inl"

using namespace caskb200;

extern "C" {

const char* emu_error() { return emu_last_error().c_str(); }
int64_t emu_live_allocations() { return dev::live_allocations(); }

// host arrays in, host arrays out (row_ptr: n + 1 entries; col / val: capacity `cap` entries).
// info[0] = nnz, info[1] = nnzs_field, info[2] = error bits, info[3] = first bad entry, info[4] = launches
int emu_coo_to_csr(int64_t n, int64_t m, int64_t L, const int32_t* rows, const int32_t* cols, const double* vals, int flags,
                   int order, int32_t* row_ptr, int32_t* col, double* val, int64_t cap, int64_t* info) {
  dev::Exec ex;
  int64_t launches = 0;
  ex.launches = &launches;
  ex.order = order;
  ingest::CsrArrays out;
  int32_t err = 0;
  int64_t first_bad = -1;
  const int rc = ingest::coo_to_csr(ex, n, m, L, rows, cols, vals, flags, &out, &err, &first_bad);
  info[0] = out.nnz; info[1] = out.nnzs_field; info[2] = err; info[3] = first_bad; info[4] = launches;
  if (rc == CASK_B200_OK && !err) {
    std::memcpy(row_ptr, out.row_ptr, sizeof(int32_t) * (size_t)(n + 1));
    if (out.nnz <= cap) {
      std::memcpy(col, out.col, sizeof(int32_t) * (size_t)out.nnz);
      std::memcpy(val, out.val, sizeof(double) * (size_t)out.nnz);
    }
  }
  dev::release(out.row_ptr); dev::release(out.col); dev::release(out.val);
  return rc;
}

}  // extern "C"

// ---- ILU(0) ------------------------------------------------------------------------------------------------
#include "../../cask_b200/csrc/precond_logic.inl"

extern "C" {

// Factorises (pc_out: nnz values in the pattern) and, if x != NULL, applies z = M^-1 x.
// info[0] = lower levels, info[1] = upper levels, info[2] = zero-pivot flag, info[3] = launches
int emu_ilu(int64_t n, int64_t nnz, const int32_t* rp, const int32_t* ci, const double* va, int order, int unit_lower,
            double* pc_out, const double* x, double* z, int64_t* info) {
  dev::Exec ex;
  int64_t launches = 0;
  ex.launches = &launches;
  ex.order = order;
  precond::IluState st;
  int rc = precond::ilu_analyse(ex, n, nnz, rp, ci, &st);
  if (rc == CASK_B200_OK) rc = precond::ilu_factor(ex, va, &st);
  int32_t zp = 0;
  if (rc == CASK_B200_OK && pc_out) std::memcpy(pc_out, st.pc, sizeof(double) * (size_t)nnz);
  if (rc == CASK_B200_OK && x) rc = precond::ilu_apply(ex, &st, unit_lower, x, z, &zp);
  info[0] = (int64_t)st.ptr_l.size() - 1; info[1] = (int64_t)st.ptr_u.size() - 1; info[2] = zp; info[3] = launches;
  precond::ilu_free(&st);
  return rc;
}

int emu_inv_diag(int64_t n, const int32_t* rp, const int32_t* ci, const double* va, double* invd) {
  dev::Exec ex;
  precond::InvDiag f{rp, ci, va, invd};
  return dev::for_each(ex, n, f);
}

}  // extern "C"

// ---- value dictionaries of the coded staged-ELL format ---------------------------------------------------------
#include "../../cask_b200/csrc/valuedict_logic.inl"

extern "C" {

// table: nlisted * 256 doubles; codes: as long as vals; ndict: nlisted.  info[0] = overflow, info[1] = max entries,
// info[2] = launches
int emu_valuedict(int64_t nlisted, const int64_t* val_off, const int32_t* width, int32_t slice_rows, const double* vals,
                  int order, double* table, uint8_t* codes, int32_t* ndict, int64_t* info) {
  dev::Exec ex;
  int64_t launches = 0;
  ex.launches = &launches;
  ex.order = order;
  int32_t overflow = 0, max_entries = 0;
  const int rc = valuedict::build(ex, nlisted, val_off, width, slice_rows, vals, table, codes, ndict, &overflow, &max_entries);
  info[0] = overflow; info[1] = max_entries; info[2] = launches;
  return rc;
}

// pair dictionaries: table nlisted * 256 records of 16 bytes {value bits, displacement * 8, 0}, npairs nlisted.
// info[0] = overflow, info[1] = longest table (slot 0 included), info[2] = launches
int emu_pairdict(int64_t nlisted, const int64_t* val_off, const int32_t* width, int32_t slice_rows, const double* vals,
                 const uint16_t* idx, int order, void* table, uint8_t* codes, int32_t* npairs, int64_t* info) {
  dev::Exec ex;
  int64_t launches = 0;
  ex.launches = &launches;
  ex.order = order;
  int32_t overflow = 0, max_pairs = 0;
  const int rc = valuedict::build_pairs(ex, nlisted, val_off, width, slice_rows, vals, idx,
                                        static_cast<valuedict::PairEntry*>(table), codes, npairs, &overflow, &max_pairs);
  info[0] = overflow; info[1] = max_pairs; info[2] = launches;
  return rc;
}

}  // extern "C"
