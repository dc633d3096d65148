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
