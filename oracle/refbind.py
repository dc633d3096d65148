"""ctypes binding of oracle/_ref/libcaskref.so — the reference's own code, compiled in place.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, tests/golden/make_golden.py and bench.py's
reference / cpu_baseline legs; never by the product package `cask_b200`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcaskref.so")

SCALAR_NAMES = (
    "nBlocks", "n", "paddingCycles", "totalCycles", "vector_load_cycles", "outSize",
    "reductionCycles", "emptyCycles", "m_colptr_unpaddedLength",
    "m_indptr_values_unpaddedLength", "len_colptr", "len_pairs",
)

PAIR_DTYPE = np.dtype([("value", "<f8"), ("indptr", "<i4")], align=False)  # 12 bytes, packed


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_last_error.restype = C.c_char_p
        L.ref_read_matrix.restype = C.c_void_p
        L.ref_read_matrix.argtypes = [C.c_char_p, C.c_int]
        L.ref_from_csr.restype = C.c_void_p
        L.ref_from_csr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.ref_get_csr.argtypes = [C.c_void_p] * 4
        L.ref_read_vector.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.ref_dot.restype = C.c_double
        L.ref_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_preprocess.restype = C.c_double
        L.ref_preprocess.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.ref_num_partitions.argtypes = [C.c_void_p]
        L.ref_partition_scalars.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_partition_arrays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_estimated_clock_cycles.restype = C.c_double
        L.ref_estimated_clock_cycles.argtypes = [C.c_void_p]
        L.ref_spmv_mock.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_slice_rows.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3
        L.ref_slice_columns.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3
        # reference solvers (ref_solvers.cpp: pcg<> / ILUPreconditioner compiled with ref_shim/mkl.h)
        if hasattr(L, "ref_pcg"):
            L.ref_solvers_last_error.restype = C.c_char_p
            L.ref_pcg.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int), C.c_int]
            L.ref_ilu.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_int)]
            L.ref_ilu_apply.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5
        assert L.ref_sizeof_pair() == 12
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefMatrix:
    """A cask::CsrMatrix held by the compiled reference."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(lib().ref_last_error().decode())
        self.h = handle
        n, m, nnz = C.c_int(), C.c_int(), C.c_int()
        lib().ref_dims(self.h, n, m, nnz)
        self.n, self.m, self.nnz = n.value, m.value, nnz.value

    @classmethod
    def read(cls, path, sym_lower=False):
        return cls(lib().ref_read_matrix(path.encode(), 1 if sym_lower else 0))

    @classmethod
    def from_csr(cls, n, m, row_ptr, col_ind, values):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        col_ind = np.ascontiguousarray(col_ind, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        return cls(lib().ref_from_csr(n, m, len(values), _p(row_ptr), _p(col_ind), _p(values)))

    def __del__(self):
        try:
            lib().ref_free(self.h)
        except Exception:
            pass

    def csr(self):
        rp = np.empty(self.n + 1, np.int32)
        ci = np.empty(self.nnz, np.int32)
        va = np.empty(self.nnz, np.float64)
        lib().ref_get_csr(self.h, _p(rp), _p(ci), _p(va))
        return rp, ci, va

    def dot(self, x, return_seconds=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n, np.float64)
        s = lib().ref_dot(self.h, _p(x), len(x), _p(y), self.n)
        return (y, s) if return_seconds else y

    def preprocess(self, arch, num_pipes, cache_size, input_width, max_rows=1 << 30, num_controllers=1):
        """Runs the reference preprocess(); returns list of (scalars dict, colptr, pairs)."""
        s = lib().ref_preprocess(self.h, arch, num_pipes, cache_size, input_width, max_rows, num_controllers)
        if s < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        self.preprocess_seconds = s
        out = []
        for p in range(lib().ref_num_partitions(self.h)):
            sc = np.zeros(12, np.int64)
            lib().ref_partition_scalars(self.h, p, _p(sc))
            colptr = np.empty(int(sc[10]), np.int32)
            pairs = np.empty(int(sc[11]), PAIR_DTYPE)
            lib().ref_partition_arrays(self.h, p, _p(colptr), _p(pairs))
            out.append((dict(zip(SCALAR_NAMES, (int(v) for v in sc))), colptr, pairs))
        return out

    def spmv_mock(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n, np.float64)
        rc = lib().ref_spmv_mock(self.h, _p(x), len(x), _p(y), self.n)
        return rc, lib().ref_last_error().decode() if rc else "", y

    def slice_rows(self, start, nrows):
        rp = np.empty(nrows + 1, np.int32)
        ci = np.empty(self.nnz, np.int32)
        va = np.empty(self.nnz, np.float64)
        k = lib().ref_slice_rows(self.h, start, nrows, _p(rp), _p(ci), _p(va))
        return rp, ci[:k].copy(), va[:k].copy()

    def slice_columns(self, block_size, b):
        rp = np.empty(self.n, np.int32)
        ci = np.empty(self.nnz, np.int32)
        va = np.empty(self.nnz, np.float64)
        k = lib().ref_slice_columns(self.h, block_size, b, _p(rp), _p(ci), _p(va))
        if k < 0:
            return None
        return rp, ci[:k].copy(), va[:k].copy()


def read_vector(path):
    n = lib().ref_read_vector(path.encode(), None, 0)
    if n < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    out = np.zeros(n, np.float64)
    lib().ref_read_vector(path.encode(), _p(out), n)
    return out


def _csr3(row_ptr, col_ind, values):
    return (np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col_ind, np.int32),
            np.ascontiguousarray(values, np.float64))


def solvers_available():
    return available() and hasattr(lib(), "ref_pcg")


def pcg(n, row_ptr, col_ind, values, rhs, x0=None, precon=0, iterations=0):
    """The reference's own pcg<double, Precon> (SparseLinearSolvers.hpp:162-239) on the CSR as given
    (its tests pass the stored lower triangle). precon: 0 identity, 1 ILU. Returns (converged, iterations, x)."""
    rp, ci, va = _csr3(row_ptr, col_ind, values)
    rhs = np.ascontiguousarray(rhs, np.float64)
    x = np.zeros(n, np.float64) if x0 is None else np.array(x0, np.float64)
    it = C.c_int(iterations)
    rc = lib().ref_pcg(n, len(va), _p(rp), _p(ci), _p(va), _p(rhs), _p(x), C.byref(it), precon)
    if rc < 0:
        raise RuntimeError(lib().ref_solvers_last_error().decode())
    return bool(rc), it.value, x


def ilu(n, row_ptr, col_ind, values):
    """ILUPreconditioner{a}.pc in the pattern of a, and its nnzs field."""
    rp, ci, va = _csr3(row_ptr, col_ind, values)
    pc = np.zeros(len(va), np.float64)
    nn = C.c_int(0)
    if lib().ref_ilu(n, len(va), _p(rp), _p(ci), _p(va), _p(pc), C.byref(nn)) < 0:
        raise RuntimeError(lib().ref_solvers_last_error().decode())
    return pc, nn.value


def ilu_apply(n, row_ptr, col_ind, values, x):
    rp, ci, va = _csr3(row_ptr, col_ind, values)
    x = np.ascontiguousarray(x, np.float64)
    z = np.zeros(n, np.float64)
    if lib().ref_ilu_apply(n, len(va), _p(rp), _p(ci), _p(va), _p(x), _p(z)) < 0:
        raise RuntimeError(lib().ref_solvers_last_error().decode())
    return z
