"""The REAL Intel MKL of this image, for the CPU side of the comparison.

TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's cpu_baseline / `--impl reference` legs, never by the
product package `cask_b200`.

There is no MKL development package here, but PyTorch's libtorch_cpu.so links oneMKL statically and exports its
inspector-executor sparse BLAS and a few BLAS-1 routines.  That is the library the reference's CPU solver path is
written against (pcg<> -> mkl_dcsrsymv / cblas_*, src/runtime/SparseLinearSolvers.hpp:162-239; ILU apply ->
mkl_dcsrtrsv, MklLayer.hpp:61-84; lib/sparse-bench/src/cpu/CgMklExplicit.cpp).  Two things are offered:

  * CsrHandle: y = A x by mkl_sparse_d_mv on all host cores (general CSR, or symmetric with only the lower triangle
    read = the contract of mkl_dcsrsymv('l')), and triangular solves by mkl_sparse_d_trsv;
  * pcg(): the reference's OWN pcg<double, Precon> / ILUPreconditioner, compiled where they lie against
    oracle/ref_shim_mkl/mkl.h (oracle/_ref/libcaskref_mkl.so), i.e. the reference's loop with MKL's arithmetic.

Nothing is imported from torch: libtorch_cpu.so is opened with ctypes (RTLD_GLOBAL, so that libcaskref_mkl.so's
undefined MKL symbols bind to it).
"""
import ctypes as C
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_MKL_PATH = os.path.join(_HERE, "_ref", "libcaskref_mkl.so")

# mkl_spblas.h (oneMKL 2024)
INDEX_ZERO, INDEX_ONE = 0, 1
OP_N = 10
TYPE_GENERAL, TYPE_SYMMETRIC, TYPE_TRIANGULAR = 20, 21, 23
FILL_LOWER, FILL_UPPER, FILL_FULL = 40, 41, 42
DIAG_NON_UNIT, DIAG_UNIT = 50, 51


class _Descr(C.Structure):
    _fields_ = [("type", C.c_int), ("mode", C.c_int), ("diag", C.c_int)]


_mkl = None
_mkl_err = None


def _find_libtorch_cpu():
    spec = importlib.util.find_spec("torch")
    if spec is None or not spec.submodule_search_locations:
        return None
    hits = glob.glob(os.path.join(list(spec.submodule_search_locations)[0], "lib", "libtorch_cpu.so"))
    return hits[0] if hits else None


def mkl():
    """libtorch_cpu.so as an MKL provider; raises RuntimeError if it is absent or does not export the routines."""
    global _mkl, _mkl_err
    if _mkl is not None:
        return _mkl
    if _mkl_err is not None:
        raise RuntimeError(_mkl_err)
    path = _find_libtorch_cpu()
    try:
        if path is None:
            raise OSError("libtorch_cpu.so not found")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        L.mkl_sparse_d_create_csr.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        L.mkl_sparse_d_mv.argtypes = [C.c_int, C.c_double, vp, _Descr, vp, C.c_double, vp]
        L.mkl_sparse_d_trsv.argtypes = [C.c_int, C.c_double, vp, _Descr, vp, vp]
        L.mkl_sparse_destroy.argtypes = [vp]
        L.MKL_Get_Version_String.argtypes = [C.c_char_p, C.c_int]
        L.mkl_get_max_threads.restype = C.c_int
        L.cblas_daxpy, L.ddot_  # noqa: B018  (AttributeError if not exported)
    except (OSError, AttributeError) as e:
        _mkl_err = "Intel MKL is not reachable through libtorch_cpu.so: %s" % e
        raise RuntimeError(_mkl_err)
    _mkl = L
    return L


def available():
    try:
        mkl()
        return True
    except RuntimeError:
        return False


def version():
    buf = C.create_string_buffer(256)
    mkl().MKL_Get_Version_String(buf, 256)
    return buf.value.decode().strip()


def max_threads():
    return int(mkl().mkl_get_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CsrHandle:
    """A 0-based CSR matrix viewed by MKL (create_csr keeps pointers into the arrays, which this object owns)."""

    def __init__(self, n, m, row_ptr, col_ind, values):
        self.n, self.m = int(n), int(m)
        self.rp = np.ascontiguousarray(row_ptr, np.int32)
        self.ci = np.ascontiguousarray(col_ind, np.int32)
        self.va = np.ascontiguousarray(values, np.float64)
        self.h = C.c_void_p()
        rc = mkl().mkl_sparse_d_create_csr(C.byref(self.h), INDEX_ZERO, self.n, self.m, _p(self.rp),
                                           C.c_void_p(self.rp.ctypes.data + 4), _p(self.ci), _p(self.va))
        if rc != 0:
            raise RuntimeError("mkl_sparse_d_create_csr: status %d" % rc)

    def __del__(self):
        try:
            if self.h:
                mkl().mkl_sparse_destroy(self.h)
        except Exception:
            pass

    def _mv(self, descr, x, y):
        rc = mkl().mkl_sparse_d_mv(OP_N, 1.0, self.h, descr, _p(x), 0.0, _p(y))
        if rc != 0:
            raise RuntimeError("mkl_sparse_d_mv: status %d" % rc)

    def spmv(self, x, y=None):
        """y = A x, A as stored."""
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.n, np.float64) if y is None else y
        self._mv(_Descr(TYPE_GENERAL, FILL_FULL, DIAG_NON_UNIT), x, y)
        return y

    def symv_lower(self, x, y=None):
        """y = A x with A symmetric and only its stored lower triangle read: mkl_dcsrsymv('l', ...)."""
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.n, np.float64) if y is None else y
        self._mv(_Descr(TYPE_SYMMETRIC, FILL_LOWER, DIAG_NON_UNIT), x, y)
        return y

    def trsv(self, x, lower, unit):
        """Solve T y = x, T the lower / upper triangle of A: mkl_dcsrtrsv(uplo, 'N', diag, ...)."""
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.n, np.float64)
        d = _Descr(TYPE_TRIANGULAR, FILL_LOWER if lower else FILL_UPPER, DIAG_UNIT if unit else DIAG_NON_UNIT)
        rc = mkl().mkl_sparse_d_trsv(OP_N, 1.0, self.h, d, _p(x), _p(y))
        if rc != 0:
            raise RuntimeError("mkl_sparse_d_trsv: status %d" % rc)
        return y


# ---- the reference's own pcg<> / ILUPreconditioner on MKL -----------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(REF_MKL_PATH) and available()


def ref():
    global _ref
    if _ref is None:
        mkl()  # RTLD_GLOBAL first: libcaskref_mkl.so's MKL symbols are undefined until then
        L = C.CDLL(REF_MKL_PATH)
        L.ref_solvers_last_error.restype = C.c_char_p
        L.ref_pcg_timed.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double)]
        L.ref_ilu_apply.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5
        L.ref_symv_timed.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_double)]
        assert L.ref_solvers_real_mkl() == 1
        _ref = L
    return _ref


def _csr3(row_ptr, col_ind, values):
    return (np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col_ind, np.int32),
            np.ascontiguousarray(values, np.float64))


def pcg(n, row_ptr, col_ind, values, rhs, x0=None, precon=0, iterations=0):
    """The reference's pcg<double, Precon> on MKL, on the CSR as given (it reads the lower triangle only).
    precon: 0 identity, 1 ILU.  Returns (converged, iterations, x, seconds of the reference's "cg:solve" timer)."""
    rp, ci, va = _csr3(row_ptr, col_ind, values)
    rhs = np.ascontiguousarray(rhs, np.float64)
    x = np.zeros(n, np.float64) if x0 is None else np.array(x0, np.float64)
    it, sec = C.c_int(iterations), C.c_double(0.0)
    rc = ref().ref_pcg_timed(n, len(va), _p(rp), _p(ci), _p(va), _p(rhs), _p(x), C.byref(it), precon, C.byref(sec))
    if rc < 0:
        raise RuntimeError(ref().ref_solvers_last_error().decode())
    return bool(rc), it.value, x, sec.value


def symv(n, row_ptr, col_ind, values, x, reps=1):
    """y = A x the way the reference's pcg<> computes it: mkl_dcsrsymv('l') on one-based copies of the CSR handed over
    (A symmetric, stored lower triangle read), through the compiled adapter.  Returns (y, seconds of `reps` calls)."""
    rp, ci, va = _csr3(row_ptr, col_ind, values)
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(n, np.float64)
    sec = C.c_double(0.0)
    if ref().ref_symv_timed(n, len(va), _p(rp), _p(ci), _p(va), _p(x), _p(y), reps, C.byref(sec)) < 0:
        raise RuntimeError(ref().ref_solvers_last_error().decode())
    return y, sec.value


def ilu_apply(n, row_ptr, col_ind, values, x):
    """ILUPreconditioner{a}.apply(x) with both triangular solves done by MKL."""
    rp, ci, va = _csr3(row_ptr, col_ind, values)
    x = np.ascontiguousarray(x, np.float64)
    z = np.zeros(n, np.float64)
    if ref().ref_ilu_apply(n, len(va), _p(rp), _p(ci), _p(va), _p(x), _p(z)) < 0:
        raise RuntimeError(ref().ref_solvers_last_error().decode())
    return z
