"""ctypes binding of oracle/libcask_oracle.so (the plain-C CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package `cask_b200` never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcask_oracle.so")

PAIR_DTYPE = np.dtype([("value", "<f8"), ("indptr", "<i4")], align=False)
SCALAR_NAMES = (
    "nBlocks", "n", "paddingCycles", "totalCycles", "vector_load_cycles", "outSize",
    "reductionCycles", "emptyCycles", "m_colptr_unpaddedLength",
    "m_indptr_values_unpaddedLength", "len_colptr", "len_pairs",
)
ARCH_SIMPLE, ARCH_SKIPEMPTY = 0, 1


class _Partition(C.Structure):
    _fields_ = [(k, C.c_int32) for k in SCALAR_NAMES[:10]] + [
        ("len_colptr", C.c_int64), ("len_pairs", C.c_int64),
        ("m_colptr", C.POINTER(C.c_int32)), ("m_indptr_values", C.c_void_p),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.oracle_count_compute_cycles.argtypes = [vp, i32, i32]
        L.oracle_encode_empty_rows.argtypes = [vp, i32, vp]
        L.oracle_preprocess.argtypes = [i32, i32, vp, vp, vp, C.c_int, i32, i32, i32, vp]
        L.oracle_free_partition.argtypes = [vp]
        L.oracle_csr_dot.argtypes = [i32, vp, vp, vp, vp, vp]
        L.oracle_partition_spmv_w.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32]
        L.oracle_pcg.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, dbl]
        L.oracle_pcg_full.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, dbl, vp]
        L.oracle_bicgstab.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp]
        L.oracle_ilu0.argtypes = [i32, vp, vp, vp, vp]
        L.oracle_ilu_apply.argtypes = [i32, vp, vp, vp, vp, vp]
        L.oracle_ilu_apply_mode.argtypes = [i32, vp, vp, vp, vp, vp, C.c_int]
        L.oracle_pcg_precond.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, dbl, C.c_int, C.c_int, vp]
        for g in ("poisson2d", "poisson3d27", "convdiff3d7"):
            f = getattr(L, "oracle_gen_" + g)
            f.restype = i64
            f.argtypes = [i32, vp, vp, vp]
        L.oracle_gen_rmat.restype = i64
        L.oracle_gen_rmat.argtypes = [i32, i32, C.c_uint64, vp, vp, vp]
        L.oracle_csr_spmv_omp.argtypes = [i64, vp, vp, vp, vp, vp]
        L.oracle_csr_spmv_omp32.argtypes = [i32, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _csr(row_ptr, col_ind, values):
    return (np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col_ind, np.int32),
            np.ascontiguousarray(values, np.float64))


def count_compute_cycles(end_offsets, width):
    v = np.ascontiguousarray(end_offsets, np.int32)
    return lib().oracle_count_compute_cycles(_p(v), len(v), width)


def encode_empty_rows(end_offsets):
    v = np.ascontiguousarray(end_offsets, np.int32)
    out = np.zeros(max(len(v), 1), np.int32)
    k = lib().oracle_encode_empty_rows(_p(v), len(v), _p(out))
    return out[:k].copy()


def preprocess(n, m, row_ptr, col_ind, values, arch, num_pipes, cache_size, input_width):
    """Returns list of (scalars dict, colptr int32[], pairs PAIR_DTYPE[]) like refbind.preprocess."""
    rp, ci, va = _csr(row_ptr, col_ind, values)
    parts = (_Partition * num_pipes)()
    rc = lib().oracle_preprocess(n, m, _p(rp), _p(ci), _p(va), arch, num_pipes, cache_size,
                                 input_width, C.byref(parts))
    if rc:
        raise RuntimeError("oracle_preprocess failed rc=%d" % rc)
    out = []
    for p in parts:
        sc = {k: int(getattr(p, k)) for k in SCALAR_NAMES}
        colptr = np.ctypeslib.as_array(p.m_colptr, shape=(max(p.len_colptr, 1),))[:p.len_colptr].copy()
        nb = 12 * p.len_pairs
        buf = (C.c_char * max(nb, 1)).from_address(p.m_indptr_values)
        pairs = np.frombuffer(bytes(buf[:nb]), dtype=PAIR_DTYPE).copy()
        out.append((sc, colptr, pairs))
        lib().oracle_free_partition(C.byref(p))
    return out


def partition_spmv(parts, cache_size, input_width, x, n_total):
    """y from the partition arrays, as the device computes it."""
    x = np.ascontiguousarray(x, np.float64)
    arr = (_Partition * len(parts))()
    keep = []
    for q, (sc, colptr, pairs) in zip(arr, parts):
        for k in SCALAR_NAMES:
            setattr(q, k, sc[k])
        colptr = np.ascontiguousarray(colptr, np.int32)
        pairs = np.ascontiguousarray(pairs)
        keep += [colptr, pairs]
        q.m_colptr = colptr.ctypes.data_as(C.POINTER(C.c_int32))
        q.m_indptr_values = pairs.ctypes.data
    y = np.zeros(n_total, np.float64)
    rc = lib().oracle_partition_spmv_w(C.byref(arr), len(parts), cache_size, input_width, _p(x), len(x),
                                       _p(y), n_total)
    if rc:
        raise RuntimeError("oracle_partition_spmv failed rc=%d" % rc)
    return y


def csr_dot(n, row_ptr, col_ind, values, x):
    rp, ci, va = _csr(row_ptr, col_ind, values)
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(n, np.float64)
    lib().oracle_csr_dot(n, _p(rp), _p(ci), _p(va), _p(x), _p(y))
    return y


def pcg(n, row_ptr, col_ind, values, rhs, x0=None, maxiters=2000, tol=1e-5, lower=True):
    """Returns (converged, iterations, x[, rs_final]) with the reference's iteration quirk."""
    rp, ci, va = _csr(row_ptr, col_ind, values)
    rhs = np.ascontiguousarray(rhs, np.float64)
    x = np.zeros(n, np.float64) if x0 is None else np.array(x0, np.float64)
    it = C.c_int32(0)
    if lower:
        ok = lib().oracle_pcg(n, _p(rp), _p(ci), _p(va), _p(rhs), _p(x), C.byref(it), maxiters, tol)
        return bool(ok), it.value, x
    rs = C.c_double(0)
    ok = lib().oracle_pcg_full(n, _p(rp), _p(ci), _p(va), _p(rhs), _p(x), C.byref(it), maxiters, tol,
                               C.byref(rs))
    return bool(ok), it.value, x, rs.value


def ilu0(n, row_ptr, col_ind, values):
    """ILUPreconditioner's pc in the pattern of the given CSR (SparseLinearSolvers.hpp:89-140)."""
    rp, ci, va = _csr(row_ptr, col_ind, values)
    pc = np.zeros(len(va), np.float64)
    lib().oracle_ilu0(n, _p(rp), _p(ci), _p(va), _p(pc))
    return pc


def ilu_apply(n, row_ptr, col_ind, pc, x, unit_lower=False):
    """ILUPreconditioner::apply (:142-150). Returns (z, zero_pivot)."""
    rp, ci, pc = _csr(row_ptr, col_ind, pc)
    x = np.ascontiguousarray(x, np.float64)
    z = np.zeros(n, np.float64)
    rc = lib().oracle_ilu_apply_mode(n, _p(rp), _p(ci), _p(pc), _p(x), _p(z), 1 if unit_lower else 0)
    return z, bool(rc)


PRECON = {"identity": 0, "ilu": 1, "jacobi": 2, "ilu_unit": 3}


def pcg_precond(n, row_ptr, col_ind, values, rhs, precon, x0=None, maxiters=2000, tol=1e-5, lower=True,
                iterations=0):
    """pcg<double, Precon> with the preconditioner left in. Returns (converged, iterations, x, rs_final);
    converged is None if the ILU solve met a zero pivot."""
    rp, ci, va = _csr(row_ptr, col_ind, values)
    rhs = np.ascontiguousarray(rhs, np.float64)
    x = np.zeros(n, np.float64) if x0 is None else np.array(x0, np.float64)
    it = C.c_int32(iterations)
    rs = C.c_double(0)
    ok = lib().oracle_pcg_precond(n, _p(rp), _p(ci), _p(va), _p(rhs), _p(x), C.byref(it), maxiters, tol,
                                  PRECON[precon] if isinstance(precon, str) else precon, 1 if lower else 0,
                                  C.byref(rs))
    return (None if ok < 0 else bool(ok)), it.value, x, rs.value


def bicgstab(n, row_ptr, col_ind, values, b, tol=np.finfo(np.float64).eps, maxit=None):
    rp, ci, va = _csr(row_ptr, col_ind, values)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(n, np.float64)
    it = C.c_int32(2 * n if maxit is None else maxit)
    te = C.c_double(tol)
    lib().oracle_bicgstab(n, _p(rp), _p(ci), _p(va), _p(b), _p(x), C.byref(it), C.byref(te))
    return x, it.value, te.value


def _gen(name, n_rows, *args):
    f = getattr(lib(), "oracle_gen_" + name)
    nnz = f(*args, None, None, None)
    rp = np.zeros(n_rows + 1, np.int32)
    ci = np.zeros(nnz, np.int32)
    va = np.zeros(nnz, np.float64)
    f(*args, _p(rp), _p(ci), _p(va))
    return n_rows, rp, ci, va


def gen_poisson2d(N):
    return _gen("poisson2d", N * N, N)


def gen_poisson3d27(N):
    return _gen("poisson3d27", N ** 3, N)


def gen_convdiff3d7(N):
    return _gen("convdiff3d7", N ** 3, N)


def gen_rmat(scale, edge_factor=15, seed=1):
    return _gen("rmat", 1 << scale, scale, edge_factor, seed)


def csr_spmv_omp(n, row_ptr, col_ind, values, x, y=None):
    """OpenMP CSR row loop (CPU baseline port). row_ptr may be int32 or int64. Returns (y, threads)."""
    if y is None:
        y = np.zeros(n, np.float64)
    if row_ptr.dtype == np.int64:
        t = lib().oracle_csr_spmv_omp(n, _p(row_ptr), _p(col_ind), _p(values), _p(x), _p(y))
    else:
        t = lib().oracle_csr_spmv_omp32(n, _p(row_ptr), _p(col_ind), _p(values), _p(x), _p(y))
    return y, t
