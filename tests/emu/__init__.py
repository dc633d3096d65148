"""TEST INFRASTRUCTURE ONLY: builds and binds tests/emu/libcask_emu.so — the device-logic sources of the product
(cask_b200/csrc/*_logic.inl) compiled against a host emulation of their backend, so that index arithmetic can be
checked against the oracle without a GPU.  The product library never contains or calls any of this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libcask_emu.so")
_lib = None


def build():
    srcs = [os.path.join(HERE, f) for f in ("emu.cpp", "dev_host.hpp")]
    csrc = os.path.join(ROOT, "cask_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith("_logic.inl")]
    if os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra",
                           "-Wno-unused-parameter", "-ffp-contract=off", "-o", LIB, os.path.join(HERE, "emu.cpp")])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.emu_error.restype = C.c_char_p
        L.emu_live_allocations.restype = C.c_int64
        vp, i64 = C.c_void_p, C.c_int64
        L.emu_coo_to_csr.argtypes = [i64, i64, i64, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, i64, vp]
        L.emu_ilu.argtypes = [i64, i64, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]
        L.emu_inv_diag.argtypes = [i64, vp, vp, vp, vp]
        L.emu_valuedict.argtypes = [i64, vp, vp, C.c_int32, vp, C.c_int, vp, vp, vp, vp]
        L.emu_pairdict.argtypes = [i64, vp, vp, C.c_int32, vp, vp, C.c_int, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


INGEST_ONE_BASED, INGEST_SYMMETRIC, INGEST_DROP_UPPER = 1, 2, 4


def coo_to_csr(n, m, rows, cols, vals, flags, order=1):
    """Returns dict(rc, err, first_bad, nnz, nnzs_field, row_ptr, col, val, launches)."""
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    L = len(vals)
    cap = 2 * L + 1
    rp = np.zeros(n + 1, np.int32)
    ci = np.zeros(cap, np.int32)
    va = np.zeros(cap, np.float64)
    info = np.zeros(5, np.int64)
    rc = lib().emu_coo_to_csr(n, m, L, _p(rows), _p(cols), _p(vals), flags, order, _p(rp), _p(ci), _p(va), cap, _p(info))
    nnz = int(info[0])
    return {"rc": rc, "err": int(info[2]), "first_bad": int(info[3]), "nnz": nnz, "nnzs_field": int(info[1]),
            "row_ptr": rp, "col": ci[:nnz].copy(), "val": va[:nnz].copy(), "launches": int(info[4]),
            "message": lib().emu_error().decode()}


def ilu(n, row_ptr, col_ind, values, x=None, unit_lower=False, order=1):
    """Returns dict(rc, pc, z, levels_lower, levels_upper, zero_pivot, launches, message)."""
    rp = np.ascontiguousarray(row_ptr, np.int32)
    ci = np.ascontiguousarray(col_ind, np.int32)
    va = np.ascontiguousarray(values, np.float64)
    pc = np.zeros(max(len(va), 1), np.float64)
    info = np.zeros(4, np.int64)
    xx = None if x is None else np.ascontiguousarray(x, np.float64)
    z = np.zeros(max(n, 1), np.float64)
    rc = lib().emu_ilu(n, len(va), _p(rp), _p(ci), _p(va), order, 1 if unit_lower else 0, _p(pc),
                       None if xx is None else _p(xx), _p(z), _p(info))
    return {"rc": rc, "pc": pc[:len(va)], "z": z[:n], "levels_lower": int(info[0]), "levels_upper": int(info[1]),
            "zero_pivot": bool(info[2]), "launches": int(info[3]), "message": lib().emu_error().decode()}


def inv_diag(n, row_ptr, col_ind, values):
    rp = np.ascontiguousarray(row_ptr, np.int32)
    ci = np.ascontiguousarray(col_ind, np.int32)
    va = np.ascontiguousarray(values, np.float64)
    out = np.zeros(max(n, 1), np.float64)
    lib().emu_inv_diag(n, _p(rp), _p(ci), _p(va), _p(out))
    return out[:n]


def valuedict(val_off, width, vals, slice_rows=1024, order=1):
    """Per-slice value dictionaries of the coded staged-ELL format.  Returns dict(rc, overflow, max_entries, table
    [nlisted, 256], codes, ndict)."""
    vo = np.ascontiguousarray(val_off, np.int64)
    w = np.ascontiguousarray(width, np.int32)
    va = np.ascontiguousarray(vals, np.float64)
    table = np.full((max(len(vo), 1), 256), np.nan)
    codes = np.full(max(len(va), 1), 255, np.uint8)
    ndict = np.full(max(len(vo), 1), -1, np.int32)
    info = np.zeros(3, np.int64)
    rc = lib().emu_valuedict(len(vo), _p(vo), _p(w), slice_rows, _p(va), order, _p(table), _p(codes), _p(ndict), _p(info))
    return {"rc": rc, "overflow": bool(info[0]), "max_entries": int(info[1]), "table": table[:len(vo)],
            "codes": codes[:len(va)], "ndict": ndict[:len(vo)], "message": lib().emu_error().decode()}


PAIR_DTYPE = np.dtype([("value_bits", "<u8"), ("disp8", "<i4"), ("zero", "<i4")])


def pairdict(val_off, width, vals, idx, slice_rows=1024, order=1):
    """Per-slice (value, displacement) pair dictionaries of the pair-coded staged-ELL format.  Returns dict(rc, overflow,
    max_pairs, table [nlisted, 256] of PAIR_DTYPE records, codes, npairs)."""
    vo = np.ascontiguousarray(val_off, np.int64)
    w = np.ascontiguousarray(width, np.int32)
    va = np.ascontiguousarray(vals, np.float64)
    ix = np.ascontiguousarray(idx, np.uint16)
    table = np.zeros((max(len(vo), 1), 256), PAIR_DTYPE)
    table["zero"] = -1
    codes = np.full(max(len(va), 1), 255, np.uint8)
    npairs = np.full(max(len(vo), 1), -1, np.int32)
    info = np.zeros(3, np.int64)
    rc = lib().emu_pairdict(len(vo), _p(vo), _p(w), slice_rows, _p(va), _p(ix), order, _p(table), _p(codes), _p(npairs),
                            _p(info))
    return {"rc": rc, "overflow": bool(info[0]), "max_pairs": int(info[1]), "table": table[:len(vo)],
            "codes": codes[:len(va)], "npairs": npairs[:len(vo)], "message": lib().emu_error().decode()}
