"""cask_b200 — B200-native SpMV / CG hot path of CASK behind a C ABI.

The product is `libcask_b200.so` (hand-written CUDA for sm_100a, `cask_b200/csrc/`) and the C++ host
mirror of the reference interface (`cask_b200/host/`).  This Python module is only the thinnest
ctypes view of that C ABI (`include/cask_b200.h`) for tests and bench.py: it holds no arithmetic and
no fallback — if the library is missing or no GPU is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcask_b200.so")

OK, ERR_INVALID_ARGUMENT, ERR_RUNTIME, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_NCCL = range(7)
ARCH_SIMPLE, ARCH_SKIPEMPTY = 0, 1
SYNTH_POISSON2D, SYNTH_POISSON3D27, SYNTH_CONVDIFF3D7 = 0, 1, 2

PAIR_DTYPE = np.dtype([("value", "<f8"), ("indptr", "<i4")], align=False)  # indptr_value, Spmv.hpp:13-20

EXPORTED_SYMBOLS = (
    "cask_b200_device_count", "cask_b200_create", "cask_b200_destroy", "cask_b200_last_error",
    "cask_b200_set_stream", "cask_b200_use_own_stream", "cask_b200_synchronize", "cask_b200_set_option", "cask_b200_preprocess",
    "cask_b200_preprocess_device", "cask_b200_plan_get_stats", "cask_b200_plan_estimate", "cask_b200_partition_get_info",
    "cask_b200_partition_export", "cask_b200_spmv", "cask_b200_spmv_device", "cask_b200_spmv_refformat",
    "cask_b200_cg", "cask_b200_cg_device", "cask_b200_bicgstab", "cask_b200_bicgstab_device",
    "cask_b200_nccl_unique_id", "cask_b200_dist_init", "cask_b200_shard_rows",
    "cask_b200_preprocess_shard_device", "cask_b200_dist_halo_counts", "cask_b200_dist_peer_active", "cask_b200_dist_vector", "cask_b200_spmv_shard", "cask_b200_halo_plan_host", "cask_b200_sparse_segments_host", "cask_b200_sparse_send_plan_host", "cask_b200_synth_rows",
    "cask_b200_synth_nnz", "cask_b200_synth_device", "cask_b200_launch_count", "cask_b200_legacy_write",
    "cask_b200_legacy_read", "cask_b200_legacy_run", "cask_b200_legacy_reset", "cask_b200_legacy_launch_count",
    "cask_b200_mm_read_info", "cask_b200_mm_read_coo", "cask_b200_mm_read_vector", "cask_b200_ingest_coo",
    "cask_b200_ingest_coo_device", "cask_b200_read_matrix", "cask_b200_csr_get_info", "cask_b200_csr_export",
    "cask_b200_csr_device_arrays", "cask_b200_csr_free", "cask_b200_preprocess_csr",
    "cask_b200_plan_value_dict", "cask_b200_pcg", "cask_b200_pcg_device", "cask_b200_precond_set_matrix", "cask_b200_ilu_factor", "cask_b200_ilu_apply",
)
PRECON_IDENTITY, PRECON_ILU, PRECON_JACOBI, PRECON_ILU_UNIT = range(4)
INGEST_ONE_BASED, INGEST_SYMMETRIC, INGEST_DROP_UPPER = 1, 2, 4


class Design(C.Structure):
    """cask_b200_design == the parameters of GeneratedSpmvImplementation (GeneratedImplSupport.hpp:58)."""
    _fields_ = [(k, C.c_int32) for k in ("num_pipes", "cache_size", "input_width", "max_rows",
                                         "num_controllers", "dram_reduction_enabled", "arch")]


class PartitionInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("nBlocks", "n", "paddingCycles", "totalCycles", "vector_load_cycles",
                                         "outSize", "reductionCycles", "emptyCycles", "m_colptr_unpaddedLength",
                                         "m_indptr_values_unpaddedLength")] + [("len_colptr", C.c_int64),
                                                                               ("len_pairs", C.c_int64)]


class PlanStats(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("nnz", C.c_int64), ("slice_rows", C.c_int32),
                ("num_slices", C.c_int32), ("slices_staged_ell", C.c_int32), ("slices_gather_csr", C.c_int32),
                ("csr_lanes_per_row", C.c_int32), ("max_row_length", C.c_int32),
                ("ell_padded_entries", C.c_int64), ("ell_nnz", C.c_int64), ("xcache_doubles_total", C.c_int64),
                ("device_bytes", C.c_int64), ("row_length_histogram", C.c_int64 * 8),
                ("csr_nnz", C.c_int64), ("csr_rows", C.c_int64), ("csr_items", C.c_int32), ("csr_kernel", C.c_int32),
                ("persist_ku", C.c_int32), ("persist_stages", C.c_int32), ("persist_ctas_per_sm", C.c_int32),
                ("value_dict", C.c_int32), ("col_reorder", C.c_int32), ("cols_referenced", C.c_int64),
                ("merge_items", C.c_int32), ("merge_ctas", C.c_int32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "row_length_histogram"}
        d["row_length_histogram"] = list(self.row_length_histogram)
        return d


def plan_estimate(stats, hbm_gbs=0.0, l2_bytes=0.0, sm_clock_hz=0.0, sms=0):
    """(bytes, seconds) of one SpMV on a plan with these statistics (dict from Context.plan_stats() or a PlanStats):
    cask_b200_plan_estimate, the selector's model.  No GPU needed."""
    if not isinstance(stats, PlanStats):
        st = PlanStats()
        for k, _ in PlanStats._fields_:
            if k == "row_length_histogram":
                for i, v in enumerate(stats.get(k, [0] * 8)):
                    st.row_length_histogram[i] = int(v)
            elif k in stats:
                setattr(st, k, int(stats[k]))
        stats = st
    b, s = C.c_double(), C.c_double()
    check(lib().cask_b200_plan_estimate(C.byref(stats), C.c_double(hbm_gbs), C.c_double(l2_bytes), C.c_double(sm_clock_hz), int(sms),
                                        C.byref(b), C.byref(s)))
    return b.value, s.value


class MmInfo(C.Structure):
    """cask_b200_mm_info == struct MmInfo (IO.hpp:39-58) + the size line."""
    _fields_ = [("type", C.c_char * 16), ("format", C.c_char * 16), ("data_type", C.c_char * 16),
                ("symmetry", C.c_char * 16), ("n", C.c_int64), ("m", C.c_int64), ("entries", C.c_int64)]

    def as_dict(self):
        return {"type": self.type.decode(), "format": self.format.decode(), "data_type": self.data_type.decode(),
                "symmetry": self.symmetry.decode(), "n": self.n, "m": self.m, "entries": self.entries}


class CaskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cask_b200 error %d: %s" % (code, msg))
        self.code = code
        self.message = msg


_lib = None


def lib():
    """Loads libcask_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libcask_b200.so is not built: run `python -m cask_b200.build` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.cask_b200_last_error.restype = C.c_char_p
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.cask_b200_create.argtypes = [C.POINTER(vp), C.c_int]
        L.cask_b200_destroy.argtypes = [vp]
        L.cask_b200_set_stream.argtypes = [vp, vp]
        L.cask_b200_synchronize.argtypes = [vp]
        L.cask_b200_use_own_stream.argtypes = [vp]
        L.cask_b200_set_option.argtypes = [vp, C.c_char_p, dbl]
        L.cask_b200_preprocess.argtypes = [vp, C.POINTER(Design), i64, i64, i64, vp, vp, vp]
        L.cask_b200_preprocess_device.argtypes = [vp, C.POINTER(Design), i64, i64, i64, vp, vp, vp]
        L.cask_b200_preprocess_shard_device.argtypes = [vp, C.POINTER(Design), i64, i64, i64, i64, i64, vp, vp, vp]
        L.cask_b200_plan_get_stats.argtypes = [vp, C.POINTER(PlanStats)]
        L.cask_b200_plan_value_dict.argtypes = [vp, vp, vp, vp]
        L.cask_b200_plan_estimate.argtypes = [C.POINTER(PlanStats), dbl, dbl, dbl, C.c_int32, C.POINTER(dbl), C.POINTER(dbl)]
        L.cask_b200_partition_get_info.argtypes = [vp, i32, C.POINTER(PartitionInfo)]
        L.cask_b200_partition_export.argtypes = [vp, i32, vp, vp]
        L.cask_b200_spmv.argtypes = [vp, vp, vp]
        L.cask_b200_spmv_device.argtypes = [vp, vp, vp]
        L.cask_b200_spmv_refformat.argtypes = [vp, vp, vp]
        L.cask_b200_cg.argtypes = [vp, vp, vp, i32, dbl, vp, vp, vp]
        L.cask_b200_cg_device.argtypes = [vp, vp, vp, i32, dbl, vp, vp, vp, vp]
        L.cask_b200_bicgstab.argtypes = [vp, vp, vp, vp, vp]
        L.cask_b200_bicgstab_device.argtypes = [vp, vp, vp, vp, vp]
        L.cask_b200_nccl_unique_id.argtypes = [vp]
        L.cask_b200_dist_init.argtypes = [vp, i32, i32, vp]
        L.cask_b200_shard_rows.argtypes = [i64, i32, i32, vp, vp]
        L.cask_b200_dist_halo_counts.argtypes = [vp, vp]
        L.cask_b200_dist_peer_active.argtypes = [vp, vp]
        L.cask_b200_dist_vector.argtypes = [vp, i32, C.POINTER(vp)]
        L.cask_b200_spmv_shard.argtypes = [vp, vp, vp]
        L.cask_b200_halo_plan_host.argtypes = [i64, i32, i32, i64, vp, vp, i64, vp, vp, vp, vp]
        L.cask_b200_sparse_segments_host.argtypes = [vp, i32, vp, i64, vp]
        L.cask_b200_sparse_send_plan_host.argtypes = [vp, i32, i32, vp, vp]
        L.cask_b200_synth_rows.argtypes = [i32, i32, vp]
        L.cask_b200_synth_nnz.argtypes = [i32, i32, i64, i64, vp]
        L.cask_b200_synth_device.argtypes = [i32, i32, i64, i64, vp, vp, vp, vp]
        L.cask_b200_launch_count.argtypes = [vp, vp]
        L.cask_b200_device_count.argtypes = [vp]
        L.cask_b200_mm_read_info.argtypes = [C.c_char_p, C.POINTER(MmInfo)]
        L.cask_b200_mm_read_coo.argtypes = [C.c_char_p, i64, vp, vp, vp, vp]
        L.cask_b200_mm_read_vector.argtypes = [C.c_char_p, i64, vp, vp]
        L.cask_b200_ingest_coo.argtypes = [vp, i64, i64, i64, vp, vp, vp, i32, C.POINTER(vp)]
        L.cask_b200_ingest_coo_device.argtypes = [vp, i64, i64, i64, vp, vp, vp, i32, C.POINTER(vp)]
        L.cask_b200_read_matrix.argtypes = [vp, C.c_char_p, i32, C.POINTER(vp)]
        L.cask_b200_csr_get_info.argtypes = [vp, vp, vp, vp, vp]
        L.cask_b200_csr_export.argtypes = [vp, vp, vp, vp, vp]
        L.cask_b200_csr_device_arrays.argtypes = [vp, vp, vp, vp]
        L.cask_b200_csr_free.argtypes = [vp]
        L.cask_b200_preprocess_csr.argtypes = [vp, C.POINTER(Design), vp]
        L.cask_b200_pcg.argtypes = [vp, vp, vp, i32, dbl, i32, vp, vp, vp]
        L.cask_b200_pcg_device.argtypes = [vp, vp, vp, i32, dbl, i32, vp, vp, vp]
        L.cask_b200_precond_set_matrix.argtypes = [vp, vp]
        L.cask_b200_ilu_factor.argtypes = [vp, vp, vp, vp]
        L.cask_b200_ilu_apply.argtypes = [vp, i32, vp, vp, vp]
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise CaskError(rc, lib().cask_b200_last_error().decode())


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw device pointer


def design(num_pipes=1, cache_size=8192, input_width=16, max_rows=0, num_controllers=1,
           dram_reduction_enabled=0, arch=ARCH_SIMPLE):
    return Design(num_pipes, cache_size, input_width, max_rows, num_controllers, dram_reduction_enabled, arch)


def shard_rows(n, world, rank):
    r0, nr = C.c_int64(), C.c_int64()
    check(lib().cask_b200_shard_rows(n, world, rank, C.byref(r0), C.byref(nr)))
    return r0.value, nr.value


def halo_plan_host(n_global, world, rank, run_col0, run_len):
    """[(peer, col0, len)] this rank must receive, from the x windows it stages (host arithmetic only)."""
    c0 = np.ascontiguousarray(run_col0, np.int64)
    ln = np.ascontiguousarray(run_len, np.int64)
    cap = 2 * len(c0) * max(world, 1) + 8
    peer, col0, length = np.zeros(cap, np.int32), np.zeros(cap, np.int64), np.zeros(cap, np.int64)
    cnt = C.c_int64()
    check(lib().cask_b200_halo_plan_host(n_global, world, rank, len(c0), _p(c0), _p(ln), cap, _p(peer), _p(col0), _p(length),
                                         C.byref(cnt)))
    return [(int(peer[i]), int(col0[i]), int(length[i])) for i in range(cnt.value)]


def sparse_segments_host(bounds, need):
    """seg[world + 1]: where each owner's columns start inside the ascending list `need` of referenced columns."""
    b = np.ascontiguousarray(bounds, np.int64)
    nd = np.ascontiguousarray(need, np.int32)
    seg = np.zeros(len(b), np.int64)
    check(lib().cask_b200_sparse_segments_host(_p(b), len(b) - 1, _p(nd), len(nd), _p(seg)))
    return seg


def sparse_send_plan_host(all_seg, rank):
    """(send_off[world + 1], dst_off[world]) of `rank` from every rank's segment boundaries (world x (world + 1))."""
    a = np.ascontiguousarray(all_seg, np.int64)
    world = a.shape[0]
    send_off, dst_off = np.zeros(world + 1, np.int64), np.zeros(world, np.int64)
    check(lib().cask_b200_sparse_send_plan_host(_p(a), world, rank, _p(send_off), _p(dst_off)))
    return send_off, dst_off


def synth_rows(kind, N):
    n = C.c_int64()
    check(lib().cask_b200_synth_rows(kind, N, C.byref(n)))
    return n.value


def synth_nnz(kind, N, row0, nrows):
    z = C.c_int64()
    check(lib().cask_b200_synth_nnz(kind, N, row0, nrows, C.byref(z)))
    return z.value


def mm_read_info(path):
    """io::readHeader + the size line (host only, no GPU needed)."""
    info = MmInfo()
    check(lib().cask_b200_mm_read_info(os.fsencode(path), C.byref(info)))
    return info.as_dict()


def mm_read_coo(path):
    """(info, rows, cols, vals): the entries of a coordinate file in file order, 1-based (host only)."""
    info = mm_read_info(path)
    L = min(info["entries"], os.path.getsize(path) // 6 + 1)   # an entry is at least 6 bytes: a corrupt size line sizes nothing
    rows, cols, vals = np.zeros(L, np.int32), np.zeros(L, np.int32), np.zeros(L, np.float64)
    cnt = C.c_int64()
    check(lib().cask_b200_mm_read_coo(os.fsencode(path), L, _p(rows), _p(cols), _p(vals), C.byref(cnt)))
    return info, rows, cols, vals


def mm_read_vector(path):
    """io::readVector (host only)."""
    n = min(mm_read_info(path)["n"], os.path.getsize(path) + 1)
    out = np.zeros(n, np.float64)
    cnt = C.c_int64()
    check(lib().cask_b200_mm_read_vector(os.fsencode(path), n, _p(out), C.byref(cnt)))
    return out


class DeviceCsr:
    """cask_b200_csr: a CSR matrix built on (and resident on) the GPU by the ingest path."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        n, m, nnz, field = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().cask_b200_csr_get_info(self.h, C.byref(n), C.byref(m), C.byref(nnz), C.byref(field)))
        self.n, self.m, self.nnz, self.nnzs_field = n.value, m.value, nnz.value, field.value

    def export(self):
        rp, ci, va = np.zeros(self.n + 1, np.int32), np.zeros(self.nnz, np.int32), np.zeros(self.nnz, np.float64)
        check(lib().cask_b200_csr_export(self.ctx.h, self.h, _p(rp), _p(ci), _p(va)))
        return rp, ci, va

    def device_arrays(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().cask_b200_csr_device_arrays(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def free(self):
        if self.h:
            lib().cask_b200_csr_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One cask_b200_ctx: one GPU, one stream."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().cask_b200_create(C.byref(self.h), device))
        self.n = self.m = 0

    def close(self):
        if self.h:
            lib().cask_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        """Share a caller's cudaStream_t (0 = CUDA's default stream), e.g. torch.cuda.current_stream().cuda_stream."""
        check(lib().cask_b200_set_stream(self.h, C.c_void_p(stream_ptr)))

    def use_own_stream(self):
        check(lib().cask_b200_use_own_stream(self.h))

    def synchronize(self):
        check(lib().cask_b200_synchronize(self.h))

    def set_option(self, name, value):
        check(lib().cask_b200_set_option(self.h, name.encode(), float(value)))

    def preprocess(self, dsg, n, m, row_ptr, col_ind, values):
        rp = np.ascontiguousarray(row_ptr, np.int32)
        ci = np.ascontiguousarray(col_ind, np.int32)
        va = np.ascontiguousarray(values, np.float64)
        check(lib().cask_b200_preprocess(self.h, C.byref(dsg), n, m, len(va), _p(rp), _p(ci), _p(va)))
        self.n, self.m = n, m

    def preprocess_device(self, dsg, n, m, nnz, d_row_ptr, d_col_ind, d_values):
        check(lib().cask_b200_preprocess_device(self.h, C.byref(dsg), n, m, nnz, _p(d_row_ptr), _p(d_col_ind),
                                                _p(d_values)))
        self.n, self.m = n, m

    def preprocess_shard_device(self, dsg, n_global, m, row0, nrows, nnz, d_row_ptr, d_col_ind, d_values):
        check(lib().cask_b200_preprocess_shard_device(self.h, C.byref(dsg), n_global, m, row0, nrows, nnz,
                                                      _p(d_row_ptr), _p(d_col_ind), _p(d_values)))
        self.n, self.m = nrows, m

    def ingest_coo(self, n, m, rows, cols, vals, flags):
        """COO (host arrays, file order) -> DeviceCsr with the reference's DokMatrix semantics."""
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        h = C.c_void_p()
        check(lib().cask_b200_ingest_coo(self.h, n, m, len(vals), _p(rows), _p(cols), _p(vals), flags, C.byref(h)))
        return DeviceCsr(self, h)

    def ingest_coo_device(self, n, m, count, d_rows, d_cols, d_vals, flags):
        h = C.c_void_p()
        check(lib().cask_b200_ingest_coo_device(self.h, n, m, count, _p(d_rows), _p(d_cols), _p(d_vals), flags, C.byref(h)))
        return DeviceCsr(self, h)

    def read_matrix(self, path, sym_lower=False):
        """io::readMatrix (sym_lower=False) / io::readSymMatrix().matrix (True), built on the GPU."""
        h = C.c_void_p()
        check(lib().cask_b200_read_matrix(self.h, os.fsencode(path), 1 if sym_lower else 0, C.byref(h)))
        return DeviceCsr(self, h)

    def preprocess_csr(self, dsg, csr):
        check(lib().cask_b200_preprocess_csr(self.h, C.byref(dsg), csr.h))
        self.n, self.m = csr.n, csr.m
        self._csr_keepalive = csr

    def plan_stats(self):
        st = PlanStats()
        check(lib().cask_b200_plan_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def value_dict(self):
        """(active, table entries staged per slice, matrix bytes one SpMV streams) of the current plan."""
        a, e, b = C.c_int32(), C.c_int32(), C.c_int64()
        check(lib().cask_b200_plan_value_dict(self.h, C.byref(a), C.byref(e), C.byref(b)))
        return bool(a.value), e.value, b.value

    def value_dict_mode(self):
        """0 uncoded, 1 value codes (3 B per stored nonzero), 2 (value, displacement) pair codes (1 B)."""
        a = C.c_int32()
        check(lib().cask_b200_plan_value_dict(self.h, C.byref(a), None, None))
        return a.value

    def partition(self, pipe, arrays=True):
        """(info dict, colptr, pairs) of reference partition `pipe`, produced by the GPU partitioner."""
        info = PartitionInfo()
        check(lib().cask_b200_partition_get_info(self.h, pipe, C.byref(info)))
        d = {k: int(getattr(info, k)) for k, _ in PartitionInfo._fields_}
        if not arrays:
            return d, None, None
        colptr = np.empty(info.len_colptr, np.int32)
        pairs = np.empty(info.len_pairs, PAIR_DTYPE)
        check(lib().cask_b200_partition_export(self.h, pipe, _p(colptr), _p(pairs)))
        return d, colptr, pairs

    def _vec(self, v, want, what, copy=False):
        """float64 contiguous view (or copy) of a host vector whose length the C side takes on trust."""
        v = np.array(v, np.float64) if copy else np.ascontiguousarray(v, np.float64)
        if v.ndim != 1 or len(v) != want:
            raise ValueError('%s has %s entries, the matrix needs %d' % (what, v.shape, want))
        return v

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        if self.m and len(x) != self.m:
            raise ValueError('x has %d entries, matrix has %d columns' % (len(x), self.m))
        y = np.empty(self.n, np.float64)
        check(lib().cask_b200_spmv(self.h, _p(x), _p(y)))
        return y

    def spmv_into(self, x, y):
        if len(x) != self.m or len(y) != self.n or x.dtype != np.float64 or y.dtype != np.float64:
            raise ValueError('spmv_into: x needs %d and y %d float64 entries' % (self.m, self.n))
        check(lib().cask_b200_spmv(self.h, _p(x), _p(y)))

    def spmv_refformat(self, x):
        x = self._vec(x, self.m, 'x')
        y = np.empty(self.n, np.float64)
        check(lib().cask_b200_spmv_refformat(self.h, _p(x), _p(y)))
        return y

    def spmv_device(self, d_x, d_y):
        check(lib().cask_b200_spmv_device(self.h, _p(d_x), _p(d_y)))

    def cg(self, rhs, x0=None, maxiters=2000, tol=1e-5, iterations=0):
        """Returns (converged, iterations, x, rs_final) — iterations with the reference's convention."""
        rhs = self._vec(rhs, self.n, 'rhs')
        x = np.zeros(self.n, np.float64) if x0 is None else self._vec(x0, self.n, 'x0', copy=True)
        it, conv, rs = C.c_int32(iterations), C.c_int32(0), C.c_double(0)
        check(lib().cask_b200_cg(self.h, _p(rhs), _p(x), maxiters, tol, C.byref(it), C.byref(conv), C.byref(rs)))
        return bool(conv.value), it.value, x, rs.value

    def cg_device(self, d_rhs, d_x, maxiters=2000, tol=1e-5, iterations=0):
        it, conv, rs, trips = C.c_int32(iterations), C.c_int32(0), C.c_double(0), C.c_int32(0)
        check(lib().cask_b200_cg_device(self.h, _p(d_rhs), _p(d_x), maxiters, tol, C.byref(it), C.byref(conv),
                                        C.byref(rs), C.byref(trips)))
        return bool(conv.value), it.value, rs.value, trips.value

    def pcg(self, rhs, precon, x0=None, maxiters=2000, tol=1e-5, iterations=0):
        """pcg<double, Precon>: returns (converged, iterations, x, rs_final); x is returned even when an ILU solve met a
        zero pivot (CaskError is raised after the download in that case)."""
        rhs = self._vec(rhs, self.n, 'rhs')
        x = np.zeros(self.n, np.float64) if x0 is None else self._vec(x0, self.n, 'x0', copy=True)
        it, conv, rs = C.c_int32(iterations), C.c_int32(0), C.c_double(0)
        check(lib().cask_b200_pcg(self.h, _p(rhs), _p(x), maxiters, tol, precon, C.byref(it), C.byref(conv), C.byref(rs)))
        return bool(conv.value), it.value, x, rs.value

    def pcg_device(self, d_rhs, d_x, precon, maxiters=2000, tol=1e-5, iterations=0):
        it, conv, rs = C.c_int32(iterations), C.c_int32(0), C.c_double(0)
        check(lib().cask_b200_pcg_device(self.h, _p(d_rhs), _p(d_x), maxiters, tol, precon, C.byref(it), C.byref(conv),
                                         C.byref(rs)))
        return bool(conv.value), it.value, rs.value

    def precond_set_matrix(self, csr):
        check(lib().cask_b200_precond_set_matrix(self.h, csr.h if csr is not None else None))
        self._precond_keepalive = csr

    def ilu_factor(self, nnz):
        """(pc in the pattern of the matrix, lower levels, upper levels): ILUPreconditioner's factors from the GPU."""
        pc = np.zeros(max(nnz, 1), np.float64)
        ll, lu = C.c_int32(), C.c_int32()
        check(lib().cask_b200_ilu_factor(self.h, _p(pc), C.byref(ll), C.byref(lu)))
        return pc[:nnz], ll.value, lu.value

    def ilu_apply(self, x, unit_lower=False):
        x = self._vec(x, self.n, 'x')
        z = np.zeros(len(x), np.float64)
        zp = C.c_int32()
        check(lib().cask_b200_ilu_apply(self.h, 1 if unit_lower else 0, _p(x), _p(z), C.byref(zp)))
        return z, bool(zp.value)

    def bicgstab(self, b, tol=0.0, maxit=0):
        b = self._vec(b, self.n, 'b')
        x = np.zeros(self.n, np.float64)
        it, te = C.c_int32(maxit), C.c_double(tol)
        check(lib().cask_b200_bicgstab(self.h, _p(b), _p(x), C.byref(it), C.byref(te)))
        return x, it.value, te.value

    def bicgstab_device(self, d_b, d_x, tol=0.0, maxit=0):
        it, te = C.c_int32(maxit), C.c_double(tol)
        check(lib().cask_b200_bicgstab_device(self.h, _p(d_b), _p(d_x), C.byref(it), C.byref(te)))
        return it.value, te.value

    def dist_init(self, rank, world, unique_id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(unique_id_bytes) if unique_id_bytes else None
        check(lib().cask_b200_dist_init(self.h, rank, world, buf))

    def halo_counts(self, world):
        out = np.zeros(world, np.int64)
        check(lib().cask_b200_dist_halo_counts(self.h, _p(out)))
        return out

    def dist_vector(self, channel=0):
        """Collective.  Device pointer (int) of full-layout vector `channel` of the symmetric arena, or None when the
        peer-memory path is not available for the current plan (use an own buffer then)."""
        ptr = C.c_void_p()
        check(lib().cask_b200_dist_vector(self.h, channel, C.byref(ptr)))
        return ptr.value

    def spmv_shard(self, x_slice, y_slice=None):
        """Collective.  Host buffers: this rank's slice of x in, its rows of y out."""
        x_slice = np.ascontiguousarray(x_slice, np.float64)
        y_slice = np.empty(len(x_slice), np.float64) if y_slice is None else y_slice
        check(lib().cask_b200_spmv_shard(self.h, _p(x_slice), _p(y_slice)))
        return y_slice

    def peer_active(self):
        a = C.c_int32()
        check(lib().cask_b200_dist_peer_active(self.h, C.byref(a)))
        return bool(a.value)

    def launch_count(self):
        c = C.c_int64()
        check(lib().cask_b200_launch_count(self.h, C.byref(c)))
        return c.value


def nccl_unique_id():
    buf = (C.c_char * 128)()
    check(lib().cask_b200_nccl_unique_id(buf))
    return bytes(buf)


def synth_device(kind, N, row0, nrows, d_row_ptr, d_col_ind, d_values, stream_ptr=0):
    check(lib().cask_b200_synth_device(kind, N, row0, nrows, _p(d_row_ptr), _p(d_col_ind), _p(d_values),
                                       C.c_void_p(stream_ptr)))
