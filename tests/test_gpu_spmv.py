"""y = A x on the GPU vs the reference's CsrMatrix::dot (golden, x[i] = 0.25 i as test_spmv.cpp:27-28)."""
import numpy as np
import pytest

from conftest import assert_y_close, row_scale

pytestmark = pytest.mark.gpu


def _check_all(golden, gpu_lib, ctx, dsg, exact, refformat=False):
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        x = golden.x(name)
        ctx.preprocess(dsg, n, m, rp, ci, va)
        got = ctx.spmv_refformat(x) if refformat else ctx.spmv(x)
        exp = golden.dots[name]
        if exact:
            assert np.array_equal(got, exp), name
        else:
            assert_y_close(got, exp, row_scale(n, rp, ci, va, x))
    return True


def test_staged_ell_is_bit_identical_to_reference_dot(golden, gpu_lib, ctx):
    """Thread-per-row, ascending column order, separate multiply and add == DokMatrix::dot's order."""
    ctx.set_option("force_kind", 0)
    _check_all(golden, gpu_lib, ctx, gpu_lib.design(1, 8192, 16), exact=False)
    st = None
    for name in ("test_dense_64", "test_cage6", "bfwb62", "test_small", "OPF_3754", "dw8192", "t2d_q9_A_01"):
        n, m, rp, ci, va = golden.csr(name)
        ctx.preprocess(gpu_lib.design(2, 24576, 16), n, m, rp, ci, va)
        st = ctx.plan_stats()
        assert st["slices_staged_ell"] > 0, (name, st)
        got, exp = ctx.spmv(golden.x(name)), golden.dots[name]
        if st["slices_gather_csr"] == 0:
            assert np.array_equal(got, exp), name
        else:  # rows of the staged slices are still bit-identical
            assert (got == exp).mean() > 0.5, (name, st)


@pytest.mark.parametrize("stream", [0, 1])
@pytest.mark.parametrize("vec", [2, 4, 8, 16, 32])
def test_gather_csr_vector_per_row(golden, gpu_lib, ctx, vec, stream):
    """stream=1: the CSR-stream variant (products staged in shared memory, then reduced row by row)."""
    ctx.set_option("force_kind", 1)
    ctx.set_option("force_csr_vec", vec)
    ctx.set_option("csr_stream", stream)
    if stream:
        ctx.set_option("csr_item_nnz", 1024 if vec == 2 else 4096)
    _check_all(golden, gpu_lib, ctx, gpu_lib.design(3, 2048, 16), exact=False)
    assert ctx.plan_stats()["slices_staged_ell"] == 0


def test_auto_selection_all_fixtures(golden, gpu_lib, ctx):
    for dsg in (gpu_lib.design(1, 2048, 16), gpu_lib.design(4, 512, 8, arch=1), gpu_lib.design(48, 64, 2)):
        _check_all(golden, gpu_lib, ctx, dsg, exact=False)


def test_small_cache_forces_gather(golden, gpu_lib, ctx):
    """A cache too small for the slice's x windows must fall back to the gather kernel, not fail."""
    n, m, rp, ci, va = golden.csr("t2d_q9_A_01")
    ctx.preprocess(gpu_lib.design(1, 64, 16), n, m, rp, ci, va)
    st = ctx.plan_stats()
    assert st["slices_gather_csr"] > 0
    x = golden.x("t2d_q9_A_01")
    assert_y_close(ctx.spmv(x), golden.dots["t2d_q9_A_01"], row_scale(n, rp, ci, va, x))


@pytest.mark.parametrize("arch", [0, 1])
def test_reference_format_describes_the_matrix(golden, gpu_lib, ctx, arch):
    """y computed from the emitted m_colptr / m_indptr_values the way the dataflow engine walks them."""
    for dsg in (gpu_lib.design(2, 8, 4, arch=arch), gpu_lib.design(1, 2048, 16, arch=arch), gpu_lib.design(5, 64, 3, arch=arch)):
        for name in golden.names:
            n, m, rp, ci, va = golden.csr(name)
            if n * ((m + dsg.cache_size - 1) // dsg.cache_size) > 20_000_000:
                continue
            x = golden.x(name)
            ctx.preprocess(dsg, n, m, rp, ci, va)
            assert_y_close(ctx.spmv_refformat(x), golden.dots[name], row_scale(n, rp, ci, va, x))


def test_edge_cases(gpu_lib, ctx, oracle):
    d = gpu_lib.design(2, 64, 4)
    # empty matrix rows, single row, single column, odd column count (bulk-copy tail), rectangular
    cases = []
    cases.append((5, 7, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0)))
    cases.append((1, 9, np.array([0, 9], np.int32), np.arange(9, dtype=np.int32), np.arange(1.0, 10.0)))
    cases.append((6, 1, np.arange(7, dtype=np.int32), np.zeros(6, np.int32), np.arange(1.0, 7.0)))
    rng = np.random.default_rng(3)
    for (n, m) in ((33, 17), (17, 1025), (1500, 1501), (2049, 33)):
        dense = (rng.random((n, m)) < 0.2) * rng.standard_normal((n, m))
        rp = np.concatenate([[0], np.cumsum((dense != 0).sum(1))]).astype(np.int32)
        ci = np.nonzero(dense)[1].astype(np.int32)
        cases.append((n, m, rp, ci, dense[dense != 0]))
    for kind in (-1, 0, 1):
        ctx.set_option("force_kind", kind)
        for n, m, rp, ci, va in cases:
            x = rng.standard_normal(m)
            ctx.preprocess(d, n, m, rp, ci, va)
            exp = oracle.csr_dot(n, rp, ci, va, x)
            assert_y_close(ctx.spmv(x), exp, row_scale(n, rp, ci, va, x))


def test_argument_checks_mirror_reference(golden, gpu_lib, ctx):
    """Messages and classes of Spmv::spmv's checks, src/runtime/Spmv.cpp:189-232."""
    n, m, rp, ci, va = golden.csr("test_small")
    x = golden.x("test_small")
    ctx.preprocess(gpu_lib.design(1, 64, 4, max_rows=8), n, m, rp, ci, va)
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.spmv(x)
    assert e.value.code == gpu_lib.ERR_INVALID_ARGUMENT
    assert e.value.message == "Matrix is too large! Maximum supported rows: 8 actual rows: 16"
    ctx.preprocess(gpu_lib.design(3, 64, 4, num_controllers=2), n, m, rp, ci, va)
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.spmv(x)
    assert e.value.code == gpu_lib.ERR_RUNTIME and e.value.message == "numPipes should be a multiple of numControllers"
    ctx.preprocess(gpu_lib.design(1, 64, 4, dram_reduction_enabled=1), n, m, rp, ci, va)
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.spmv(x)
    assert e.value.message.startswith("Matrix is too small! Minimum supported rows with DRAM reduction: 35000")
    with pytest.raises(gpu_lib.CaskError):
        gpu_lib.Context(0).spmv(x)


# ---- merge-path tiles (csr_kernel = 1; automatic from 2^20 gather nonzeros) -----------------------------------------
def test_merge_path_all_fixtures(golden, gpu_lib, ctx):
    """Every fixture through spmv_csr_merge_kernel + fix-up: forced gather everywhere, and mixed with staged slices."""
    ctx.set_option("csr_kernel", 1)
    ctx.set_option("force_kind", 1)
    _check_all(golden, gpu_lib, ctx, gpu_lib.design(3, 2048, 16), exact=False)
    st = ctx.plan_stats()
    assert st["slices_staged_ell"] == 0 and st["csr_kernel"] == 1 and st["csr_items"] > 0
    ctx.set_option("force_kind", -1)
    for dsg in (gpu_lib.design(1, 2048, 16), gpu_lib.design(4, 512, 8, arch=1), gpu_lib.design(48, 64, 2)):
        _check_all(golden, gpu_lib, ctx, dsg, exact=False)


def test_merge_path_edge_cases(gpu_lib, ctx, oracle):
    """Empty matrices, rows of zero length in long stretches, one row far longer than a tile (4352 merge items), rectangular."""
    ctx.set_option("csr_kernel", 1)
    ctx.set_option("force_kind", 1)
    d = gpu_lib.design(2, 64, 4)
    rng = np.random.default_rng(11)
    cases = [(5, 7, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0)),
             (1, 9, np.array([0, 9], np.int32), np.arange(9, dtype=np.int32), np.arange(1.0, 10.0))]
    import scipy.sparse as sp
    for n, m, hub_rows, hub_len, p_empty in ((3000, 50000, (7, 1500, 2999), 30000, 0.7), (20000, 20000, (0,), 19000, 0.95),
                                             (9000, 300, (), 0, 0.0), (4353, 4353, (4352,), 4353, 0.5)):
        lens = np.where(rng.random(n) < p_empty, 0, rng.integers(1, 12, n))
        lens = np.minimum(lens, m)
        rows = np.repeat(np.arange(n), lens)
        cols = rng.integers(0, m, len(rows))
        a = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, m)).tolil()
        for r in hub_rows:
            a[r, rng.choice(m, hub_len, replace=False)] = 1.0
        a = a.tocsr()
        a.sort_indices()
        a.data = rng.standard_normal(len(a.data))
        cases.append((n, m, a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)))
    for n, m, rp, ci, va in cases:
        x = rng.standard_normal(m)
        ctx.preprocess(d, n, m, rp, ci, va)
        assert ctx.plan_stats()["csr_kernel"] == 1
        exp = oracle.csr_dot(n, rp, ci, va, x)
        got = ctx.spmv(x)
        assert_y_close(got, exp, row_scale(n, rp, ci, va, x))
        assert np.array_equal(got[np.diff(rp) == 0], np.zeros(int((np.diff(rp) == 0).sum())))   # empty rows are written, as zeros
        assert np.array_equal(got, ctx.spmv(x))                                                 # deterministic


def test_merge_path_rmat_matches_row_group_kernel(gpu_lib, ctx, oracle):
    """R-MAT twin (scale 16, 1.4M nonzeros: hubs of 8 000 nonzeros, a third of the rows empty): automatic selection picks the merge kernel from 2^20 gather nonzeros; its
    result agrees with the row-group kernel and the oracle."""
    n, rp, ci, va = oracle.gen_rmat(16, 24, 3)
    x = np.random.default_rng(2).standard_normal(n)
    d = gpu_lib.design(1, 8192, 16)
    ctx.preprocess(d, n, n, rp, ci, va)
    st = ctx.plan_stats()
    assert st["csr_nnz"] >= 1 << 20 and st["csr_kernel"] == 1, st
    y_merge = ctx.spmv(x)
    ctx.set_option("csr_kernel", 0)
    ctx.preprocess(d, n, n, rp, ci, va)
    assert ctx.plan_stats()["csr_kernel"] == 0
    y_rows = ctx.spmv(x)
    scale = row_scale(n, rp, ci, va, x)
    assert_y_close(y_merge, oracle.csr_dot(n, rp, ci, va, x), scale)
    assert_y_close(y_merge, y_rows, scale)


def test_merge_path_fused_dot_in_cg(gpu_lib, ctx, oracle):
    """kDot instantiation: CG on an SPD system forced through the merge kernel stops where the row-group kernel and the
    restated pcg stop (+-1) and reaches the same solution."""
    n, rp, ci, va = oracle.gen_poisson3d27(14)
    b = oracle.csr_dot(n, rp, ci, va, 1.0 + 0.25 * (np.arange(n) % 4))
    ctx.set_option("force_kind", 1)
    res = {}
    for k in (0, 1):
        ctx.set_option("csr_kernel", k)
        ctx.preprocess(gpu_lib.design(1, 8192, 16), n, n, rp, ci, va)
        assert ctx.plan_stats()["csr_kernel"] == k and ctx.plan_stats()["slices_staged_ell"] == 0
        res[k] = ctx.cg(b)
    oc, oi, ox, _ = oracle.pcg(n, rp, ci, va, b, lower=False)
    assert res[1][0] and abs(res[1][1] - oi) <= 1 and abs(res[1][1] - res[0][1]) <= 1
    assert np.allclose(res[1][2], ox, rtol=0, atol=1e-6)


# ---- hub clustering of the gather path (option col_reorder, off by default) -----------------------------------------
def test_col_reorder_all_fixtures_and_edge_cases(golden, gpu_lib, ctx, oracle):
    """Columns renumbered by descending reference count + x permuted in front of every SpMV: same y (tolerance of the
    gather kernels) on every fixture, on rectangular matrices whose referenced-column count is odd, on matrices most of
    whose columns are never referenced, and with a fused dot inside CG."""
    ctx.set_option("csr_kernel", 1)
    ctx.set_option("force_kind", 1)
    ctx.set_option("col_reorder", 1)
    _check_all(golden, gpu_lib, ctx, gpu_lib.design(3, 2048, 16), exact=False)
    st = ctx.plan_stats()
    assert st["col_reorder"] == 1 and 0 < st["cols_referenced"] <= st["m"], st
    ctx.set_option("force_kind", -1)   # mixed plans: staged slices read x, gather slices the permuted copy
    _check_all(golden, gpu_lib, ctx, gpu_lib.design(4, 512, 8, arch=1), exact=False)
    ctx.set_option("force_kind", 1)
    rng = np.random.default_rng(5)
    import scipy.sparse as sp
    for n, m, used in ((700, 90001, 333), (5000, 4097, 4097), (1, 9, 9), (3000, 3000, 1)):
        cols_pool = rng.choice(m, used, replace=False)
        lens = rng.integers(0, 9, n)
        rows = np.repeat(np.arange(n), lens)
        a = sp.csr_matrix((np.ones(len(rows)), (rows, cols_pool[rng.integers(0, used, len(rows))])), shape=(n, m))
        a.sum_duplicates()
        a.sort_indices()
        a.data = rng.standard_normal(len(a.data))
        rp, ci, va = a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)
        x = rng.standard_normal(m)
        ctx.preprocess(gpu_lib.design(2, 64, 4), n, m, rp, ci, va)
        st = ctx.plan_stats()
        if len(ci):
            assert st["col_reorder"] == 1 and st["cols_referenced"] == len(np.unique(ci)), st
        got = ctx.spmv(x)
        assert_y_close(got, oracle.csr_dot(n, rp, ci, va, x), row_scale(n, rp, ci, va, x))
        x2 = rng.standard_normal(m)                                        # x is re-permuted on every call
        assert_y_close(ctx.spmv(x2), oracle.csr_dot(n, rp, ci, va, x2), row_scale(n, rp, ci, va, x2))
        assert np.array_equal(got, ctx.spmv(x))
    n, rp, ci, va = oracle.gen_poisson3d27(12)
    b = oracle.csr_dot(n, rp, ci, va, 1.0 + 0.25 * (np.arange(n) % 4))
    ctx.preprocess(gpu_lib.design(1, 8192, 16), n, n, rp, ci, va)
    assert ctx.plan_stats()["col_reorder"] == 1
    conv, iters, sol, _ = ctx.cg(b)
    oc, oi, ox, _ = oracle.pcg(n, rp, ci, va, b, lower=False)
    assert conv and abs(iters - oi) <= 1 and np.allclose(sol, ox, rtol=0, atol=1e-6)


def test_col_reorder_rmat(gpu_lib, ctx, oracle):
    """R-MAT twin: the reordered plan agrees with the plain merge kernel, the oracle, and puts the hubs first."""
    n, rp, ci, va = oracle.gen_rmat(16, 24, 3)
    x = np.random.default_rng(2).standard_normal(n)
    d = gpu_lib.design(1, 8192, 16)
    ctx.preprocess(d, n, n, rp, ci, va)
    st = ctx.plan_stats()
    assert st["csr_kernel"] == 1 and st["col_reorder"] == 0, st     # opt-in
    y0 = ctx.spmv(x)
    ctx.set_option("col_reorder", 1)
    ctx.preprocess(d, n, n, rp, ci, va)
    st = ctx.plan_stats()
    assert st["col_reorder"] == 1 and st["cols_referenced"] == len(np.unique(ci)), st
    y1 = ctx.spmv(x)
    scale = row_scale(n, rp, ci, va, x)
    assert_y_close(y1, oracle.csr_dot(n, rp, ci, va, x), scale)
    assert_y_close(y1, y0, scale)


# ---- pageable caller buffers: the library's own pinned rings + host copy threads (hostcopy.hpp) -----------------------
def test_pageable_host_buffers_are_staged_by_the_library(gpu_lib, ctx, oracle):
    """numpy arrays are pageable memory, like the std::vector<double> inside every cask::Vector: from 4 MB of vectors on,
    cask_b200_spmv moves them through pinned 4 MB ring slots with its own copy threads.  Same y, bit for bit, as with the
    driver's staging (host_staging = 0) and as the reference's dot - on the chunked pipeline (staged ELL), on a one-chunk
    plan (merge-path gather) and with ring wrap-around in both directions (9.7 MB per vector = 3 slots of 4)."""
    n, rp, ci, va = oracle.gen_poisson2d(1100)
    x = np.random.default_rng(8).standard_normal(n)
    exp = oracle.csr_dot(n, rp, ci, va, x)
    d = gpu_lib.design(1, 8192, 16)
    ctx.preprocess(d, n, n, rp, ci, va)
    assert ctx.plan_stats()["slices_gather_csr"] == 0
    y_staged = ctx.spmv(x)
    assert np.array_equal(y_staged, exp)
    for rep in range(3):                                   # rings and events are reused across calls
        x2 = np.random.default_rng(20 + rep).standard_normal(n)
        assert np.array_equal(ctx.spmv(x2), oracle.csr_dot(n, rp, ci, va, x2))
    ctx.set_option("host_staging", 0)
    assert np.array_equal(ctx.spmv(x), y_staged)
    ctx.set_option("host_staging", 1)
    ctx.set_option("force_kind", 1)
    ctx.set_option("csr_kernel", 1)
    ctx.preprocess(d, n, n, rp, ci, va)
    assert ctx.plan_stats()["csr_kernel"] == 1
    assert_y_close(ctx.spmv(x), exp, row_scale(n, rp, ci, va, x))
    ctx.set_option("col_reorder", 1)
    ctx.preprocess(d, n, n, rp, ci, va)
    assert_y_close(ctx.spmv(x), exp, row_scale(n, rp, ci, va, x))


def test_col_reorder_automatic_rule(gpu_lib, ctx, oracle):
    """Default options: a single-rank gather plan of >= 2^24 nonzeros over an x of >= 64 MB is hub-clustered when its column
    reference counts are skewed (here: column = n u^6, the most referenced 1/64 of the columns draw over 40 % of the
    gathers; tiles re-cut for 5 items per thread) and left alone when they are not (uniformly scattered columns: the
    permutation of x would be pure cost).  Same y either way - one nonzero per row, so y is exact."""
    n = 1 << 24
    rng = np.random.default_rng(6)
    rp = np.arange(n + 1, dtype=np.int32)
    va = rng.standard_normal(n)
    x = rng.standard_normal(n)
    d = gpu_lib.design(1, 8192, 16)
    for skewed in (True, False):
        u = rng.random(n)
        ci = np.minimum((n * (u ** 6 if skewed else u)).astype(np.int64), n - 1).astype(np.int32)
        ctx.set_option("col_reorder", -1)
        ctx.preprocess(d, n, n, rp, ci, va)
        st = ctx.plan_stats()
        assert st["csr_kernel"] == 1 and st["slices_staged_ell"] == 0, st
        assert st["col_reorder"] == (1 if skewed else 0), st
        if skewed:
            assert st["cols_referenced"] == len(np.unique(ci))
        assert np.array_equal(ctx.spmv(x), va * x[ci])
        ctx.set_option("col_reorder", 0)
        ctx.preprocess(d, n, n, rp, ci, va)
        assert ctx.plan_stats()["col_reorder"] == 0 and np.array_equal(ctx.spmv(x), va * x[ci])
