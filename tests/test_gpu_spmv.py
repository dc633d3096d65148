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
