"""Coded staged ELL (option value_dict, cask_b200_plan_value_dict): 8-bit codes into per-slice tables of the distinct
values, 3 bytes per stored nonzero instead of 10 (mode 1), or into tables of (value, x-cache displacement) pairs, 1 byte
per stored nonzero and no index stream (mode 2; the tests at the end of the file).  Bar: y, CG and BiCGStab results IDENTICAL, bit for bit, to the
uncoded kernel on the same plan (same doubles multiplied in the same order), and therefore the same parity with the
reference's CsrMatrix::dot as tests/test_gpu_spmv.py; matrices with more than 256 distinct values per slice silently
keep the uncoded format.  The table builder itself is checked on CPU in tests/test_valuedict_emu.py.
File name sorts after the other GPU suites on purpose: written after the last GPU session of round 1, off by default."""
import numpy as np
import pytest

from conftest import assert_y_close, row_scale

pytestmark = pytest.mark.gpu

STENCILS = [("gen_poisson2d", 96, 3), ("gen_poisson2d", 257, 3), ("gen_poisson3d27", 24, 3), ("gen_convdiff3d7", 20, 6)]


def both(gpu_lib, ctx, dsg, n, m, rp, ci, va, fn, mode=1, **opts):
    """fn(ctx) on the uncoded and on the coded plan of the same matrix; returns (plain, coded, value_dict info)."""
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_option("value_dict", 0)
    ctx.preprocess(dsg, n, m, rp, ci, va)
    assert ctx.value_dict()[0] is False
    plain = fn(ctx)
    ctx.set_option("value_dict", mode)
    ctx.preprocess(dsg, n, m, rp, ci, va)
    info = ctx.value_dict()
    coded = fn(ctx)
    ctx.set_option("value_dict", 0)
    return plain, coded, info


@pytest.mark.parametrize("ku", [0, 2, 4])
@pytest.mark.parametrize("gen,N,max_entries", STENCILS)
def test_stencils_are_coded_and_bit_identical(gpu_lib, ctx, oracle, gen, N, max_entries, ku):
    n, rp, ci, va = getattr(oracle, gen)(N)
    x = np.random.default_rng(1).standard_normal(n)
    # the z = N-1 boundary slice of the 24^3 27-point grid has ELL fill 0.706, below the planner's default 0.75
    # (plan.cu: ell_min_fill): lowered here so that every slice is staged and therefore coded
    plain, coded, (active, entries, mbytes) = both(gpu_lib, ctx, gpu_lib.design(2, 8192, 16), n, n, rp, ci, va,
                                                   lambda c: c.spmv(x), persist_ku=ku, ell_min_fill=0.5)
    st = ctx.plan_stats()
    ctx.set_option("ell_min_fill", 0.75)
    ctx.set_option("persist_ku", 0)
    assert active and 2 <= entries <= max_entries + 1 and entries % 2 == 0
    assert st["slices_gather_csr"] == 0
    assert mbytes == 3 * st["ell_padded_entries"] + 8 * entries * st["slices_staged_ell"]
    assert np.array_equal(plain, coded)
    assert np.array_equal(coded, oracle.csr_dot(n, rp, ci, va, x))   # the reference's summation order


@pytest.mark.parametrize("ctas", [1, 3])
def test_ctas_per_sm_override(gpu_lib, ctx, oracle, ctas):
    n, rp, ci, va = oracle.gen_poisson2d(300)
    x = np.random.default_rng(2).standard_normal(n)
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va, lambda c: c.spmv(x),
                              persist_ctas=ctas)
    ctx.set_option("persist_ctas", 0)
    assert info[0] and np.array_equal(plain, coded)


def test_signed_zero_nan_and_many_values(gpu_lib, ctx, oracle):
    """Tables hold bit patterns: -0.0, NaN and infinities survive; 256 distinct values per slice still fit."""
    n, rp, ci, va = oracle.gen_poisson2d(64)
    va = va.copy()
    va[::7] = -0.0
    va[5::11] = (np.arange(len(va[5::11])) % 250) + 0.5       # + {4, -1, -0.0}: at most 253 distinct values per slice
    x = np.random.default_rng(3).standard_normal(n)
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va, lambda c: c.spmv(x))
    assert info[0] and 200 <= info[1] <= 256
    assert np.array_equal(plain.view(np.uint64), coded.view(np.uint64))
    va[3] = np.nan
    va[9] = np.inf
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va, lambda c: c.spmv(x))
    assert info[0] and np.array_equal(plain.view(np.uint64), coded.view(np.uint64))


def test_too_many_values_keep_the_uncoded_format(golden, gpu_lib, ctx, oracle):
    n, rp, ci, va = oracle.gen_poisson2d(96)
    va = np.random.default_rng(4).standard_normal(len(va))
    x = np.random.default_rng(5).standard_normal(n)
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va, lambda c: c.spmv(x))
    assert info[0] is False and info[1] == 0 and np.array_equal(plain, coded)
    assert info[2] == 10 * ctx.plan_stats()["ell_padded_entries"]
    # every reference fixture under the option: coded where a slice allows it, uncoded elsewhere; same y either way and
    # the same parity bar against the reference's CsrMatrix::dot
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        x = golden.x(name)
        plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(2, 24576, 16), n, m, rp, ci, va, lambda c: c.spmv(x))
        assert np.array_equal(plain.view(np.uint64), coded.view(np.uint64)), (name, info)
        assert_y_close(coded, golden.dots[name], row_scale(n, rp, ci, va, x))


def test_mixed_plan_gather_slices_beside_coded_slices(gpu_lib, ctx, oracle):
    """A few long rows push their slices to the gather-CSR kernel; the staged slices stay coded."""
    n, rp, ci, va = oracle.gen_poisson2d(80)
    import scipy.sparse as sp
    a = sp.csr_matrix((va, ci, rp), shape=(n, n)).tolil()
    rng = np.random.default_rng(6)
    for r in (10, 2500, 2501):
        cols = rng.choice(n, 900, replace=False)
        a[r, cols] = 2.0
    a = a.tocsr()
    a.sort_indices()
    rp2, ci2, va2 = a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)
    x = rng.standard_normal(n)
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp2, ci2, va2, lambda c: c.spmv(x))
    st = ctx.plan_stats()
    assert st["slices_gather_csr"] > 0 and st["slices_staged_ell"] > 0 and info[0]
    assert np.array_equal(plain, coded)
    assert_y_close(coded, oracle.csr_dot(n, rp2, ci2, va2, x), row_scale(n, rp2, ci2, va2, x))


@pytest.mark.parametrize("gen,N", [("gen_poisson2d", 120), ("gen_poisson3d27", 20)])
def test_cg_identical_iterates(gpu_lib, ctx, oracle, gen, N):
    """The fused SpMV + p.Ap kernel in its coded instantiation: same iterations, same solution bits."""
    n, rp, ci, va = getattr(oracle, gen)(N)
    b = oracle.csr_dot(n, rp, ci, va, 1.0 + 0.25 * (np.arange(n) % 4))
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va, lambda c: c.cg(b))
    assert info[0]
    assert plain[0] and coded[0] and plain[1] == coded[1] and plain[3] == coded[3]
    assert np.array_equal(plain[2], coded[2])
    oc, oi, ox, _ = oracle.pcg(n, rp, ci, va, b, lower=False)
    assert abs(coded[1] - oi) <= 1


def test_bicgstab_identical_iterates(gpu_lib, ctx, oracle):
    n, rp, ci, va = oracle.gen_convdiff3d7(20)
    b = oracle.csr_dot(n, rp, ci, va, np.ones(n))
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va,
                              lambda c: c.bicgstab(b, tol=1e-10))
    assert info[0] and plain[1] == coded[1] and np.array_equal(plain[0], coded[0])


def test_full_size_c2_properties(gpu_lib, ctx):
    """BASELINE configs[1] at full size under the coded format: closed-form result for x = 0.25 (k mod 1024)."""
    import torch
    G = 4096
    n = G * G
    dev = torch.device("cuda", 0)
    nnz = gpu_lib.synth_nnz(gpu_lib.SYNTH_POISSON2D, G, 0, n)
    rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    va = torch.empty(nnz, dtype=torch.float64, device=dev)
    gpu_lib.synth_device(gpu_lib.SYNTH_POISSON2D, G, 0, n, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), 0)
    torch.cuda.synchronize()
    ctx.set_option("value_dict", 1)
    ctx.preprocess_device(gpu_lib.design(1, 8192, 16), n, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    active, entries, mbytes = ctx.value_dict()
    assert active and entries == 4 and mbytes < 0.31 * 10 * ctx.plan_stats()["ell_padded_entries"]
    x = ((torch.arange(n, device=dev) % 1024).double() * 0.25).contiguous()
    y = torch.empty(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()          # the context runs on its own non-blocking stream
    ctx.spmv_device(x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    xg = x.view(G, G)
    ref = 4 * xg.clone()
    ref[:, 1:] -= xg[:, :-1]
    ref[:, :-1] -= xg[:, 1:]
    ref[1:] -= xg[:-1]
    ref[:-1] -= xg[1:]
    assert torch.equal(y.view(G, G), ref)
    ctx.set_option("value_dict", 0)


# ---- mode 2: (value, displacement) pair codes ----------------------------------------------------------------------
PAIR_STENCILS = [("gen_poisson2d", 96, 5), ("gen_poisson2d", 257, 5), ("gen_poisson3d27", 24, 27), ("gen_convdiff3d7", 20, 7)]


@pytest.mark.parametrize("ku", [0, 2, 4])
@pytest.mark.parametrize("gen,N,points", PAIR_STENCILS)
def test_pair_codes_stencils_bit_identical(gpu_lib, ctx, oracle, gen, N, points, ku):
    """One pair per stencil point (+ slot 0 for padding), whatever the slice; y identical to the uncoded kernel and to
    the reference's summation order."""
    n, rp, ci, va = getattr(oracle, gen)(N)
    x = np.random.default_rng(1).standard_normal(n)
    plain, coded, (active, entries, mbytes) = both(gpu_lib, ctx, gpu_lib.design(2, 8192, 16), n, n, rp, ci, va,
                                                   lambda c: c.spmv(x), mode=2, persist_ku=ku, ell_min_fill=0.5)
    st = ctx.plan_stats()
    assert active and entries % 8 == 0 and points + 1 <= entries <= ((2 * points + 1 + 7) & ~7)
    assert st["slices_gather_csr"] == 0
    assert mbytes == st["ell_padded_entries"] + 16 * entries * st["slices_staged_ell"]
    assert np.array_equal(plain, coded)
    assert np.array_equal(coded, oracle.csr_dot(n, rp, ci, va, x))
    ctx.set_option("persist_ku", 0)


def test_pair_codes_mode_is_reported_and_falls_back(gpu_lib, ctx, oracle):
    n, rp, ci, va = oracle.gen_poisson2d(96)
    x = np.random.default_rng(5).standard_normal(n)
    dsg = gpu_lib.design(1, 8192, 16)
    plain, coded, info = both(gpu_lib, ctx, dsg, n, n, rp, ci, va, lambda c: (c.value_dict_mode(), c.spmv(x)), mode=2)
    assert plain[0] == 0 and coded[0] == 2 and np.array_equal(plain[1], coded[1])
    # few distinct VALUES but a scattered pattern: every row of a slice reads its own far column, so the slice holds one
    # displacement per row (> 255 pairs) -> value codes
    rng = np.random.default_rng(9)
    rows = np.arange(n)
    far = (rows * 37 + 11) % n
    import scipy.sparse as sp
    a = (sp.csr_matrix((va, ci, rp), shape=(n, n)) + sp.csr_matrix((np.full(n, 0.5), (rows, far)), shape=(n, n))).tocsr()
    a.sort_indices()
    rp2, ci2, va2 = a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)
    ctx.set_option("force_kind", 0)      # staged ELL wherever the x windows fit: n = 9 216 columns do
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 24576, 16), n, n, rp2, ci2, va2,
                              lambda c: (c.value_dict_mode(), c.plan_stats()["slices_staged_ell"], c.spmv(x)), mode=2)
    ctx.set_option("force_kind", -1)
    assert coded[1] > 0 and coded[0] == 1, coded[:2]
    assert np.array_equal(plain[2], coded[2])
    # random values on top: no coding at all
    va3 = rng.standard_normal(len(va2))
    ctx.set_option("force_kind", 0)
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 24576, 16), n, n, rp2, ci2, va3,
                              lambda c: (c.value_dict_mode(), c.spmv(x)), mode=2)
    ctx.set_option("force_kind", -1)
    assert coded[0] == 0 and np.array_equal(plain[1], coded[1])


def test_pair_codes_every_fixture_and_special_values(golden, gpu_lib, ctx, oracle):
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        x = golden.x(name)
        plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(2, 24576, 16), n, m, rp, ci, va, lambda c: c.spmv(x), mode=2)
        assert np.array_equal(plain.view(np.uint64), coded.view(np.uint64)), (name, info)
        assert_y_close(coded, golden.dots[name], row_scale(n, rp, ci, va, x))
    n, rp, ci, va = oracle.gen_poisson2d(64)
    va = va.copy()
    va[::7] = -0.0
    va[3] = np.nan
    va[9] = np.inf
    x = np.random.default_rng(3).standard_normal(n)
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va,
                              lambda c: (c.value_dict_mode(), c.spmv(x)), mode=2)
    assert coded[0] == 2 and np.array_equal(plain[1].view(np.uint64), coded[1].view(np.uint64))


@pytest.mark.parametrize("gen,N", [("gen_poisson2d", 120), ("gen_poisson3d27", 20)])
def test_pair_codes_cg_identical_iterates(gpu_lib, ctx, oracle, gen, N):
    n, rp, ci, va = getattr(oracle, gen)(N)
    b = oracle.csr_dot(n, rp, ci, va, 1.0 + 0.25 * (np.arange(n) % 4))
    plain, coded, info = both(gpu_lib, ctx, gpu_lib.design(1, 8192, 16), n, n, rp, ci, va,
                              lambda c: (c.value_dict_mode(),) + tuple(c.cg(b)), mode=2)
    assert coded[0] == 2
    assert plain[1] and coded[1] and plain[2] == coded[2] and plain[4] == coded[4]
    assert np.array_equal(plain[3], coded[3])


def test_pair_codes_full_size_c2_properties(gpu_lib, ctx):
    """BASELINE configs[1] at full size under pair codes: closed-form result for x = 0.25 (k mod 1024)."""
    import torch
    G = 4096
    n = G * G
    dev = torch.device("cuda", 0)
    nnz = gpu_lib.synth_nnz(gpu_lib.SYNTH_POISSON2D, G, 0, n)
    rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    va = torch.empty(nnz, dtype=torch.float64, device=dev)
    gpu_lib.synth_device(gpu_lib.SYNTH_POISSON2D, G, 0, n, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), 0)
    torch.cuda.synchronize()
    ctx.set_option("value_dict", 2)
    ctx.preprocess_device(gpu_lib.design(1, 8192, 16), n, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    active, entries, mbytes = ctx.value_dict()
    assert ctx.value_dict_mode() == 2 and entries == 8 and mbytes < 0.11 * 10 * ctx.plan_stats()["ell_padded_entries"]
    x = ((torch.arange(n, device=dev) % 1024).double() * 0.25).contiguous()
    y = torch.empty(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()          # the context runs on its own non-blocking stream
    ctx.spmv_device(x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    xg = x.view(G, G)
    ref = 4 * xg.clone()
    ref[:, 1:] -= xg[:, :-1]
    ref[:, :-1] -= xg[:, 1:]
    ref[1:] -= xg[:-1]
    ref[:-1] -= xg[1:]
    assert torch.equal(y.view(G, G), ref)
    ctx.set_option("value_dict", 0)
