"""Pins the solver restatements to the reference's code running on the REAL Intel MKL.

The reference's CPU solver path is written against MKL (pcg<> -> mkl_dcsrsymv / cblas_*; ILUPreconditioner::apply ->
mkl_dcsrtrsv).  The image has no MKL package, but libtorch_cpu.so exports oneMKL's inspector-executor sparse BLAS;
oracle/ref_shim_mkl/mkl.h maps the two NIST-style names onto it and oracle/_ref/libcaskref_mkl.so is the reference's
own pcg<> / ILUPreconditioner compiled against that header (oracle/mklbind.py).  What this file establishes:

  * the adapter's enum values / calling convention are right (checked by behaviour against numpy);
  * MKL's y = A x agrees with the oracle's CsrMatrix::dot restatement to 1e-12 (north_star's bar for y);
  * the reference's known answers (test/LinearSolvers.cpp:14-146) hold with MKL's arithmetic;
  * the reference's ILU apply really is a NON-unit lower solve with MKL's own trsv (the finding recorded in
    DESIGN.md section 2 does not depend on the stand-in);
  * iteration counts of the reference loop on MKL are within +-1 of the restated loop - the bar north_star sets for CG.
CPU only."""
import numpy as np
import pytest

from oracle import mklbind as M
from test_oracle_precond import ARROW, _random_spd, csr_to_dense, dense_to_csr, ulp_close

needs_mkl = pytest.mark.skipif(not M.available(), reason="libtorch_cpu.so does not export MKL's sparse BLAS here")
needs_ref_mkl = pytest.mark.skipif(not M.ref_available(), reason="oracle/_ref/libcaskref_mkl.so not built (needs /root/reference)")


@needs_mkl
def test_mkl_identifies_itself():
    assert "Math Kernel Library" in M.version()
    assert M.max_threads() >= 1


@needs_mkl
def test_adapter_enums_by_behaviour():
    """General / symmetric-lower products and the four triangular solves against dense numpy."""
    rng = np.random.default_rng(11)
    a = _random_spd(rng, 23, 0.3)
    n, rp, ci, va = dense_to_csr(a)
    x = rng.standard_normal(n)
    h = M.CsrHandle(n, n, rp, ci, va)
    assert np.allclose(h.spmv(x), a @ x, rtol=1e-13, atol=1e-13)
    # only the stored lower triangle is read: feed a matrix whose upper triangle is garbage
    junk = np.tril(a) + np.triu(rng.standard_normal((n, n)), 1)
    nj, rpj, cij, vaj = dense_to_csr(junk)
    assert np.allclose(M.CsrHandle(nj, nj, rpj, cij, vaj).symv_lower(x), a @ x, rtol=1e-13, atol=1e-13)
    # and the lower triangle alone gives the full symmetric product (what the reference's pcg relies on)
    nl, rpl, cil, val = dense_to_csr(np.tril(a))
    assert np.allclose(M.CsrHandle(nl, nl, rpl, cil, val).symv_lower(x), a @ x, rtol=1e-13, atol=1e-13)
    for lower in (True, False):
        t = np.tril(a) if lower else np.triu(a)
        for unit in (False, True):
            tt = t.copy()
            if unit:
                np.fill_diagonal(tt, 1.0)
            assert np.allclose(h.trsv(x, lower, unit), np.linalg.solve(tt, x), rtol=1e-10, atol=1e-12), (lower, unit)


@needs_mkl
@pytest.mark.parametrize("gen,arg", [("gen_poisson2d", 64), ("gen_poisson3d27", 12), ("gen_convdiff3d7", 16)])
def test_mkl_product_agrees_with_the_oracle(oracle, gen, arg):
    n, rp, ci, va = getattr(oracle, gen)(arg)
    x = np.random.default_rng(1).random(n)
    exp = oracle.csr_dot(n, rp, ci, va, x)
    got = M.CsrHandle(n, n, rp, ci, va).spmv(x)
    scale = np.bincount(np.repeat(np.arange(n), np.diff(rp)), weights=np.abs(va * x[ci]), minlength=n)
    assert np.all(np.abs(got - exp) <= 1e-12 * np.maximum(np.abs(exp), scale))


@needs_mkl
def test_mkl_product_on_the_reference_fixtures(oracle, golden):
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        if n != m or len(va) == 0:
            continue
        x = golden.x(name)
        exp = golden.dots[name]
        got = M.CsrHandle(n, m, rp, ci, va).spmv(x)
        scale = np.bincount(np.repeat(np.arange(n), np.diff(rp)), weights=np.abs(va * x[ci]), minlength=n)
        assert np.all(np.abs(got - exp) <= 1e-12 * np.maximum(np.abs(exp), scale)), name


@needs_ref_mkl
@pytest.mark.parametrize("name,sol", [("tiny", [1, 2, 3, 4]), ("tinysym", [-2, 2, 3, 3])])
def test_reference_cg_known_answers_on_mkl(golden, name, sol):
    """CGWithIdentityPC / CGSymWithIdentityPC, test/LinearSolvers.cpp:14-52, the reference's loop on MKL."""
    s = golden.systems[name]
    conv, it, x, sec = M.pcg(s["n"], s["row_ptr"], s["col_ind"], s["values"], s["rhs"], precon=0)
    assert conv and ulp_close(x, sol) and sec >= 0.0


@needs_ref_mkl
def test_reference_ilu_known_answers_on_mkl(oracle, golden):
    """ILUComputeAndApply (:125-146) and CGSymWithILUPC (:54-77) with MKL's trsv / symv / ddot."""
    n, rp, ci, va = dense_to_csr(ARROW)
    assert ulp_close(M.ilu_apply(n, rp, ci, va, [1.0, 2.0, 3.0, 4.0]), [-16.25, 7, 11, 15])
    s = golden.systems["tinysym"]
    conv, it, x, _ = M.pcg(s["n"], s["row_ptr"], s["col_ind"], s["values"], s["rhs"], precon=1)
    exp = [-1.9982580059252246, 2.0000862488691915, 3.0001293733037859, 2.9987581910958183]
    assert (conv, it) == (False, 1999)  # the loop never converges; the asserted doubles are iterate 2000
    assert ulp_close(x, exp)
    # the restatement lands within 4 ulp of the same asserted doubles (so within 8 ulp of MKL's iterate)
    oc, oit, ox, _ = oracle.pcg_precond(s["n"], s["row_ptr"], s["col_ind"], s["values"], s["rhs"], "ilu", lower=True)
    assert (oc, oit) == (conv, it) and ulp_close(ox, exp) and ulp_close(ox, x, ulps=8)


@needs_ref_mkl
def test_reference_ilu_apply_is_a_non_unit_lower_solve_on_mkl(oracle):
    """apply() = U^-1 (D + L)^-1 x with MKL's own mkl_sparse_d_trsv: equal to the oracle's non-unit variant to a few
    ulp of the solve, and far from the unit-lower (textbook) variant."""
    rng = np.random.default_rng(4)
    for a in [_random_spd(rng, 31, 0.2), _random_spd(rng, 12, 0.5)]:
        n, rp, ci, va = dense_to_csr(a)
        x = rng.standard_normal(n)
        pc = oracle.ilu0(n, rp, ci, va)
        z_mkl = M.ilu_apply(n, rp, ci, va, x)
        z_non_unit, bad0 = oracle.ilu_apply(n, rp, ci, pc, x, False)
        z_unit, bad1 = oracle.ilu_apply(n, rp, ci, pc, x, True)
        assert not bad0 and not bad1
        assert np.allclose(z_mkl, z_non_unit, rtol=1e-12, atol=1e-14)
        assert np.abs(z_mkl - z_unit).max() > 1e-3 * np.abs(z_unit).max()


@needs_ref_mkl
def test_reference_ilu_pcg_stalls_on_an_spd_stencil_on_mkl(oracle):
    """With MKL's arithmetic too, the reference's ILU-preconditioned CG does not converge on an SPD stencil when the
    ILU is built from the matrix itself (pcg handed the full symmetric CSR: mkl_dcsrsymv('l') reads its lower triangle,
    ILUPreconditioner{a} factors all of it) - while identity-preconditioned CG on the same call converges."""
    n, rp, ci, va = oracle.gen_poisson3d27(6)
    xt = 1.0 + 0.25 * (np.arange(n) % 4)
    b = oracle.csr_dot(n, rp, ci, va, xt)
    conv, it, x, _ = M.pcg(n, rp, ci, va, b, precon=1)
    assert (conv, it) == (False, 1999) and np.abs(x - xt).max() > 1e-3
    oc, oit, ox, _ = oracle.pcg_precond(n, rp, ci, va, b, "ilu", lower=False)
    assert (oc, oit) == (False, 1999)
    conv0, it0, x0, _ = M.pcg(n, rp, ci, va, b, precon=0)
    assert conv0 and it0 < 100 and np.abs(x0 - xt).max() < 1e-4


@needs_ref_mkl
def test_cg_iteration_counts_on_mkl_within_one_of_the_restatement(oracle):
    """north_star's CG bar (iteration counts within +-1), here between the reference loop on MKL and the C restatement
    the GPU suites compare against - on random SPD systems and on the down-scaled twins of BASELINE's stencils."""
    rng = np.random.default_rng(9)
    cases = []
    for _ in range(5):
        a = _random_spd(rng, int(rng.integers(20, 120)), 0.1)
        n, rp, ci, va = dense_to_csr(np.tril(a))
        cases.append((n, rp, ci, va, rng.standard_normal(n)))
    for gen, arg in (("gen_poisson2d", 24), ("gen_poisson3d27", 10)):
        n, rp, ci, va = getattr(oracle, gen)(arg)
        b = oracle.csr_dot(n, rp, ci, va, 1.0 + 0.25 * (np.arange(n) % 4))
        keep = ci <= np.repeat(np.arange(n), np.diff(rp))
        rows = np.repeat(np.arange(n), np.diff(rp))[keep]
        rpl = np.zeros(n + 1, np.int32)
        rpl[1:] = np.cumsum(np.bincount(rows, minlength=n))
        cases.append((n, rpl, ci[keep], va[keep], b))
    for n, rp, ci, va, b in cases:
        conv_m, it_m, x_m, _ = M.pcg(n, rp, ci, va, b, precon=0)
        conv_o, it_o, x_o, _ = oracle.pcg_precond(n, rp, ci, va, b, "identity", lower=True)
        assert conv_m and conv_o
        assert abs(it_m - it_o) <= 1, (n, it_m, it_o)
        assert np.allclose(x_m, x_o, rtol=0, atol=1e-6)
