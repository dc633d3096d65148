"""ILU(0) / Jacobi device logic of libcask_b200.so (cask_b200/csrc/precond_logic.inl) executed through the host
emulation of its backend (tests/emu): level analysis, the level-by-level factorisation and both triangular solves,
rows of a level visited in scrambled order, against the oracle's sequential restatement of ILUPreconditioner
(itself pinned to the reference's known answers and compiled code in tests/test_oracle_precond.py).  Bit-exact.
CPU only; the GPU run of the same logic is tests/test_gpu_w_precond.py."""
import numpy as np
import pytest

import emu
from test_oracle_precond import ARROW, dense_to_csr, _random_spd


def run(oracle, n, rp, ci, va, seed=0):
    x = np.random.default_rng(seed).standard_normal(n)
    pc = oracle.ilu0(n, rp, ci, va)
    for order in (0, 1, 2):
        for unit in (False, True):
            r = emu.ilu(n, rp, ci, va, x, unit, order)
            assert r["rc"] == 0, r["message"]
            z, bad = oracle.ilu_apply(n, rp, ci, pc, x, unit)
            assert np.array_equal(r["pc"], pc)
            assert np.array_equal(r["z"], z, equal_nan=True)
            assert r["zero_pivot"] == bad
    assert emu.lib().emu_live_allocations() == 0
    return r


@pytest.mark.parametrize("gen,arg,levels", [("gen_poisson2d", 24, 47), ("gen_poisson3d27", 7, 43), ("gen_convdiff3d7", 9, 25)])
def test_stencils_bit_exact_and_level_counts(oracle, gen, arg, levels):
    n, rp, ci, va = getattr(oracle, gen)(arg)
    r = run(oracle, n, rp, ci, va)
    # 5-point: level = i + j (2N - 1 wavefronts); 27-point: x + 2y + 4z... (7(N-1) + 1); 7-point: x + y + z (3(N-1) + 1)
    assert r["levels_lower"] == levels and r["levels_upper"] == levels


def test_reference_known_answers():
    """test/LinearSolvers.cpp:79-99 (ILUCompute2) and :125-146 (ILUComputeAndApply) through the device logic."""
    n, rp, ci, va = dense_to_csr(ARROW)
    r = emu.ilu(n, rp, ci, va, [1, 2, 3, 4])
    assert r["pc"].tolist() == [2, 1, 1, 1, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]
    assert r["z"].tolist() == [-16.25, 7, 11, 15]
    assert (r["levels_lower"], r["levels_upper"]) == (2, 2)


def test_random_patterns_lower_only_and_zero_pivots(oracle):
    rng = np.random.default_rng(9)
    for trial in range(30):
        a = _random_spd(rng, int(rng.integers(1, 50)), float(rng.uniform(0.05, 0.5)))
        if trial % 3 == 0:
            a = np.tril(a)                      # what the reference's pcg hands the constructor
        if trial % 5 == 0 and a.shape[0] > 2:
            a[1, 1] = 0.0                       # missing diagonal -> flagged, same infinities as the oracle
        if trial % 7 == 0:
            a = a * (rng.random(a.shape) < 0.7)  # non-symmetric pattern
        n, rp, ci, va = dense_to_csr(a)
        run(oracle, n, rp, ci, va, trial)


def test_rmat_power_law_pattern(oracle):
    n, rp, ci, va = oracle.gen_rmat(9, 6, 3)
    run(oracle, n, rp, ci, va)


def test_unsorted_rows_are_refused():
    r = emu.ilu(2, [0, 2, 3], [1, 0, 1], [1.0, 2.0, 3.0])
    assert r["rc"] == 5 and "ascending" in r["message"]  # CASK_B200_ERR_UNSUPPORTED
    r = emu.ilu(2, [0, 2, 3], [0, 0, 1], [1.0, 2.0, 3.0])
    assert r["rc"] == 5
    assert emu.lib().emu_live_allocations() == 0


def test_empty_matrix_and_empty_rows():
    r = emu.ilu(0, [0], [], [], [])
    assert r["rc"] == 0
    r = emu.ilu(3, [0, 0, 1, 1], [1], [2.0], [1.0, 4.0, 5.0], unit_lower=False)
    assert r["rc"] == 0 and r["zero_pivot"] and r["z"][1] == 1.0  # (4 / 2) / 2; rows 0 and 2 have no pivot


def test_jacobi_inverse_diagonal(oracle):
    n, rp, ci, va = oracle.gen_poisson3d27(5)
    assert np.array_equal(emu.inv_diag(n, rp, ci, va), np.full(n, 1.0 / 26.0))
    n, rp, ci, va = dense_to_csr([[0, 1], [1, 4]])
    assert emu.inv_diag(n, rp, ci, va).tolist() == [1.0, 0.25]
