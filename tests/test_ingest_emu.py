"""COO -> CSR ingest logic of libcask_b200.so (cask_b200/csrc/ingest_logic.inl), executed through the host emulation
of its device backend (tests/emu) - the same functor bodies and orchestration the GPU runs, visited in scrambled
order - and compared with the reference's DokMatrix semantics.  CPU only; the GPU run of the same logic is
tests/test_gpu_w_ingest.py."""
import os

import numpy as np
import pytest

import emu
from ingest_model import NotSymmetric, dok_ingest
from oracle import refbind as R

ONE, SYM, DROP = emu.INGEST_ONE_BASED, emu.INGEST_SYMMETRIC, emu.INGEST_DROP_UPPER


def check(n, m, rows, cols, vals, flags, order=1):
    out = emu.coo_to_csr(n, m, rows, cols, vals, flags, order)
    assert out["rc"] == 0, out["message"]
    assert emu.lib().emu_live_allocations() == 0
    try:
        rp, ci, va, nnzs = dok_ingest(n, m, rows, cols, vals, bool(flags & SYM), bool(flags & ONE), bool(flags & DROP))
    except NotSymmetric:
        assert out["err"] == 2
        return out
    assert out["err"] == 0
    assert np.array_equal(out["row_ptr"], rp)
    assert np.array_equal(out["col"], ci) and np.array_equal(out["val"], va)
    assert out["nnzs_field"] == nnzs
    return out


def shuffled_coo(golden, name, rng, lower_only=False):
    n, m, rp, ci, va = golden.csr(name)
    rows = np.repeat(np.arange(n), np.diff(rp))
    keep = ci <= rows if lower_only else np.ones(len(ci), bool)
    perm = rng.permutation(int(keep.sum()))
    return n, m, (rows[keep] + 1)[perm], (ci[keep] + 1)[perm], va[keep][perm]


@pytest.mark.parametrize("order", [0, 1, 2])
def test_golden_matrices_from_shuffled_entries(golden, order):
    rng = np.random.default_rng(1)
    for name in golden.names[:: max(1, len(golden.names) // 12)]:
        n, m, rows, cols, vals = shuffled_coo(golden, name, rng)
        out = check(n, m, rows, cols, vals, ONE, order)
        _, _, rp, ci, va = golden.csr(name)
        assert np.array_equal(out["row_ptr"], rp) and np.array_equal(out["col"], ci) and np.array_equal(out["val"], va)


def test_symmetric_expansion_rebuilds_the_full_golden_matrix(golden):
    """tinysym as io::readMatrix gives it = explicitSymmetric of the stored triangle (test/LinearSolvers.cpp:101-123)."""
    s = golden.systems["tinysym"]
    rows = np.repeat(np.arange(s["n"]), np.diff(s["row_ptr"])) + 1
    out = check(s["n"], s["n"], rows, np.array(s["col_ind"]) + 1, s["values"], ONE | SYM)
    assert out["row_ptr"].tolist() == [0, 2, 3, 4, 6] and out["col"].tolist() == [0, 3, 1, 2, 0, 3]
    n, m, rp, ci, va = golden.csr("tinysym")
    assert np.array_equal(out["row_ptr"], rp) and np.array_equal(out["col"], ci) and np.array_equal(out["val"], va)


def test_duplicates_keep_the_last_value_and_count_every_set():
    rows = [1, 2, 1, 3, 1, 2]
    cols = [1, 2, 1, 1, 1, 2]
    vals = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    out = check(3, 3, rows, cols, vals, ONE)
    assert out["val"].tolist() == [5.0, 6.0, 4.0] and out["nnz"] == 3 and out["nnzs_field"] == 6


def test_stored_transpose_pairs():
    # equal values: the reference only warns, keeps both, and counts each visit twice (nnzs 1 + 2 + 2 = 5 for 3 entries)
    out = check(3, 3, [1, 2, 1], [1, 1, 2], [9.0, 4.0, 4.0], ONE | SYM)
    assert out["nnz"] == 3 and out["nnzs_field"] == 5
    # different values: "Matrix is not symmetric"
    out = check(3, 3, [1, 2, 1], [1, 1, 2], [9.0, 4.0, 5.0], ONE | SYM)
    assert out["err"] == 2
    # a duplicate that is overwritten before the comparison does not count as asymmetric
    out = check(3, 3, [2, 1, 2], [1, 2, 1], [7.0, 5.0, 5.0], ONE | SYM)
    assert out["err"] == 0 and out["val"].tolist() == [5.0, 5.0]


def test_bad_indices_are_reported_with_the_first_offender():
    out = emu.coo_to_csr(3, 3, [1, 4, 0, 2], [1, 1, 2, 2], [1.0, 2.0, 3.0, 4.0], ONE)
    assert out["err"] & 1 and out["first_bad"] == 1
    out = emu.coo_to_csr(2, 5, [1, 1], [1, 4], [1.0, 2.0], ONE | SYM)  # mirror (3, 0) does not exist in a 2 x 5 matrix
    assert out["err"] & 1 and out["first_bad"] == 1
    assert emu.lib().emu_live_allocations() == 0


def test_empty_and_degenerate_shapes():
    out = check(5, 5, [], [], [], ONE)
    assert out["row_ptr"].tolist() == [0] * 6 and out["nnz"] == 0
    out = check(6, 4, [3, 3, 5], [4, 1, 2], [1.0, 2.0, 3.0], ONE)   # empty rows at both ends and in the middle
    assert out["row_ptr"].tolist() == [0, 0, 0, 2, 2, 3, 3]
    out = check(1, 1, [0], [0], [2.5], 0)                            # 0-based input
    assert out["val"].tolist() == [2.5]
    out = check(4, 4, [1, 2], [2, 3], [1.0, 1.0], ONE | DROP)        # everything dropped
    assert out["nnz"] == 0 and out["row_ptr"].tolist() == [0] * 5


def test_lower_triangle_expansion_is_the_symmetric_product_pcg_uses(golden):
    """DROP_UPPER | SYMMETRIC on a matrix given in full keeps what mkl_dcsrsymv('l') reads and mirrors it
    (SparseLinearSolvers.hpp:189,206): for a symmetric input that is the input itself."""
    rng = np.random.default_rng(2)
    n, m, rp, ci, va = golden.csr("tinysym")
    rows = np.repeat(np.arange(n), np.diff(rp))
    out = check(n, m, rows, ci, va, SYM | DROP)
    assert np.array_equal(out["row_ptr"], rp) and np.array_equal(out["col"], ci) and np.array_equal(out["val"], va)
    n, m, rows, cols, vals = shuffled_coo(golden, "test_cage6", rng)
    check(n, m, rows, cols, vals, ONE | SYM | DROP)


def test_random_matrices_with_duplicates_all_orders():
    rng = np.random.default_rng(3)
    for trial in range(60):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        L = int(rng.integers(0, 200))
        sym = trial % 3 == 0
        if sym:
            m = n
        rows = rng.integers(1, n + 1, L)
        cols = rng.integers(1, m + 1, L)
        vals = rng.integers(-3, 4, L).astype(float) if trial % 2 else rng.standard_normal(L)
        if sym and trial % 6 == 0:   # make it consistent so that the expansion succeeds: value is a function of the pair
            vals = ((np.minimum(rows, cols) * 31 + np.maximum(rows, cols)) % 7).astype(float)
        flags = ONE | (SYM if sym else 0) | (DROP if trial % 5 == 0 else 0)
        for order in (0, 1, 2):
            check(n, m, rows, cols, vals, flags, order)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_model_matches_the_compiled_reference(tmp_path):
    """The pure-Python model above (and with it the emulated device logic) against io::readMatrix itself, on files
    with repeated keys and stored transpose pairs."""
    rng = np.random.default_rng(4)
    for trial in range(25):
        sym = trial % 2 == 0
        n = int(rng.integers(2, 30))
        m = n if sym else int(rng.integers(2, 30))
        L = int(rng.integers(1, 120))
        rows = rng.integers(1, n + 1, L)
        cols = rng.integers(1, m + 1, L)
        vals = ((np.minimum(rows, cols) * 31 + np.maximum(rows, cols)) % 7 + 1).astype(float) if sym else rng.standard_normal(L)
        if sym and trial % 4 == 0:
            vals = rng.integers(0, 2, L).astype(float)   # likely to clash across the diagonal
        p = str(tmp_path / ("m%d.mtx" % trial))
        with open(p, "w") as f:
            f.write("%%%%MatrixMarket matrix coordinate real %s\n%d %d %d\n" % ("symmetric" if sym else "general", n, m, L))
            for r, c, v in zip(rows, cols, vals):
                f.write("%d %d %r\n" % (r, c, float(v)))
        try:
            rp, ci, va, nnzs = dok_ingest(n, m, rows, cols, vals, sym)
        except NotSymmetric:
            with pytest.raises(RuntimeError) as e:
                R.RefMatrix.read(p)
            assert "Matrix is not symmetric" in str(e.value)
            assert emu.coo_to_csr(n, m, rows, cols, vals, ONE | SYM)["err"] == 2
            continue
        ref = R.RefMatrix.read(p)
        rrp, rci, rva = ref.csr()
        assert np.array_equal(rrp[:-1], rp[:-1]) and rrp[-1] == nnzs   # the reference's row_ptr[n] is its nnzs FIELD
        assert np.array_equal(rci, ci) and np.array_equal(rva, va)
        check(n, m, rows, cols, vals, ONE | (SYM if sym else 0))
