"""GPU ingest (cask_b200_ingest_coo / cask_b200_read_matrix / cask_b200_preprocess_csr) through the C ABI against the
reference's DokMatrix semantics: committed golden CSRs (what io::readMatrix returned for the reference's own files),
the pure-Python model pinned to the compiled reference (tests/ingest_model.py), and y = A x from the oracle.
File name sorts after the other GPU suites on purpose: this path was written after the last GPU session of round 1."""
import numpy as np
import pytest

from ingest_model import NotSymmetric, dok_ingest
from test_mmio_host import coo_of, write_mtx

pytestmark = pytest.mark.gpu


def same(csr, rp, ci, va):
    grp, gci, gva = csr.export()
    assert np.array_equal(grp, rp) and np.array_equal(gci, ci) and np.array_equal(gva, va)


def test_golden_matrices_from_shuffled_entries(gpu_lib, ctx, golden):
    rng = np.random.default_rng(1)
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        rows, cols, vals = coo_of(n, rp, ci, va)
        perm = rng.permutation(len(vals))
        csr = ctx.ingest_coo(n, m, rows[perm], cols[perm], vals[perm], gpu_lib.INGEST_ONE_BASED)
        assert (csr.n, csr.m, csr.nnz, csr.nnzs_field) == (n, m, len(va), len(va))
        same(csr, rp, ci, va)
        csr.free()


def test_read_matrix_symmetric_and_general_files(gpu_lib, ctx, golden, tmp_path):
    # general file
    n, m, rp, ci, va = golden.csr("test_cage6")
    p = str(tmp_path / "g.mtx")
    write_mtx(p, n, m, *coo_of(n, rp, ci, va))
    same(ctx.read_matrix(p), rp, ci, va)
    # symmetric file: lower triangle stored; readMatrix expands it, readSymMatrix keeps it (IO.hpp:151-176)
    s = golden.systems["tinysym"]
    lrp, lci, lva = np.array(s["row_ptr"], np.int32), np.array(s["col_ind"], np.int32), np.array(s["values"], float)
    p = str(tmp_path / "s.mtx")
    write_mtx(p, s["n"], s["n"], *coo_of(s["n"], lrp, lci, lva), symmetry="symmetric")
    n, m, rp, ci, va = golden.csr("tinysym")
    same(ctx.read_matrix(p), rp, ci, va)
    same(ctx.read_matrix(p, sym_lower=True), lrp, lci, lva)
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.read_matrix(str(tmp_path / "g.mtx"), sym_lower=True)
    assert "is not symmetric" in e.value.message


def test_duplicates_transposes_and_errors(gpu_lib, ctx):
    ONE, SYM, DROP = gpu_lib.INGEST_ONE_BASED, gpu_lib.INGEST_SYMMETRIC, gpu_lib.INGEST_DROP_UPPER
    csr = ctx.ingest_coo(3, 3, [1, 2, 1, 3, 1, 2], [1, 2, 1, 1, 1, 2], [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], ONE)
    assert csr.export()[2].tolist() == [5.0, 6.0, 4.0] and (csr.nnz, csr.nnzs_field) == (3, 6)
    csr = ctx.ingest_coo(3, 3, [1, 2, 1], [1, 1, 2], [9.0, 4.0, 4.0], ONE | SYM)
    assert (csr.nnz, csr.nnzs_field) == (3, 5)
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.ingest_coo(3, 3, [1, 2, 1], [1, 1, 2], [9.0, 4.0, 5.0], ONE | SYM)
    assert e.value.message == "Matrix is not symmetric"
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.ingest_coo(3, 3, [1, 4, 0], [1, 1, 2], [1.0, 2.0, 3.0], ONE)
    assert "entry 2 has an index outside the matrix" in e.value.message
    csr = ctx.ingest_coo(5, 5, [], [], [], ONE)
    assert csr.export()[0].tolist() == [0] * 6
    csr = ctx.ingest_coo(6, 4, [3, 3, 5], [4, 1, 2], [1.0, 2.0, 3.0], ONE)
    assert csr.export()[0].tolist() == [0, 0, 0, 2, 2, 3, 3]
    csr = ctx.ingest_coo(4, 4, [1, 2], [2, 3], [1.0, 1.0], ONE | DROP)
    assert csr.nnz == 0


def test_random_matrices_against_the_model(gpu_lib, ctx):
    ONE, SYM, DROP = gpu_lib.INGEST_ONE_BASED, gpu_lib.INGEST_SYMMETRIC, gpu_lib.INGEST_DROP_UPPER
    rng = np.random.default_rng(3)
    for trial in range(40):
        n, m = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        L = int(rng.integers(0, 5000))
        sym = trial % 3 == 0
        if sym:
            m = n
        rows, cols = rng.integers(1, n + 1, L), rng.integers(1, m + 1, L)
        vals = ((np.minimum(rows, cols) * 31 + np.maximum(rows, cols)) % 7).astype(float) if sym else rng.standard_normal(L)
        flags = ONE | (SYM if sym else 0) | (DROP if trial % 5 == 0 else 0)
        rp, ci, va, nnzs = dok_ingest(n, m, rows, cols, vals, sym, True, bool(flags & DROP))
        csr = ctx.ingest_coo(n, m, rows, cols, vals, flags)
        same(csr, rp, ci, va)
        assert csr.nnzs_field == nnzs


def test_ingested_matrix_feeds_preprocess_without_visiting_the_host(gpu_lib, ctx, oracle):
    """file entries -> device CSR -> preprocess -> spmv; y equals the oracle's CsrMatrix::dot bit for bit."""
    n, rp, ci, va = oracle.gen_poisson2d(96)
    rows, cols, vals = coo_of(n, rp, ci, va)
    perm = np.random.default_rng(7).permutation(len(vals))
    csr = ctx.ingest_coo(n, n, rows[perm], cols[perm], vals[perm], gpu_lib.INGEST_ONE_BASED)
    ctx.preprocess_csr(gpu_lib.design(num_pipes=2, cache_size=8192, input_width=16), csr)
    x = np.random.default_rng(0).random(n)
    assert np.array_equal(ctx.spmv(x), oracle.csr_dot(n, rp, ci, va, x))
    # symmetric expansion of the stored lower triangle gives the same operator
    low = ci <= np.repeat(np.arange(n), np.diff(rp))
    csr2 = ctx.ingest_coo(n, n, rows[low], cols[low], vals[low], gpu_lib.INGEST_ONE_BASED | gpu_lib.INGEST_SYMMETRIC)
    same(csr2, rp, ci, va)


def test_large_ingest_is_sorted_and_complete(gpu_lib, ctx, oracle):
    """Size-independent properties at a size the model cannot check entry by entry: 4M entries with 25% duplicates."""
    rng = np.random.default_rng(11)
    n = 1 << 20
    L = 1 << 22
    rows = rng.integers(1, n + 1, L).astype(np.int32)
    cols = rng.integers(1, n + 1, L).astype(np.int32)
    rows[: L // 4] = rows[L // 4: L // 2]
    cols[: L // 4] = cols[L // 4: L // 2]
    vals = rng.standard_normal(L)
    csr = ctx.ingest_coo(n, n, rows, cols, vals, gpu_lib.INGEST_ONE_BASED)
    rp, ci, va = csr.export()
    key = (np.repeat(np.arange(n, dtype=np.int64), np.diff(rp)) << 32) | ci
    assert rp[0] == 0 and rp[-1] == csr.nnz and np.all(np.diff(rp) >= 0)
    assert np.all(np.diff(key) > 0)                                   # strictly ascending (row, column): sorted, no duplicates
    want = np.unique(((rows.astype(np.int64) - 1) << 32) | (cols - 1))
    assert np.array_equal(key, want)
    # last value wins: compare with a stable sort on the host
    order = np.argsort(((rows.astype(np.int64) - 1) << 32) | (cols - 1), kind="stable")
    k_sorted = (((rows.astype(np.int64) - 1) << 32) | (cols - 1))[order]
    last = np.r_[k_sorted[1:] != k_sorted[:-1], True]
    assert np.array_equal(va, vals[order][last])
