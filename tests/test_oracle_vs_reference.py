"""Live comparison of the C restatement with the compiled reference (oracle/_ref/libcaskref.so).
Skipped where that library is absent; the golden-fixture tests in test_oracle.py cover that case."""
import numpy as np
import pytest

from oracle import refbind as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")


def _same(parts_a, parts_b):
    assert len(parts_a) == len(parts_b)
    for (sa, ca, pa), (sb, cb, pb) in zip(parts_a, parts_b):
        assert sa == sb
        assert np.array_equal(ca, cb)
        assert pa.tobytes() == pb.tobytes()


@pytest.mark.parametrize("gen,arg", [("gen_poisson2d", 40), ("gen_poisson3d27", 9), ("gen_convdiff3d7", 10), ("gen_rmat", 9)])
def test_preprocess_on_synthetic_twins(oracle, gen, arg):
    n, rp, ci, va = getattr(oracle, gen)(arg)
    m = R.RefMatrix.from_csr(n, n, rp, ci, va)
    for arch in (0, 1):
        for pipes, cache, width in ((1, 256, 16), (3, 50, 4), (n + 5, 64, 2)):
            _same(oracle.preprocess(n, n, rp, ci, va, arch, pipes, cache, width), m.preprocess(arch, pipes, cache, width))
    x = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(oracle.csr_dot(n, rp, ci, va, x), m.dot(x))


def test_random_matrices_including_unsorted_rows(oracle):
    rng = np.random.default_rng(11)
    for trial in range(20):
        n, mcols = int(rng.integers(1, 60)), int(rng.integers(1, 70))
        rows = [rng.permutation(mcols)[: rng.integers(0, min(mcols, 9) + 1)] for _ in range(n)]
        if trial % 2 == 0:
            rows = [np.sort(r) for r in rows]
        rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
        ci = (np.concatenate(rows) if rp[-1] else np.zeros(0)).astype(np.int32)
        va = rng.standard_normal(len(ci))
        m = R.RefMatrix.from_csr(n, mcols, rp, ci, va)
        for arch in (0, 1):
            pipes, cache, width = int(rng.integers(1, 9)), int(rng.integers(1, 40)), int(rng.integers(1, 9))
            _same(oracle.preprocess(n, mcols, rp, ci, va, arch, pipes, cache, width), m.preprocess(arch, pipes, cache, width))
        x = rng.standard_normal(max(mcols, n))
        assert np.array_equal(oracle.csr_dot(n, rp, ci, va, x), m.dot(x)[:n])


def test_slices(oracle, golden):
    """sliceRows / sliceColumns unit pins (test/SparseMatrix.cpp:139-161)."""
    n, m, rp, ci, va = golden.csr("test_cage6")
    ref = R.RefMatrix.from_csr(n, m, rp, ci, va)
    srp, sci, sva = ref.slice_rows(10, 20)
    assert np.array_equal(srp, rp[10:31] - rp[10]) and np.array_equal(sci, ci[rp[10]:rp[30]])
    (sc, colptr, pairs), = oracle.preprocess(n, m, rp, ci, va, 0, 1, 16, 1)
    nb = sc["nBlocks"]
    off = 0
    for b in range(nb):
        brp, bci, bva = ref.slice_columns(16, b)
        assert np.array_equal(colptr[b * n:(b + 1) * n], brp)
        assert np.array_equal(pairs["indptr"][off:off + len(bci)], bci)
        assert np.array_equal(pairs["value"][off:off + len(bva)], bva)
        off += len(bci)


def test_mock_flow_returns_zeros_and_checks_arguments(golden):
    """with the mock callbacks Spmv::spmv computes nothing; its checks still fire."""
    n, m, rp, ci, va = golden.csr("test_small")
    ref = R.RefMatrix.from_csr(n, m, rp, ci, va)
    ref.preprocess(0, 1, 8, 4)
    rc, msg, y = ref.spmv_mock(golden.x("test_small"))
    assert rc == 0 and not y.any()
    ref.preprocess(0, 1, 8, 4, max_rows=8)
    rc, msg, _ = ref.spmv_mock(golden.x("test_small"))
    assert rc == 1 and msg == "Matrix is too large! Maximum supported rows: 8 actual rows: 16"
    ref.preprocess(0, 3, 8, 4, num_controllers=2)
    rc, msg, _ = ref.spmv_mock(golden.x("test_small"))
    assert rc == 2 and msg == "numPipes should be a multiple of numControllers"
