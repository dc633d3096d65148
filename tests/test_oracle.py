"""CPU suite (no GPU): pins the C restatement (oracle/) to the reference.

Two anchors: (1) the committed golden fixtures generated from the compiled reference
(tests/golden/make_golden.py), always; (2) the compiled reference itself (oracle/_ref), live, where
it is present (it is built from /root/reference in the build container and travels as a prebuilt .so).
"""
import json

import numpy as np
import pytest

from conftest import sha256

SCALARS = ("nBlocks", "n", "paddingCycles", "totalCycles", "vector_load_cycles", "outSize", "reductionCycles",
           "emptyCycles", "m_colptr_unpaddedLength", "m_indptr_values_unpaddedLength", "len_colptr", "len_pairs")


def test_pair_record_is_12_bytes(oracle):
    assert oracle.PAIR_DTYPE.itemsize == 12  # #pragma pack(1) struct indptr_value, Spmv.hpp:13-20


def test_survey_golden_vectors(golden, oracle):
    """vectors captured from the shim-compiled reference during the survey."""
    E = lambda k: int(np.int32(np.uint32(k | 0x80000000)))
    n, m, rp, ci, va = golden.csr("test_break")
    p0, p1 = oracle.preprocess(n, m, rp, ci, va, 0, 2, 2, 2)
    assert p0[1].tolist() == [1, 2, 3, 4] + [0] * 12 and p0[2]["indptr"].tolist() == [0, 0, 0, 0]
    assert (p0[0]["totalCycles"], p0[0]["reductionCycles"], p0[0]["paddingCycles"]) == (24, 16, 44)
    assert p1[1].tolist() == [1, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 2, 0, 0, 1, 1]
    assert p1[2]["indptr"].tolist() == [1, 0, 0, 1, 0, 0]
    s0, s1 = oracle.preprocess(n, m, rp, ci, va, 1, 2, 2, 2)
    assert s0[1].tolist() == [1, 2, 3, 4, E(4), E(4), 0, 0, 0, 0]
    assert (s0[0]["totalCycles"], s0[0]["reductionCycles"], s0[0]["emptyCycles"]) == (18, 10, 6)
    assert s1[1].tolist() == [1, 1, 1, 1, E(4), E(1), 1, E(1), 2, 0, 0, 1, 1]
    assert (s1[0]["totalCycles"], s1[0]["reductionCycles"], s1[0]["emptyCycles"]) == (21, 13, 3)
    n, m, rp, ci, va = golden.csr("test_tiny_odd")
    (a,) = oracle.preprocess(n, m, rp, ci, va, 0, 1, 4, 2)
    (b,) = oracle.preprocess(n, m, rp, ci, va, 1, 1, 4, 2)
    assert (a[0]["nBlocks"], len(a[1]), len(a[2]), a[0]["totalCycles"]) == (5, 85, 18, 105)
    assert (len(b[1]), b[0]["totalCycles"], b[0]["emptyCycles"]) == (52, 72, 33)
    n, m, rp, ci, va = golden.csr("test_small")
    (c,) = oracle.preprocess(n, m, rp, ci, va, 0, 1, 8, 4)
    assert (c[0]["nBlocks"], len(c[1]), len(c[2]), c[0]["totalCycles"], c[0]["reductionCycles"],
            c[0]["paddingCycles"], c[0]["outSize"]) == (2, 32, 28, 49, 32, 32, 384)


def test_partitions_match_golden(golden, oracle):
    checked = 0
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        for case in golden.partitions[name]:
            parts = oracle.preprocess(n, m, rp, ci, va, case["arch"], case["num_pipes"], case["cache_size"],
                                      case["input_width"])
            assert len(parts) == len(case["partitions"])
            for (sc, colptr, pairs), g in zip(parts, case["partitions"]):
                assert all(sc[k] == g[k] for k in SCALARS), (name, case["arch"], case["num_pipes"])
                assert sha256(colptr) == g["colptr_sha256"] and sha256(pairs) == g["pairs_sha256"]
                checked += 1
    assert checked > 4000


def test_dot_matches_golden_bit_for_bit(golden, oracle):
    """CsrMatrix::dot of the reference on x[i] = 0.25 i (test_spmv.cpp:27-28)."""
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        assert np.array_equal(oracle.csr_dot(n, rp, ci, va, golden.x(name)), golden.dots[name]), name


def test_reference_unit_test_vectors(oracle):
    """test/SparseMatrix.cpp:63-73,164-191: DokDotProduct, CsrDotProduct, SymCsrDotProduct."""
    def csr(dense):
        d = np.array(dense, float)
        rp = np.concatenate([[0], np.cumsum((d != 0).sum(1))])
        return len(d), rp, np.nonzero(d)[1], d[d != 0]
    n, rp, ci, va = csr([[1, 0, 0, 0], [1, 1, 0, 0], [1, 0, 1, 0], [1, 0, 0, 1]])
    assert oracle.csr_dot(n, rp, ci, va, [1, 2, 3, 4]).tolist() == [1, 3, 4, 5]
    n, rp, ci, va = csr([[1, 0, 0, 0], [1, 0, 1, 0], [0, 1, 1, 0], [0, 0, 1, 1]])
    assert oracle.csr_dot(n, rp, ci, va, [1, 2, 3, 4]).tolist() == [1, 4, 5, 7]
    n, rp, ci, va = csr([[1, 1, 1, 1], [1, 1, 0, 0], [1, 0, 1, 1], [1, 0, 1, 1]])
    assert oracle.csr_dot(n, rp, ci, va, [1, 2, 3, 4]).tolist() == [10, 3, 8, 8]


def test_partition_format_spmv_agrees_with_dot(golden, oracle):
    """Consuming the partition arrays the way the dataflow engine does reproduces A x."""
    from conftest import assert_y_close, row_scale
    for name in ("test_small", "test_break", "test_cage6", "bfwb62", "test_tols90", "test_dense_32", "test_two_rows_2",
                 "test_some_empty_rows", "test_one_row", "tinysym"):
        n, m, rp, ci, va = golden.csr(name)
        x = golden.x(name)
        for arch in (0, 1):
            for pipes, cache, width in ((1, 2048, 16), (2, 8, 4), (5, 64, 3), (48, 32, 2)):
                parts = oracle.preprocess(n, m, rp, ci, va, arch, pipes, cache, width)
                y = oracle.partition_spmv(parts, cache, width, x, n)
                assert_y_close(y, golden.dots[name], row_scale(n, rp, ci, va, x))


def test_pcg_reference_known_answers(golden, oracle):
    """test/LinearSolvers.cpp:14-52 (ASSERT_DOUBLE_EQ = 4 ulp): tiny -> {1,2,3,4}, tinysym -> {-2,2,3,3},
    fed the lower triangle exactly as readSymMatrix produces it."""
    for name, s in golden.systems.items():
        conv, iters, x = oracle.pcg(s["n"], s["row_ptr"], s["col_ind"], s["values"], s["rhs"])
        exp = np.array(s["asserted_solution"], float)
        assert conv
        assert (np.abs(x - exp) / np.spacing(np.abs(exp))).max() <= 4, (name, x)
        assert np.allclose(s["sol_file"], exp)


def test_pcg_iteration_convention(oracle):
    """iterations is assigned only at the end of a non-converged iteration (SparseLinearSolvers.hpp:231)."""
    n, rp, ci, va = oracle.gen_poisson2d(8)
    b = oracle.csr_dot(n, rp, ci, va, np.ones(n))
    conv, it, x, rs = oracle.pcg(n, rp, ci, va, b, lower=False)
    conv2, it2, x2, _ = oracle.pcg(n, rp, ci, va, b, maxiters=it + 1, lower=False)
    assert conv and not conv2 and it2 == it  # one iteration short: not converged, same reported index
    conv3, it3, _, _ = oracle.pcg(n, rp, ci, va, b, maxiters=1, lower=False)
    assert not conv3 and it3 == 0


def test_bicgstab_against_direct_solve(oracle):
    """PARITY UNPINNED (Eigen 3.3.1 absent, no reference test): sanity against scipy's direct solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n, rp, ci, va = oracle.gen_convdiff3d7(8)
    a = sp.csr_matrix((va, ci, rp), shape=(n, n))
    b = a @ np.ones(n)
    x, it, err = oracle.bicgstab(n, rp, ci, va, b, tol=1e-12)
    assert err <= 1e-12 and it < 2 * n
    assert np.abs(x - spla.spsolve(a.tocsc(), b)).max() < 1e-9
    x0, it0, err0 = oracle.bicgstab(n, rp, ci, va, np.zeros(n))
    assert it0 == 0 and not x0.any()


def test_generators(oracle):
    import scipy.sparse as sp
    n, rp, ci, va = oracle.gen_poisson2d(4096 // 64)
    assert len(va) == 5 * n - 4 * 64
    a = sp.csr_matrix((va, ci, rp), shape=(n, n))
    assert (abs(a - a.T)).nnz == 0 and a.has_sorted_indices
    n, rp, ci, va = oracle.gen_poisson3d27(10)
    assert len(va) == (3 * 10 - 2) ** 3  # 766^3 at N = 256
    a = sp.csr_matrix((va, ci, rp), shape=(n, n))
    assert (abs(a - a.T)).nnz == 0 and np.all(a.diagonal() == 26)
    n, rp, ci, va = oracle.gen_convdiff3d7(9)
    assert len(va) == 7 * n - 6 * 81
    a = sp.csr_matrix((va, ci, rp), shape=(n, n))
    assert (abs(a - a.T)).nnz > 0 and np.all(a.diagonal() > -(a - sp.diags(a.diagonal())).sum(1).A1 - 1e-12)
    n, rp, ci, va = oracle.gen_rmat(10, 8, 1)
    a = sp.csr_matrix((va, ci, rp), shape=(n, n))
    assert a.has_sorted_indices and len(va) <= 8 * n and np.diff(rp).max() > 20 * np.diff(rp).mean() / 4


def test_openmp_port_equals_dot_order(oracle):
    n, rp, ci, va = oracle.gen_poisson3d27(12)
    x = np.random.default_rng(1).random(n)
    y, threads = oracle.csr_spmv_omp(n, rp, ci, va, x)
    assert threads >= 1 and np.array_equal(y, oracle.csr_dot(n, rp, ci, va, x))
