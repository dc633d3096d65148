"""GPU partitioner vs the compiled reference: Partition scalars and raw array bytes must be
identical (Spmv::preprocess / do_blocking, src/runtime/Spmv.cpp:42-107,329-365)."""
import numpy as np
import pytest

from conftest import sha256

pytestmark = pytest.mark.gpu

SCALARS = ("nBlocks", "n", "paddingCycles", "totalCycles", "vector_load_cycles", "outSize", "reductionCycles",
           "emptyCycles", "m_colptr_unpaddedLength", "m_indptr_values_unpaddedLength", "len_colptr", "len_pairs")


def test_partitions_match_reference_bit_for_bit(golden, gpu_lib, ctx):
    checked = 0
    for name in golden.names:
        n, m, rp, ci, va = golden.csr(name)
        for case in golden.partitions[name]:
            d = gpu_lib.design(case["num_pipes"], case["cache_size"], case["input_width"], arch=case["arch"])
            ctx.preprocess(d, n, m, rp, ci, va)
            for p, g in enumerate(case["partitions"]):
                info, colptr, pairs = ctx.partition(p)
                for k in SCALARS:
                    assert info[k] == g[k], (name, case["arch"], case["num_pipes"], case["cache_size"],
                                             case["input_width"], p, k, info[k], g[k])
                assert sha256(colptr) == g["colptr_sha256"], (name, case, p, "colptr")
                assert sha256(pairs) == g["pairs_sha256"], (name, case, p, "pairs")
                if "colptr" in g:
                    assert colptr.tolist() == g["colptr"]
                    assert pairs["indptr"].tolist() == g["pairs_idx"]
                    assert pairs["value"].tolist() == g["pairs_val"]
                checked += 1
    assert checked == sum(len(c["partitions"]) for cs in golden.partitions.values() for c in cs)


def test_partitions_match_oracle_on_synthetic_twins(oracle, gpu_lib, ctx):
    """Down-scaled twins of the BASELINE configs: 2D Poisson 64^2, 3D 27-pt 12^3,
    conv-diff 16^3, R-MAT scale 10 — compared with the C restatement (itself pinned to the reference)."""
    twins = [oracle.gen_poisson2d(64), oracle.gen_poisson3d27(12), oracle.gen_convdiff3d7(16), oracle.gen_rmat(10, 8, 1)]
    for n, rp, ci, va in twins:
        for arch in (0, 1):
            for pipes, cache, width in ((1, 512, 16), (4, 256, 4), (7, 100, 3)):
                exp = oracle.preprocess(n, n, rp, ci, va, arch, pipes, cache, width)
                ctx.preprocess(gpu_lib.design(pipes, cache, width, arch=arch), n, n, rp, ci, va)
                for p, (sc, colptr, pairs) in enumerate(exp):
                    info, gc, gp = ctx.partition(p)
                    assert info == sc
                    assert np.array_equal(gc, colptr)
                    assert gp.tobytes() == pairs.tobytes()


def test_unsorted_rows_keep_reference_order(oracle, gpu_lib, ctx):
    """sliceColumns keeps the row's original entry order inside a block (SparseMatrix.hpp:473-479)."""
    rng = np.random.default_rng(7)
    n = m = 50
    rows = []
    for i in range(n):
        k = rng.integers(0, 12)
        rows.append(rng.permutation(m)[:k])
    rp = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    ci = np.concatenate(rows).astype(np.int32) if rp[-1] else np.zeros(0, np.int32)
    va = rng.standard_normal(len(ci))
    for arch in (0, 1):
        exp = oracle.preprocess(n, m, rp, ci, va, arch, 3, 7, 4)
        ctx.preprocess(gpu_lib.design(3, 7, 4, arch=arch), n, m, rp, ci, va)
        for p, (sc, colptr, pairs) in enumerate(exp):
            info, gc, gp = ctx.partition(p)
            assert info == sc and np.array_equal(gc, colptr) and gp.tobytes() == pairs.tobytes()


def test_export_refuses_reference_overflow(gpu_lib, ctx, oracle):
    """rows x blocks beyond INT32_MAX: the reference's own counters overflow (Spmv.cpp:68)."""
    n, rp, ci, va = oracle.gen_poisson2d(512)  # 262144 rows
    ctx.preprocess(gpu_lib.design(1, 16, 4), n, n, rp, ci, va)  # 16384 blocks -> 4.3e9 cells
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.partition(0)
    assert e.value.code == gpu_lib.ERR_UNSUPPORTED
