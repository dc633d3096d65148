"""Preconditioned CG on the GPU (cask_b200_pcg / cask_b200_ilu_factor / cask_b200_ilu_apply) through the C ABI against
the oracle's restatement of pcg<double, Precon> and ILUPreconditioner (pinned to the reference's known answers and to
its compiled code in tests/test_oracle_precond.py).  Bars: ILU factors and solves bit-identical (same arithmetic, level
order only changes which independent rows run together); CG iteration counts within +-1, same stopping rule.
File name sorts after the other GPU suites on purpose: this path was written after the last GPU session of round 1."""
import numpy as np
import pytest

from test_oracle_precond import ARROW, dense_to_csr, ulp_close

pytestmark = pytest.mark.gpu


def prep(gpu_lib, ctx, n, rp, ci, va):
    ctx.preprocess(gpu_lib.design(num_pipes=1, cache_size=8192, input_width=16), n, n, rp, ci, va)


@pytest.mark.parametrize("gen,arg", [("gen_poisson2d", 48), ("gen_poisson3d27", 12), ("gen_convdiff3d7", 14)])
def test_ilu_factors_and_solves_bit_exact(gpu_lib, ctx, oracle, gen, arg):
    n, rp, ci, va = getattr(oracle, gen)(arg)
    prep(gpu_lib, ctx, n, rp, ci, va)
    pc, ll, lu = ctx.ilu_factor(len(va))
    exp = oracle.ilu0(n, rp, ci, va)
    assert np.array_equal(pc, exp)
    assert ll == lu and ll > 1
    x = np.random.default_rng(0).standard_normal(n)
    for unit in (False, True):
        z, zp = ctx.ilu_apply(x, unit)
        ez, bad = oracle.ilu_apply(n, rp, ci, exp, x, unit)
        assert not zp and not bad and np.array_equal(z, ez)


def test_reference_known_answers(gpu_lib, ctx):
    """ILUCompute2 and ILUComputeAndApply, test/LinearSolvers.cpp:79-99, 125-146."""
    n, rp, ci, va = dense_to_csr(ARROW)
    prep(gpu_lib, ctx, n, rp, ci, va)
    pc, ll, lu = ctx.ilu_factor(len(va))
    assert pc.tolist() == [2, 1, 1, 1, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]
    z, zp = ctx.ilu_apply([1.0, 2.0, 3.0, 4.0])
    assert not zp and z.tolist() == [-16.25, 7, 11, 15]


@pytest.mark.parametrize("gen,arg", [("gen_poisson2d", 40), ("gen_poisson3d27", 10)])
def test_pcg_iteration_counts_match_the_oracle(gpu_lib, ctx, oracle, gen, arg):
    n, rp, ci, va = getattr(oracle, gen)(arg)
    prep(gpu_lib, ctx, n, rp, ci, va)
    xt = 1.0 + 0.25 * (np.arange(n) % 4)
    b = oracle.csr_dot(n, rp, ci, va, xt)
    for name, code in (("identity", gpu_lib.PRECON_IDENTITY), ("jacobi", gpu_lib.PRECON_JACOBI), ("ilu_unit", gpu_lib.PRECON_ILU_UNIT)):
        oc, oit, ox, ors = oracle.pcg_precond(n, rp, ci, va, b, name, lower=False)
        conv, it, x, rs = ctx.pcg(b, code)
        assert conv and oc
        assert abs(it - oit) <= 1, (name, it, oit)
        assert rs <= 1e-10 and np.abs(x - xt).max() < 1e-4
    # identity through this loop agrees with the tuned loop
    c0, it0, x0, rs0 = ctx.cg(b)
    c1, it1, x1, rs1 = ctx.pcg(b, gpu_lib.PRECON_IDENTITY)
    assert c0 and c1 and abs(it0 - it1) <= 1


def test_reference_ilu_pairing_on_tinysym(gpu_lib, ctx, golden):
    """CGSymWithILUPC, test/LinearSolvers.cpp:54-77: the product uses the symmetric matrix, the ILU is built from the
    stored LOWER TRIANGLE (what pcg hands Precon{a}); the loop never converges and x ends at the asserted values.
    The iteration is a stagnating fixed point, so the different summation order of the GPU dots is expected to land on
    the same doubles to ~1e-12; 4 ulp (ASSERT_DOUBLE_EQ) is what the sequential oracle reaches."""
    s = golden.systems["tinysym"]
    n, m, rp, ci, va = golden.csr("tinysym")           # full symmetric matrix (io::readMatrix)
    prep(gpu_lib, ctx, n, rp, ci, va)
    rows = np.repeat(np.arange(n), np.diff(s["row_ptr"]))
    low = ctx.ingest_coo(n, n, rows, s["col_ind"], s["values"], 0)
    ctx.precond_set_matrix(low)
    conv, it, x, rs = ctx.pcg(s["rhs"], gpu_lib.PRECON_ILU)
    exp = [-1.9982580059252246, 2.0000862488691915, 3.0001293733037859, 2.9987581910958183]
    assert (conv, it) == (False, 1999)
    assert np.allclose(x, exp, rtol=1e-9, atol=0)
    print("tinysym ILU-PCG: max ulp distance to the reference's asserted x =",
          float(np.max(np.abs(x - exp) / np.spacing(np.abs(exp)))), "4-ulp equal:", bool(ulp_close(x, exp)))


def test_zero_pivot_is_reported(gpu_lib, ctx):
    n, rp, ci, va = dense_to_csr([[0, 1, 0], [1, 4, 1], [0, 1, 4]])
    prep(gpu_lib, ctx, n, rp, ci, va)
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.pcg([1.0, 2.0, 3.0], gpu_lib.PRECON_ILU, maxiters=3)
    assert e.value.code == gpu_lib.ERR_RUNTIME and "pivot" in e.value.message


def test_unsorted_rows_are_refused(gpu_lib, ctx):
    ctx.preprocess(gpu_lib.design(), 2, 2, [0, 2, 3], [1, 0, 1], [1.0, 4.0, 3.0])
    with pytest.raises(gpu_lib.CaskError) as e:
        ctx.ilu_factor(3)
    assert e.value.code == gpu_lib.ERR_UNSUPPORTED


def test_ilu_level_kernel_and_graph_replay_the_same_solves(gpu_lib, ctx, oracle):
    """An ILU application runs as ONE cooperative kernel that walks all levels of both solves behind grid barriers (default),
    as a CUDA graph of the per-level kernels (ilu_persistent = 0), or kernel by kernel (ilu_graph = 0 as well): the same
    iterates bit for bit, across solves, preconditioner kinds and a change of matrix."""
    res = {}
    for mode, (persistent, graph) in enumerate(((1, 1), (0, 1), (0, 0))):
        ctx.set_option("ilu_persistent", persistent)
        ctx.set_option("ilu_graph", graph)
        out = []
        for gen, arg in (("gen_poisson2d", 40), ("gen_poisson3d27", 10), ("gen_convdiff3d7", 12)):
            n, rp, ci, va = getattr(oracle, gen)(arg)
            prep(gpu_lib, ctx, n, rp, ci, va)
            b = oracle.csr_dot(n, rp, ci, va, 1.0 + 0.25 * (np.arange(n) % 4))
            for code in (gpu_lib.PRECON_ILU_UNIT, gpu_lib.PRECON_ILU_UNIT, gpu_lib.PRECON_ILU):
                conv, it, x, rs = ctx.pcg(b, code, maxiters=60)
                out.append((conv, it, rs, x))
        res[mode] = out
    for other in (1, 2):
        for a, b in zip(res[0], res[other]):
            assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
    assert res[0][0][0] and res[0][3][0]      # the unit-lower solves converge on the SPD systems
    ctx.set_option("ilu_persistent", 1)
    ctx.set_option("ilu_graph", 1)
