"""Device generators of the BASELINE matrices vs their CPU restatement, and SpMV on them."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_y_close, row_scale

pytestmark = pytest.mark.gpu


def _gen_on_device(gpu_lib, kind, N, row0=0, nrows=None):
    import torch
    n = gpu_lib.synth_rows(kind, N)
    nrows = n - row0 if nrows is None else nrows
    nnz = gpu_lib.synth_nnz(kind, N, row0, nrows)
    rp = torch.empty(nrows + 1, dtype=torch.int32, device="cuda")
    ci = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")
    va = torch.empty(max(nnz, 1), dtype=torch.float64, device="cuda")
    gpu_lib.synth_device(kind, N, row0, nrows, rp.data_ptr(), ci.data_ptr(), va.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return n, nnz, rp, ci, va


@pytest.mark.parametrize("kind,gen,N", [(0, "gen_poisson2d", 37), (1, "gen_poisson3d27", 9), (2, "gen_convdiff3d7", 11),
                                        (0, "gen_poisson2d", 300), (1, "gen_poisson3d27", 20)])
def test_generators_match_oracle(gpu_lib, oracle, kind, gen, N):
    n, rp, ci, va = getattr(oracle, gen)(N)
    gn, nnz, d_rp, d_ci, d_va = _gen_on_device(gpu_lib, kind, N)
    assert gn == n and nnz == len(va)
    assert np.array_equal(d_rp.cpu().numpy(), rp)
    assert np.array_equal(d_ci.cpu().numpy()[:nnz], ci)
    assert np.array_equal(d_va.cpu().numpy()[:nnz], va)
    # any row stripe is generated independently (sliceRows semantics: row_ptr rebased, columns global)
    for world in (3, 8):
        for rank in range(world):
            r0, nr = gpu_lib.shard_rows(n, world, rank)
            _, z, s_rp, s_ci, s_va = _gen_on_device(gpu_lib, kind, N, r0, nr)
            assert np.array_equal(s_rp.cpu().numpy(), rp[r0:r0 + nr + 1] - rp[r0])
            assert np.array_equal(s_ci.cpu().numpy()[:z], ci[rp[r0]:rp[r0 + nr]])
            assert np.array_equal(s_va.cpu().numpy()[:z], va[rp[r0]:rp[r0 + nr]])


@pytest.mark.parametrize("gen,N", [("gen_poisson2d", 256), ("gen_poisson3d27", 32), ("gen_convdiff3d7", 64)])
def test_spmv_on_twins_is_bit_identical(gpu_lib, ctx, oracle, gen, N):
    """Down-scaled twins C2' C4' C5' (SURVEY 8d): stencils go to the staged-ELL kernel, whose row sums
    use the reference's summation order, so y equals the oracle bit for bit."""
    n, rp, ci, va = getattr(oracle, gen)(N)
    ctx.preprocess(gpu_lib.design(8, 8192, 16), n, n, rp, ci, va)
    st = ctx.plan_stats()
    assert st["slices_gather_csr"] == 0, st
    rng = np.random.default_rng(42)
    for x in (np.arange(n) % 1024 * 0.25, rng.random(n)):
        assert np.array_equal(ctx.spmv(x), oracle.csr_dot(n, rp, ci, va, x))


def test_spmv_on_rmat_twin(gpu_lib, ctx, oracle):
    n, rp, ci, va = oracle.gen_rmat(14, 12, 1)
    ctx.preprocess(gpu_lib.design(4, 8192, 16), n, n, rp, ci, va)
    st = ctx.plan_stats()
    assert st["slices_gather_csr"] > 0
    x = np.random.default_rng(5).random(n)
    assert_y_close(ctx.spmv(x), oracle.csr_dot(n, rp, ci, va, x), row_scale(n, rp, ci, va, x))


def test_full_size_poisson2d_properties(gpu_lib, ctx):
    """BASELINE config C2 (4096^2 grid, 16.7M rows, 83.9M nnz) at full size: size-independent
    properties of the 5-point operator instead of an oracle run."""
    import torch
    N = 4096
    n, nnz, rp, ci, va = _gen_on_device(gpu_lib, 0, N)
    assert n == N * N and nnz == 5 * n - 4 * N
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)  # stream-ordered with the torch tensors below
    ctx.preprocess_device(gpu_lib.design(1, 8192, 16), n, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    st = ctx.plan_stats()
    assert st["slices_gather_csr"] == 0 and st["ell_nnz"] == nnz
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    # (1) A * ones = number of missing neighbours (exact in fp64)
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    ctx.spmv_device(x.data_ptr(), y.data_ptr()); ctx.synchronize()
    g = y.view(N, N)
    exp = torch.zeros(N, N, dtype=torch.float64, device="cuda")
    exp[0, :] += 1; exp[-1, :] += 1; exp[:, 0] += 1; exp[:, -1] += 1
    assert torch.equal(g, exp)
    # (2) x = 0.25 * (k mod 1024): interior rows vanish exactly, every value is a multiple of 0.25
    k = torch.arange(n, device="cuda")
    x = (k % 1024).double() * 0.25
    ctx.spmv_device(x.data_ptr(), y.data_ptr()); ctx.synchronize()
    xg = x.view(N, N)
    ref = 4 * xg
    ref[1:, :] -= xg[:-1, :]; ref[:-1, :] -= xg[1:, :]; ref[:, 1:] -= xg[:, :-1]; ref[:, :-1] -= xg[:, 1:]
    assert torch.equal(y.view(N, N), ref)
    # (3) symmetry: <Ax, z> == <x, Az> to rounding; linearity: A(ax + z) == a Ax + Az on exact inputs
    z = ((k * 7) % 513).double() * 0.5
    yz = torch.empty_like(y)
    ctx.spmv_device(z.data_ptr(), yz.data_ptr()); ctx.synchronize()
    assert abs(float(torch.dot(y, z) - torch.dot(x, yz))) <= 1e-12 * abs(float(torch.dot(y, z)))
    comb = 2.0 * x + z
    yc = torch.empty_like(y)
    ctx.spmv_device(comb.data_ptr(), yc.data_ptr()); ctx.synchronize()
    assert torch.equal(yc, 2.0 * y + yz)
