"""Host-side tools on the GPU: bin/cask_dse (the architecture selector) and bin/test_client's
extra group (the host mirror's ILUPreconditioner / pcg<T, Precon> / io::gpu readers).  They drive the same C ABI as the
suites before them through C++ programs; the file name sorts last so that a failure here cannot keep pytest -x from
reaching the ABI-level suites of the ingest, preconditioner and coded-format paths."""
import numpy as np
import pytest

from test_mmio_host import coo_of, write_mtx

pytestmark = pytest.mark.gpu


def test_architecture_selector_tool(tmp_path, golden, oracle):
    """bin/cask_dse (src/main.cpp + Dse.cpp of the reference, B200 edition): candidates preprocessed on the GPU, the
    reference's table on stdout, dse_out.json with the reference's keys; the stencil is scored all-staged-ELL."""
    import json
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "cask_b200", "host", "bin", "cask_dse")
    assert os.path.exists(exe), "run `make -C cask_b200/host` (__graft_entry__.build())"
    n, rp, ci, va = oracle.gen_poisson2d(128)
    p1, p2, out = str(tmp_path / "poisson.mtx"), str(tmp_path / "cage.mtx"), str(tmp_path / "dse_out.json")
    write_mtx(p1, n, n, *coo_of(n, rp, ci, va))
    n2, m2, rp2, ci2, va2 = golden.csr("test_cage6")
    write_mtx(p2, n2, m2, *coo_of(n2, rp2, ci2, va2))
    r = subprocess.run([exe, "--out", out, p1, p2], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "File Architecture CacheSize InputWidth NumPipes" in r.stdout and " Best " in r.stdout
    assert r.stdout.count("/poisson.mtx SkipEmpty ") >= 4  # one line per cache size + the best line
    d = json.load(open(out))
    best = d["best_architectures"]
    assert 1 <= len(best) <= 2 and sum(len(b["matrices"]) for b in best) == 2
    for b in best:
        assert b["name"] == "SkipEmpty" and float(b["estimated_gflops"]) > 0
        assert set(b["architecture_params"]) == {"num_pipes", "cache_size", "input_width", "max_rows", "num_controllers"}
    pb = [b for b in best if p1 in b["matrices"]][0]
    assert int(pb["estimated_impl_params"]["slices_gather_csr"]) == 0 and float(pb["estimated_impl_params"]["ell_fill"]) > 0.95


def test_host_mirror_reference_ilu_suites(tmp_path, golden):
    """test/LinearSolvers.cpp:54-146 (CGSymWithILUPC, ILUCompute2, ILUCompute, ILUComputeAndApply) on the host mirror's
    ILUPreconditioner / pcg<double, ILUPreconditioner>, plus the Jacobi / unit-ILU extensions and io::gpu readers."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "cask_b200", "host", "bin", "test_client")
    assert os.path.exists(exe)
    for name, s in golden.systems.items():
        n = s["n"]
        rp, ci, va = np.array(s["row_ptr"]), np.array(s["col_ind"]), np.array(s["values"])
        rows = np.repeat(np.arange(n), np.diff(rp))
        with open(tmp_path / (name + ".mtx"), "w") as f:
            f.write("%%%%MatrixMarket matrix coordinate real symmetric\n%%\n%d %d %d\n" % (n, n, len(va)))
            f.write("".join("%d %d %s\n" % (r + 1, c + 1, repr(float(v))) for r, c, v in zip(rows, ci, va)))
        for suffix, vec in (("_b", s["rhs"]), ("_sol", s["sol_file"])):
            with open(tmp_path / (name + suffix + ".mtx"), "w") as f:
                f.write("%%%%MatrixMarket matrix array real general\n%%\n%d 1\n" % n)
                f.write("".join("%s\n" % repr(float(v)) for v in vec))
    p = subprocess.run([exe, str(tmp_path), "extra"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0 and "PASSED (0 failures)" in p.stdout, p.stdout[-4000:]
