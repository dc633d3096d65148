"""The C++ side on the GPU: the reference's harnesses rebuilt on the host mirror (cask_b200/host), and the
reference's OWN unmodified test_spmv.cpp + Spmv.cpp driving the GPU through the B200 plugin
(oracle/_ref/test_spmv_reference_on_b200, prebuilt in the container that has /root/reference)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

HOST_BIN = os.path.join(ROOT, "cask_b200", "host", "bin")
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "test_spmv_reference_on_b200")
MATRICES = ["test_small", "test_break", "test_tiny", "test_tiny_odd", "test_dense_48", "test_cage6", "bfwb62",
            "test_tols90", "test_long_row", "test_one_row", "test_two_rows_2", "test_some_empty_rows",
            "test_empty_last_rows_small", "test_partition", "test_non_multiple", "test_large_empty", "test_tsopf1",
            "test_wa", "OPF_3754", "OPF_6000", "TSOPF_RS_b39_c7", "dw8192", "t2d_q9_A_01"]


def write_mtx(path, n, m, rp, ci, va):
    rows = np.repeat(np.arange(n), np.diff(rp))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%\n")
        f.write("%d %d %d\n" % (n, m, len(va)))
        f.write("".join("%d %d %s\n" % (r + 1, c + 1, repr(float(v))) for r, c, v in zip(rows, ci, va)))


@pytest.fixture(scope="module")
def mtx_dir(tmp_path_factory, golden):
    d = tmp_path_factory.mktemp("mtx")
    for name in MATRICES:
        n, m, rp, ci, va = golden.csr(name)
        write_mtx(str(d / (name + ".mtx")), n, m, rp, ci, va)
    return d


def _run(cmd):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    return p.returncode, p.stdout


def test_host_mirror_test_spmv(mtx_dir):
    """test/test_spmv.cpp's flow on cask::spmv::Spmv (host mirror): exit 0 and 'All tests passed!'."""
    exe = os.path.join(HOST_BIN, "test_spmv")
    assert os.path.exists(exe), "run `make -C cask_b200/host` (__graft_entry__.build())"
    for name in MATRICES:
        rc, out = _run([exe, str(mtx_dir / (name + ".mtx"))])
        assert rc == 0 and "Test passed!" in out and "All tests passed!" in out, (name, out[-2000:])
    for impl in ("0", "1", "2"):
        rc, out = _run([exe, str(mtx_dir / "test_cage6.mtx"), impl])
        assert rc == 0, out[-2000:]
        # the complete "Result  <key>=" set of Spmv.cpp:266-301, in the reference's order, one cycle figure per pipe
        keys = [l.split("=")[0][len("Result  "):] for l in out.splitlines() if l.startswith("Result  ")]
        assert keys == ["Total cycles", "Padding cycles", "Reduction cycles", "Input width ", "Pipes ", "Iterations",
                        "Took (ms)", "Est (ms)", "Gflops (est)", "Gflops (actual)", "BWidth (est)"], keys
        pipes = int([l for l in out.splitlines() if l.startswith("Result  Pipes ")][0].split("=")[1].rstrip(","))
        total = [l for l in out.splitlines() if l.startswith("Result  Total cycles")][0].split("=")[1]
        assert len([v for v in total.split(",") if v]) == pipes and all(int(v) > 0 for v in total.split(",") if v)
    rc, out = _run([exe, str(mtx_dir / "test_cage6.mtx"), "7"])  # vector::at -> std::out_of_range like the reference
    assert rc != 0


def test_host_mirror_solver_and_client_suites(tmp_path, golden):
    """LinearSolvers.cpp / CgTest.cpp / ClientTestSpmv.cpp / ClientTestCg.cpp on the GPU path."""
    exe = os.path.join(HOST_BIN, "test_client")
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
    rc, out = _run([exe, str(tmp_path)])
    assert rc == 0 and "PASSED (0 failures)" in out, out[-4000:]
    log = open(tmp_path / "sol.upc.tinysym.log").read()
    assert log.startswith('{"setup took":"') and '"iterations":"' in log  # Benchmark.hpp:54-71 record


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref harness not built (needs /root/reference at build time)")
def test_unmodified_reference_harness_on_the_gpu(mtx_dir):
    """The reference's own test_spmv.cpp and Spmv::spmv, unmodified, against libSpmv_b200's run/write/read
    callbacks: the check that fails on the reference's mock flow (zeros) passes on the B200."""
    for name in MATRICES:
        rc, out = _run([REF_HARNESS, str(mtx_dir / (name + ".mtx"))])
        assert rc == 0 and "All tests passed!" in out, (name, out[-3000:])
    for impl in ("0", "1", "2"):
        rc, out = _run([REF_HARNESS, str(mtx_dir / "bfwb62.mtx"), impl])
        assert rc == 0 and "Running on DFE" in out, out[-3000:]
