"""The architecture selector's traffic model and dse_out.json writer (cask_b200/host/include/Dse.hpp; the B200 edition
of src/runtime/Dse.cpp:32-74 and src/main.cpp:81-117), on fabricated plan statistics - no GPU needed.  The GPU run of
the whole tool (bin/cask_dse) is in tests/test_gpu_w_ingest.py."""
import json
import os
import subprocess

from conftest import ROOT


def test_model_and_json_shape(tmp_path):
    from cask_b200 import build
    build.build()
    exe, out = str(tmp_path / "dse_check"), str(tmp_path / "dse_out.json")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++14", "-Wall", "-Wno-sign-compare",
                           "-I" + os.path.join(ROOT, "cask_b200", "host", "include"),
                           os.path.join(ROOT, "tests", "cpp", "dse_check.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "cask_b200"), "-lcask_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "cask_b200")])
    lines = subprocess.check_output([exe, out], text=True).strip().split("\n")
    b2, s2, g2, fill2, b3, s3, g3 = (float(v) for v in lines[0].split())
    # C2: 10 B per stored entry + x once + y once = 1.107 GB -> 169 us at the measured 6 553 GB/s (measured kernel: 176 us)
    assert b2 == 10 * 5 * 16777216 + 16 * 16777216 and abs(s2 - b2 / 6553e9) < 1e-12 and abs(fill2 - 83869696 / (5 * 16777216)) < 1e-15
    assert 160e-6 < s2 < 180e-6 and 930 < g2 < 1050
    # R-MAT: gathers of a 268 MB x mostly miss a 126 MB L2
    miss = 1 - 126e6 / (8 * 33554432)
    assert abs(b3 - (12 * 5e8 + 4 * 32768 * 1024 + 16 * 33554432 + 32 * 5e8 * miss)) < 1
    assert lines[1] == "a" and lines[2].startswith("SkipEmpty 8192 16 1 1 ")
    assert lines[3].split() == ["2048", "4096", "8192", "16384"]
    d = json.load(open(out))
    assert set(d) == {"date", "took", "device", "best_architectures"} and d["took"] == "1.5" and d["device"] == "B200"
    a, b = d["best_architectures"]
    assert a["name"] == "SkipEmpty" and a["matrices"] == ["/m/poisson.mtx", "/m/other.mtx"]
    assert a["architecture_params"] == {"num_pipes": "1", "cache_size": "8192", "input_width": "16", "max_rows": "16777216",
                                        "num_controllers": "1"}   # the keys of main.cpp:70-79, values as strings (ptree)
    assert set(a["estimated_impl_params"]) >= {"memory_bandwidth", "device_bytes", "slices_staged_ell", "ell_fill"}
    assert float(a["estimated_gflops"]) > float(b["estimated_gflops"])
