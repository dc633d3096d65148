"""Static checks of bench.py / __graft_entry__.py that need no GPU: every name a function loads is bound somewhere
(a NameError on the GPU box would cost the round its bench line), and the reference arm runs end to end on CPU."""
import ast
import builtins
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _bound_names(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.arg):
            out.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
    return out


def _undefined(path):
    tree = ast.parse(open(path).read())
    module_names = _bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__"}
    bad = []

    def visit(fn, outer):
        scope = outer | _bound_names(fn)
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in scope:
                bad.append("%s:%d %s" % (os.path.basename(path), n.lineno, n.id))

    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)):
            visit(n, module_names)
    return bad


@pytest.mark.parametrize("name", ["bench.py", "__graft_entry__.py", os.path.join("cask_b200", "__init__.py")])
def test_no_unbound_names(name):
    assert _undefined(os.path.join(ROOT, name)) == []


def test_reference_arm_runs_on_cpu():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--no-cpu"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    impls = line["cpu_baseline"]["implementations"]
    assert set(impls) >= {"reference_code", "reference_symv"} and line["cpu_baseline"]["kind"] == "reference"
    # value is the faster of the reference's own two code paths; the arm's config matches the GPU arm's workload string
    assert line["value"] == max(v["gflops"] for v in impls.values() if "gflops" in v)
    other = line["other_cpu"]["implementations"]
    assert set(other) >= {"port", "mkl"} and other["port"]["gflops"] > 0
    assert line["config"]["workload"].startswith("C2: 2D 5-pt Poisson 4096x4096")
    assert line["cg"].get("value", 0) > 0 and line["cg"]["kind"] in ("reference", "port")


def test_clock_sampler_windows():
    """The sampler keeps the samples of the timed region, widens to the soak window when the region is too short to
    hold three samples, and reports throttle reasons only from the window it used."""
    import datetime
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    S = bench.ClockSampler
    base = 1.7e9

    def line(t, sm, power_cap="Not Active", thermal="Not Active"):
        ts = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return "%s, 0, %d, 1965, 700.0, 0x0000000000000004, Not Active, %s, Not Active, %s" % (ts, sm, thermal, power_cap)

    idle = [S.parse(base + 0.02 * i, line(base + 0.02 * i, 1200, thermal="Active")) for i in range(10)]
    soak = [S.parse(base + 0.3 + 0.02 * i, line(base + 0.3 + 0.02 * i, 1950)) for i in range(20)]
    timed = [S.parse(base + 0.8 + 0.02 * i, line(base + 0.8 + 0.02 * i, 1935, power_cap="Active")) for i in range(5)]
    assert all(p is not None for p in idle + soak + timed)
    assert abs(soak[0][0] - (base + 0.3)) < 2e-3 and soak[0][1:3] == (1950.0, 1965.0)
    r = S.summarize(idle + soak + timed, base + 0.8, base + 0.9, base + 0.3)
    assert r["window"] == "timed region" and r["samples"] == 5 and r["sm_mhz"] == 1935 and r["reasons"] == ["sw_power_cap"]
    r = S.summarize(idle + soak + timed, base + 0.8, base + 0.83, base + 0.3)   # 2 samples only: widen to the soak
    assert r["window"].startswith("soak + timed") and r["samples"] == 22 and r["sm_mhz"] == 1950
    assert "hw_thermal_slowdown" not in r["reasons"]                            # the idle samples are outside the window
    assert S.parse(base, "garbage") is None and S.parse(base, line(base, 1).replace(" 1,", " [N/A],")) is None
    # a clock that disagrees with the host by hours (time zone): the arrival time is used
    assert S.parse(base + 7200.0, line(base, 1800))[0] == base + 7200.0
    assert S.summarize([], 0.0, 1.0, 0.0)["sm_mhz"] is None


def test_value_dict_probe_is_contained():
    """The coded-format probe is a child process: without a GPU the child fails ("no CPU fallback") and the parent
    records the failure instead of raising; a child that outlives its timeout is killed and recorded the same way."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = argparse.Namespace(steps=5, warmup=3, soak=10, grid=64, cache=8192, probe_timeout=300.0, no_cg=True)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = bench.value_dict_probe(a)
    if has_gpu:
        assert "error" in r or r["value"] > 0
    else:
        assert "no CPU fallback" in r["error"]
    a.probe_timeout = 0.2
    assert "killed" in bench.value_dict_probe(a)["error"]
    assert "killed" in bench.rmat_stream_probe(a)["error"]
