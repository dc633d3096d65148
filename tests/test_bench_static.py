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
