#!/usr/bin/env python
"""Generates the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref/libcaskref.so,
built in place from /root/reference by `make -C oracle ref`).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  matrices.npz   every test/matrices/*.mtx, test/systems/*.mtx and two test/test-benchmark
                 matrices AS THE REFERENCE READS THEM (io::readMatrix -> CsrMatrix arrays);
                 data fixtures only, no reference source.
  partitions.json  for each matrix x parameter set x {Simple, SkipEmpty}: all Partition scalars
                 (Spmv.hpp:25-29) + sha256 of the raw m_colptr / m_indptr_values bytes; the arrays
                 themselves for tiny cases.
  dots.npz       y = CsrMatrix::dot(x), x[i] = 0.25*i (test_spmv.cpp:27-28), from the reference.
  systems.json   test/systems: lower-triangle CSR as io::readSymMatrix gives it, rhs, the
                 solutions test/LinearSolvers.cpp:14-52 asserts.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refbind as R  # noqa: E402

REF = "/root/reference"

# (num_pipes, cache_size, input_width): first is the reference's smallest practical design shape,
# the rest stress stripe remainders, many blocks, width > row length, pipes > rows.
PARAM_SETS = [(1, 2048, 16), (2, 8, 4), (3, 16, 8), (1, 4, 2), (5, 64, 3), (48, 32, 2)]
BIG_PARAM_SETS = [(1, 2048, 16), (2, 4096, 8), (6, 1024, 48)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert R.available(), "run `make -C oracle ref` first"
    mats = {}
    files = sorted(os.listdir(f"{REF}/test/matrices"))
    paths = [f"{REF}/test/matrices/{f}" for f in files if f.endswith(".mtx")]
    paths += [f"{REF}/test/matrices/failing/test_large_empty2.mtx"]
    paths += [f"{REF}/test/systems/tiny.mtx", f"{REF}/test/systems/tinysym.mtx"]
    paths += [f"{REF}/test/test-benchmark/dw8192.mtx", f"{REF}/test/test-benchmark/t2d_q9_A_01.mtx"]
    npz, dots, parts_json = {}, {}, {}
    for p in paths:
        name = os.path.basename(p)[:-4]
        m = R.RefMatrix.read(p)
        rp, ci, va = m.csr()
        npz[name + ".dims"] = np.array([m.n, m.m, m.nnz], np.int64)
        npz[name + ".row_ptr"], npz[name + ".col_ind"], npz[name + ".values"] = rp, ci, va
        x = np.arange(m.m, dtype=np.float64) * 0.25
        dots[name] = m.dot(x)
        big = m.n > 5000
        cases = []
        for (pipes, cache, width) in (BIG_PARAM_SETS if big else PARAM_SETS):
            if m.n * ((m.m + cache - 1) // cache) > 40_000_000:
                continue
            for arch in (0, 1):
                res = m.preprocess(arch, pipes, cache, width)
                case = {"arch": arch, "num_pipes": pipes, "cache_size": cache, "input_width": width,
                        "estimated_clock_cycles": R.lib().ref_estimated_clock_cycles(m.h),
                        "partitions": []}
                for sc, colptr, pairs in res:
                    e = dict(sc)
                    e["colptr_sha256"] = sha(colptr)
                    e["pairs_sha256"] = sha(pairs)
                    if m.n <= 32 and len(colptr) <= 128 and len(pairs) <= 96:
                        e["colptr"] = [int(v) for v in colptr]
                        e["pairs_idx"] = [int(v) for v in pairs["indptr"]]
                        e["pairs_val"] = [float(v) for v in pairs["value"]]
                    case["partitions"].append(e)
                cases.append(case)
        parts_json[name] = cases
        print(name, m.n, m.m, m.nnz, len(cases), "cases")
    np.savez_compressed(os.path.join(HERE, "matrices.npz"), **npz)
    np.savez_compressed(os.path.join(HERE, "dots.npz"), **dots)
    with open(os.path.join(HERE, "partitions.json"), "w") as f:
        json.dump(parts_json, f, separators=(",", ":"))

    systems = {}
    for name, sol in (("tiny", [1, 2, 3, 4]), ("tinysym", [-2, 2, 3, 3])):
        a = R.RefMatrix.read(f"{REF}/test/systems/{name}.mtx", sym_lower=True)
        rp, ci, va = a.csr()
        systems[name] = {
            "n": a.n, "row_ptr": rp.tolist(), "col_ind": ci.tolist(), "values": va.tolist(),
            "rhs": R.read_vector(f"{REF}/test/systems/{name}_b.mtx").tolist(),
            "sol_file": R.read_vector(f"{REF}/test/systems/{name}_sol.mtx").tolist(),
            "asserted_solution": sol,  # test/LinearSolvers.cpp:25,45 (ASSERT_DOUBLE_EQ)
        }
    with open(os.path.join(HERE, "systems.json"), "w") as f:
        json.dump(systems, f, indent=1)


if __name__ == "__main__":
    main()
