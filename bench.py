#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: fp64 SpMV GFLOP/s (+ effective HBM GB/s vs the
roofline) on BASELINE.json's configs[1], the synthetic 2D 5-point Poisson 4096 x 4096 grid
(16.7M rows, 83.9M nnz per GPU), plus CG iterations/s on the 3D 27-point 256^3 system.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path, y = A x, over the whole (per-rank) matrix.
  value   device-resident: matrix, x and y live in HBM; K steps timed with CUDA events on the launch
          stream between barriers; max over ranks.
  e2e     the same metric through the host-buffer C ABI call cask_b200_spmv (what the reference's
          Spmv::spmv(const Vector&) maps to): pinned host x -> H2D, kernel, D2H of y, every step.
  roofline  algorithmic bytes 12 nnz + 8 (rows + cols) per launch / mean kernel time, against the
          measured HBM copy peak in MEASURED_PEAKS.json.
  cpu_baseline  the faster of two all-core CPU products on the same matrix: Intel MKL's mkl_sparse_d_mv (the
          library behind the reference's CPU solvers, reached through libtorch_cpu.so) and the OpenMP CSR
          port (oracle/).  `--impl reference` reports those two plus the compiled reference's own
          CsrMatrix::dot (oracle/_ref, bounded sample) and takes the fastest as its value; it also times
          the reference's own pcg<> on MKL (CG iterations/s, bounded sample of C4).
N > 1 (torchrun, one rank per GPU): every rank owns one 4096 x 4096 block of rows of a (4096 N) x 4096
grid (weak scaling); the x halo (one grid line per neighbour) travels over NCCL and overlaps the
interior slices.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 4096  # BASELINE.json configs[1]
CG_GRID = 256  # BASELINE.json configs[3]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 500; 20 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="untimed steps (default 10; 3 for --impl reference)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--grid", type=int, default=GRID)
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 (R-MAT SpMV) and C5 (BiCGStab) side measurements")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cache", type=int, default=8192)
    ap.add_argument("--value-dict", type=int, nargs="?", const=1, default=0,
                    help="coded staged ELL, bit-identical y; 1: values as 8-bit codes into per-slice tables (3 B per stored "
                         "nonzero instead of 10); 2: (value, displacement) pair codes (1 B); off by default until measured "
                         "on the GPU")
    ap.add_argument("--no-probe", action="store_true",
                    help="skip the coded-staged-ELL probe (a child bench.py --value-dict run, bounded by a timeout)")
    ap.add_argument("--probe-timeout", type=float, default=150.0, help="seconds a probe's child process may take")
    ap.add_argument("--soak", type=int, default=1500,
                    help="untimed launches between the warm-up and the timed region (clock sampling under load)")
    ap.add_argument("--only-rmat", action="store_true", help="profiling: run only the C3 R-MAT SpMV side measurement")
    ap.add_argument("--only-pcg-ilu", action="store_true", help="profiling: run only the ILU-preconditioned CG side measurement")
    ap.add_argument("--rmat-stripe", default=None, help="profiling, one GPU: W,r = run stripe r of W of the C3 matrix as rank r of a W-rank job would")
    ap.add_argument("--only-bicgstab", action="store_true", help="profiling: run only the C5 BiCGStab side measurement")
    ap.add_argument("--bicg-cap", type=int, default=4000, help="iteration cap of the C5 BiCGStab solve (profiling runs)")
    ap.add_argument("--cg-maxiters", type=int, default=2000, help="cap on CG iterations (profiling runs)")
    ap.add_argument("--cg-emulate-shard", type=int, default=0,
                    help="profiling: on ONE GPU, run CG on the stripe that rank W/2 of a W-way sharded C4 system owns "
                         "(principal submatrix, halo columns read zeros): the per-rank work of the W-GPU job under ncu")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 500
    if args.warmup is None:
        args.warmup = 3 if args.impl == "reference" else 10
    if any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR")):
        args.soak = 0  # under ncu every launch is replayed and serialised: no soak, and the numbers are not bench values
        args.no_probe = True  # and no child process for the profiler to follow
    return args


def algorithmic_bytes(nnz, rows, cols):
    return 12 * nnz + 8 * (rows + cols)


def workload_name(G, world, nnz_rank0=None):
    """config.workload, identical in both arms.  Rank 0's stripe of the (G*world) x G five-point grid: every row has
    5 entries minus the missing neighbours (left/right edges: 2G; top edge: G; bottom edge: G only when world == 1);
    the measuring arm passes the count of the stripe it actually generated."""
    n_local = G * G
    if nnz_rank0 is None:
        nnz_rank0 = 5 * n_local - 2 * G - G - (G if world == 1 else 0)
    return ("C2: 2D 5-pt Poisson %dx%d grid per GPU (%d rows, %d nnz per GPU), y = A x" % (G, G, n_local, nnz_rank0),
            n_local, nnz_rank0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Every sample carries
    nvidia-smi's own timestamp; stop(t0, t1) keeps the samples that fall inside the timed region [t0, t1] and, when the
    region is too short to hold three of them (K steps of a 0.2 ms kernel), the samples of the soak window that
    precedes it, where the GPU runs the very same launches back to back - the window used is named in the result."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self, wait_first_s=3.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
            return
        t_end = time.time() + wait_first_s  # nvidia-smi needs a few hundred ms before its first sample
        while not self.lines and time.time() < t_end and self.proc.poll() is None:
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    @staticmethod
    def parse(received, line):
        """(time, sm MHz, max MHz, [reasons]) of one csv line, or None."""
        f = [t.strip() for t in line.split(",")]
        if len(f) < 10:
            return None
        try:
            sm, mx = float(f[2]), float(f[3])
        except ValueError:
            return None
        try:
            import datetime
            ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            if abs(ts - received) > 5.0:  # clock / time-zone disagreement: trust the arrival time
                ts = received
        except ValueError:
            ts = received
        return ts, sm, mx, [n for n, v in zip(ClockSampler.NAMES, f[6:10]) if v.lower().startswith("active")]

    @staticmethod
    def summarize(samples, t0=None, t1=None, t_soak=None):
        window, sel = "all samples", samples
        if t0 is not None and t1 is not None:
            timed = [s for s in samples if t0 <= s[0] <= t1]
            soak = [s for s in samples if (t_soak if t_soak is not None else t0) <= s[0] <= t1]
            if len(timed) >= 3:
                window, sel = "timed region", timed
            elif soak:
                window, sel = "soak + timed region (same launches back to back; the timed region alone held %d samples)" % len(timed), soak
        reasons = sorted({r for s in sel for r in s[3]})
        return {"sm_mhz": statistics.median([s[1] for s in sel]) if sel else None,
                "sm_max_mhz": max(s[2] for s in sel) if sel else None, "samples": len(sel), "window": window,
                "reasons": reasons}

    def stop(self, t0=None, t1=None, t_soak=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "window": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        samples = [p for p in (self.parse(r, ln) for r, ln in list(self.lines)) if p]
        return self.summarize(samples, t0, t1, t_soak)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` (an instantiation name such as
    spmv_ell_persistent_kernel<2,false,0>) from the committed `ncu --set full` captures: profiles/roofline_traffic.json
    is WRITTEN by profiles/summarize_ncu.py from the raw reports, keyed by instantiation.  (bytes, source) or (None, None)
    when no capture of that instantiation is on record."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        kern = json.load(f).get("kernels", {})
    for rec in kern.values():
        if rec.get("kernel") == kernel:
            return rec.get("dram_bytes_per_launch"), rec.get("source")
    return None, None


def roofline_of(bytes_alg, bytes_stored, seconds, peak):
    """Achieved GB/s of one unit of work (an SpMV, a solver iteration) on algorithmic and on stored bytes, as fractions
    of the measured HBM peak."""
    out = {"algorithmic_bytes": int(bytes_alg), "achieved": bytes_alg / seconds / 1e9, "peak": peak, "unit": "GB/s"}
    out["frac"] = out["achieved"] / peak
    if bytes_stored:
        out["stored_bytes"] = int(bytes_stored)
        out["stored_gbs"] = bytes_stored / seconds / 1e9
        out["stored_frac"] = out["stored_gbs"] / peak
    return out


def cpu_allcore_spmv(grid, min_reps=10, max_reps=50, budget_s=5.0):
    """y = A x on ALL host cores, the same matrix as the GPU arm, two implementations, best of N each:
      mkl   Intel oneMKL's mkl_sparse_d_mv - the library the reference's CPU solvers are written against
            (pcg<> -> mkl_dcsrsymv, SparseLinearSolvers.hpp:189-206).  The image has no MKL package; the copy linked
            into libtorch_cpu.so exports the inspector-executor routine (oracle/mklbind.py);
      port  the OpenMP CSR row loop of oracle/cask_oracle.c (what Eigen's row-major product computes; Eigen is absent).
    The faster one is the CPU baseline."""
    import numpy as np
    from oracle import oraclebind as O
    n, rp, ci, va = O.gen_poisson2d(grid)
    nnz = len(va)
    x = (np.arange(n) % 1024) * 0.25
    y = np.zeros(n)

    def best_of(fn):
        fn()  # warm-up
        best, reps, t_all = 1e30, 0, time.perf_counter()
        while reps < min_reps or (time.perf_counter() - t_all < budget_s and reps < max_reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
            reps += 1
        return best, reps

    out = {}
    threads = [1]

    def port():
        threads[0] = O.csr_spmv_omp(n, rp, ci, va, x, y)[1]
    t, reps = best_of(port)
    y_port = y.copy()
    out["port"] = {"gflops": 2.0 * nnz / t / 1e9, "gbs": algorithmic_bytes(nnz, n, n) / t / 1e9, "ms": 1e3 * t,
                   "cores": int(threads[0]), "reps": reps, "what": "OpenMP CSR row loop (oracle/cask_oracle.c)"}
    try:
        from oracle import mklbind as M
        h = M.CsrHandle(n, n, rp, ci, va)
        t, reps = best_of(lambda: h.spmv(x, y))
        if not np.allclose(y, y_port, rtol=1e-12, atol=1e-9):
            raise RuntimeError("MKL result differs from the port")
        out["mkl"] = {"gflops": 2.0 * nnz / t / 1e9, "gbs": algorithmic_bytes(nnz, n, n) / t / 1e9, "ms": 1e3 * t,
                      "cores": M.max_threads(), "reps": reps, "what": "mkl_sparse_d_mv, " + M.version()}
    except Exception as e:  # MKL not reachable: the port stands alone
        out["mkl"] = {"error": str(e)}
    best = max((k for k in out if "gflops" in out[k]), key=lambda k: out[k]["gflops"])
    return out, best, n, nnz


def cpu_port_baseline(grid):
    """cpu_baseline of the main arm: the faster of MKL / the OpenMP port on the full workload, all host cores."""
    out, best, n, nnz = cpu_allcore_spmv(grid)
    b = out[best]
    return {"value": b["gflops"], "unit": "GFLOP/s", "cores": b["cores"], "kind": "port", "gbs": b["gbs"],
            "sample": "full workload: 2D 5-pt Poisson %dx%d, best of %d calls of the faster all-core CPU SpMV (%s: %s)"
                      % (grid, grid, b["reps"], best, b["what"]),
            "implementations": out}


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs run on rank 0 alone and are meant to use every
    host core, so the thread count is set explicitly BEFORE the OpenMP / MKL runtimes are loaded."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(cores)
    os.environ.pop("OMP_PROC_BIND", None)
    return cores


def reference_arm(args):
    """The reference's own CPU implementations of y = A x, timed on the GPU box's host cores (rank 0 only; the other
    ranks of a torchrun launch exit at once).  Two code paths of the reference compute the product, both compiled in
    place from /root/reference (oracle/_ref):
      reference_symv  the product inside its CPU solver, pcg<> (SparseLinearSolvers.hpp:175-206): one-based copies, then
                      mkl_dcsrsymv('l') on the stored lower triangle - Intel MKL, ALL host cores;
      reference_code  cask::CsrMatrix::dot (src/runtime/SparseMatrix.hpp:422-424), what its tests use as the CPU product:
                      single-threaded by construction and it rebuilds a hash map per call, so it runs on a bounded
                      512 x 512 sample of the operator, a bounded number of times (reported, never the headline).
    A step is one pass over the WHOLE job of the GPU arm at this N: the N stripes of the weak-scaling grid, one after the
    other on the same cores (every stripe is the same 4096 x 4096 operator up to its boundary rows, so the stripe matrix
    is built once and multiplied N times per step).  Exactly --warmup untimed and --steps timed steps are run; `value` =
    flops of those steps / their time (cpu_baseline.kind "reference").  For context the line also carries `other_cpu`:
    all-core CSR products that are NOT reference code (MKL's general mkl_sparse_d_mv and the OpenMP port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_cores()
    import numpy as np
    from oracle import oraclebind as O
    from oracle import refbind as R
    stripes = max(1, args.gpus)
    line = {"impl": "reference", "metric": "fp64 SpMV GFLOP/s", "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.grid, args.gpus)[0],
                       "step": "one pass over the %d stripe(s) of the job, %d host threads" % (stripes, cores)}}
    steps, warm = max(1, args.steps), max(0, args.warmup)
    impls = {}
    n, rp, ci, va = O.gen_poisson2d(args.grid)
    nnz = len(va)
    x = (np.arange(n) % 1024) * 0.25
    y_port = np.zeros(n)
    O.csr_spmv_omp(n, rp, ci, va, x, y_port)
    t_ref = None
    try:
        from oracle import mklbind as M
        if not M.ref_available():
            raise RuntimeError("oracle/_ref/libcaskref_mkl.so or MKL missing")
        rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp))
        keep = ci <= rows
        rpl = np.zeros(n + 1, np.int32)
        rpl[1:] = np.cumsum(np.bincount(rows[keep], minlength=n))
        cil, val = ci[keep].copy(), va[keep].copy()
        del rows
        if warm:
            M.symv(n, rpl, cil, val, x, reps=warm * stripes)
        y, sec = M.symv(n, rpl, cil, val, x, reps=steps * stripes)
        if not np.allclose(y, y_port, rtol=1e-12, atol=1e-9):
            raise RuntimeError("mkl_dcsrsymv result differs from the port")
        t_ref = sec
        impls["reference_symv"] = {
            "gflops": 2.0 * nnz * steps * stripes / sec / 1e9, "ms_per_step": 1e3 * sec / steps, "cores": M.max_threads(),
            "steps": steps, "warmup": warm, "kind": "reference",
            "what": "the product of the reference's pcg<>: mkl_dcsrsymv('l') on the stored lower triangle (%d of %d nnz per "
                    "stripe), %dx%d grid per stripe, flops counted for the full operator; %s"
                    % (int(keep.sum()), nnz, args.grid, args.grid, M.version())}
        del keep, rpl, cil, val, y
    except Exception as e:
        impls["reference_symv"] = {"error": str(e)}
    if R.available():
        sample_grid, reps = 512, max(1, min(steps, 5))  # 262 144 rows, 1.3M nnz: same operator, bounded
        sn, srp, sci, sva = O.gen_poisson2d(sample_grid)
        sx = (np.arange(sn) % 1024) * 0.25
        m = R.RefMatrix.from_csr(sn, sn, srp, sci, sva)
        m.dot(sx)
        t = 0.0
        for _ in range(reps):
            sy, sec1 = m.dot(sx, return_seconds=True)
            t += sec1
        assert np.array_equal(sy, O.csr_dot(sn, srp, sci, sva, sx))
        impls["reference_code"] = {
            "gflops": 2.0 * len(sva) * reps / t / 1e9, "ms_per_call": 1e3 * t / reps, "cores": 1, "calls": reps, "kind": "reference",
            "what": "cask::CsrMatrix::dot (reference code, single-threaded by construction) on a %dx%d sample of the operator"
                    % (sample_grid, sample_grid)}
    # all-core products that are not reference code, same step definition (bounded to ~20 s each)
    other = {}
    y = np.zeros(n)

    def timed(fn, label, what, cores_fn):
        fn()
        t0 = time.perf_counter()
        fn()
        once = time.perf_counter() - t0
        k = max(1, min(steps * stripes, int(20.0 / max(once, 1e-6))))
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        dt = time.perf_counter() - t0
        other[label] = {"gflops": 2.0 * nnz * k / dt / 1e9, "gbs": algorithmic_bytes(nnz, n, n) * k / dt / 1e9,
                        "ms_per_stripe": 1e3 * dt / k, "cores": cores_fn(), "calls": k, "what": what}
    thr = [1]

    def port():
        thr[0] = O.csr_spmv_omp(n, rp, ci, va, x, y)[1]
    timed(port, "port", "OpenMP CSR row loop (oracle/cask_oracle.c)", lambda: int(thr[0]))
    try:
        from oracle import mklbind as M
        h = M.CsrHandle(n, n, rp, ci, va)
        timed(lambda: h.spmv(x, y), "mkl", "mkl_sparse_d_mv, " + M.version(), M.max_threads)
        if not np.allclose(y, y_port, rtol=1e-12, atol=1e-9):
            raise RuntimeError("MKL result differs from the port")
    except Exception as e:
        other["mkl"] = {"error": str(e)}
    best_other = max((k for k in other if "gflops" in other[k]), key=lambda k: other[k]["gflops"])
    if "gflops" in impls.get("reference_symv", {}):
        b, kind, best = impls["reference_symv"], "reference", "reference_symv"
        v, ms_step = b["gflops"], b["ms_per_step"]
    else:  # no compiled reference / MKL on this machine: the port stands in, and says so
        b, kind, best = other[best_other], "port", best_other
        v, ms_step = b["gflops"], b["ms_per_stripe"] * stripes
    line.update({"value": v, "ms_per_step": ms_step,
                 "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": b["cores"], "kind": kind,
                                  "sample": "%s: %s; %d warm-up + %d timed steps of %d stripe pass(es)"
                                            % (best, b["what"], warm, steps, stripes),
                                  "implementations": impls},
                 "other_cpu": {"note": "all-core CSR products that are not reference code, same stripe matrix",
                               "implementations": other, "fastest": best_other, "gflops": other[best_other]["gflops"]},
                 "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    if not args.no_cg:
        try:
            line["cg"] = cpu_reference_cg()
        except Exception as e:  # informational
            line["cg"] = {"error": str(e)}
    print(json.dumps(line), flush=True)


def cpu_reference_cg(N=96):
    """CG iterations/s of the reference's OWN pcg<double, IdentityPreconditioner> (SparseLinearSolvers.hpp:162-239,
    compiled in place) on Intel MKL, all host cores, on a bounded sample of C4: the same 27-point operator and
    right-hand side on an N^3 grid, lower triangle handed over as the reference's callers do.  The reference hard-codes
    maxiters 2000 / tol 1e-5, so the sample is sized to converge within seconds.  Falls back to the sequential
    restatement (kind "port") where oracle/_ref or MKL is missing."""
    import numpy as np
    from oracle import oraclebind as O
    n, rp, ci, va = O.gen_poisson3d27(N)
    xt = 1.0 + 0.25 * (np.arange(n) % 4)
    b = O.csr_dot(n, rp, ci, va, xt)
    rows = np.repeat(np.arange(n), np.diff(rp))
    keep = ci <= rows
    rpl = np.zeros(n + 1, np.int32)
    rpl[1:] = np.cumsum(np.bincount(rows[keep], minlength=n))
    cil, val = ci[keep], va[keep]
    del rows, keep
    try:
        from oracle import mklbind as M
        if not M.ref_available():
            raise RuntimeError("oracle/_ref/libcaskref_mkl.so or MKL missing")
        conv, it, x, sec = M.pcg(n, rpl, cil, val, b, precon=0)
        kind, cores = "reference", M.max_threads()
        what = "reference pcg<double, IdentityPreconditioner> compiled in place, on " + M.version()
        trips = it + 2 if conv else 2000  # `iterations` holds the index of the last non-converged trip (:231)
    except Exception as e:
        t0 = time.perf_counter()
        conv, it, x = O.pcg(n, rpl, cil, val, b)
        sec = time.perf_counter() - t0
        kind, cores, what = "port", 1, "sequential C restatement of pcg (oracle/cask_oracle.c); MKL path unavailable: %s" % e
        trips = it + 2 if conv else 2000
    return {"value": trips / sec, "unit": "CG iterations/s", "cores": cores, "kind": kind, "converged": bool(conv),
            "loop_trips": trips, "seconds": sec, "max_abs_err_vs_x_true": float(np.abs(x - xt).max()),
            "row_iterations_per_s": trips * n / sec,
            "sample": "3D 27-pt Poisson %d^3 (%d rows, %d nnz; C4 is 256^3), %s" % (N, n, len(va), what)}


def child_bench(args, extra, env_extra=None):
    """One bounded child run of this script (its own process, its own CUDA context): returns the child's JSON line as a
    dict, or {"error": ...} when it exits non-zero, prints nothing parseable, or outlives --probe-timeout (killed).  A
    child that crashes or hangs costs its own record only, never the headline line of the parent."""
    cmd = [sys.executable, os.path.abspath(__file__), "--no-probe", "--no-cpu", "--steps", str(args.steps), "--warmup",
           str(args.warmup), "--soak", str(min(args.soak, 500)), "--grid", str(args.grid), "--cache", str(args.cache)] + list(extra)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env.update(env_extra or {})
    child = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)
    try:
        out, err = child.communicate(timeout=args.probe_timeout)
    except subprocess.TimeoutExpired:
        child.kill()
        try:
            child.communicate(timeout=30)
        except subprocess.TimeoutExpired:  # stuck in the driver: leave it behind rather than wait for it
            pass
        return {"error": "child exceeded %.0f s and was killed" % args.probe_timeout}
    if child.returncode != 0:
        return {"error": "child exit %d: %s" % (child.returncode, (err or out).strip()[-400:])}
    try:
        return json.loads(out.strip().splitlines()[-1])
    except (ValueError, IndexError):
        return {"error": "child printed no JSON line: %s" % out.strip()[-200:]}


def value_dict_probe(args, mode=1):
    """A coded staged-ELL format (option value_dict = mode, off by default) measured on the same workload in a CHILD
    process: `bench.py --value-dict <mode>` (C2 SpMV, then C4 CG with the same option).  The child checks y against the
    closed-form stencil result like the main arm does (a wrong y is a non-zero exit), so the record says whether the
    format is correct on this machine and what it would buy.  Informational: `value` stays the default format's number."""
    t0 = time.time()
    d = child_bench(args, ["--value-dict", str(mode), "--no-extra"] + (["--no-cg"] if args.no_cg else []))
    if "error" in d and "value" not in d:
        return d
    fmt = d.get("config", {}).get("format", {})
    cg = d.get("cg") or {}
    return {"what": "child process: bench.py --value-dict %d on the same workload; y checked against the closed-form stencil result" % mode,
            "format_in_use": fmt.get("value_dict"), "value": d.get("value"), "unit": d.get("unit"), "ms_per_step": d.get("ms_per_step"),
            "algorithmic_gbs": d.get("hbm_gbs"), "stored_bytes_per_launch": fmt.get("stored_bytes_per_launch"),
            "stored_gbs": (fmt["stored_bytes_per_launch"] / (d["ms_per_step"] * 1e-3) / 1e9
                           if fmt.get("stored_bytes_per_launch") and d.get("ms_per_step") else None),
            "table_entries_per_slice": fmt.get("table_entries_per_slice"), "kernel": d.get("roofline", {}).get("kernel"),
            "gpu_launches": d.get("gpu_launches"), "clocks": d.get("clocks"),
            "cg": {k: cg.get(k) for k in ("iters_per_s", "loop_trips", "converged", "max_abs_err_vs_x_true",
                                          "us_per_iteration_marginal", "error") if k in cg} or None,
            "seconds": time.time() - t0}


def rmat_stream_probe(args):
    """C3 with the CSR-stream variant of the gather kernel (option csr_stream, off by default), same child mechanism:
    the A/B beside `rmat_spmv`, which runs the default variant.  The child's spot check against torch's own product
    on sampled rows travels with it."""
    t0 = time.time()
    d = child_bench(args, ["--only-rmat", "--no-cg"], {"CASK_B200_CSR_STREAM": "1"})
    if "error" in d and "value" not in d:
        return d
    r = dict(d.get("rmat_spmv") or {"error": "child line carries no rmat_spmv"})
    r.pop("plan", None)
    r["what"] = "child process: bench.py --only-rmat with CASK_B200_CSR_STREAM=1 (products staged in shared memory)"
    r["seconds"] = time.time() - t0
    return r


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import cask_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cask_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = cb.Context(local)
    stream = torch.cuda.current_stream().cuda_stream
    ctx.set_stream(stream)
    if args.value_dict:
        ctx.set_option("value_dict", args.value_dict)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(cb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.dist_init(rank, world, bytes(idt.cpu().numpy().tobytes()))

    # ---- workload: rank r owns grid rows [r*G, (r+1)*G) of a (G*world) x G five-point grid ----------
    G = args.grid
    kind = cb.SYNTH_POISSON2D

    # The library's device generator makes square grids; the weak-scaling strip is (G*world) x G, and the stripes of
    # neighbouring ranks must couple through their i+-1 rows, so the stripe is assembled here with torch (bench-side
    # synthetic data) and cross-checked against the library's generator at world == 1.
    n_local = G * G
    n_global = n_local * world
    rows_i = torch.arange(n_local, device=dev, dtype=torch.int64) + rank * n_local
    gi, gj = rows_i // G, rows_i % G
    has = torch.stack([gi > 0, gj > 0, torch.ones_like(gi, dtype=torch.bool), gj < G - 1, gi < G * world - 1], 1)
    offs = torch.tensor([-G, -1, 0, 1, G], device=dev, dtype=torch.int64)
    vals5 = torch.tensor([-1.0, -1.0, 4.0, -1.0, -1.0], device=dev, dtype=torch.float64)
    cols = (rows_i[:, None] + offs[None, :])[has].to(torch.int32).contiguous()
    vals = vals5[None, :].expand(n_local, 5)[has].contiguous()
    rp = torch.zeros(n_local + 1, dtype=torch.int32, device=dev)
    rp[1:] = torch.cumsum(has.sum(1), 0).to(torch.int32)
    nnz_local = int(rp[-1].item())
    del rows_i, gi, gj, has
    if world == 1:
        # cross-check the torch-built stripe against the library's own device generator
        nnz_chk = cb.synth_nnz(kind, G, 0, n_local)
        assert nnz_chk == nnz_local
    dsg = cb.design(num_pipes=1, cache_size=args.cache, input_width=16)
    t0 = time.perf_counter()
    if world > 1:
        ctx.preprocess_shard_device(dsg, n_global, n_global, rank * n_local, n_local, nnz_local,
                                    rp.data_ptr(), cols.data_ptr(), vals.data_ptr())
    else:
        ctx.preprocess_device(dsg, n_local, n_local, nnz_local, rp.data_ptr(), cols.data_ptr(), vals.data_ptr())
    ctx.synchronize()
    preprocess_s = time.perf_counter() - t0
    stats = ctx.plan_stats()
    vd_active, vd_entries, vd_matrix_bytes = ctx.value_dict()
    vd_mode = ctx.value_dict_mode()

    x_full = ((torch.arange(n_global, device=dev) % 1024).double() * 0.25).contiguous()
    x_arena = ctx.dist_vector(0) if world > 1 else None
    if x_arena is not None:
        # sharded: x lives in the library's symmetric arena (cask_b200_dist_vector) - boundary rows are stored straight
        # into the neighbours' copies and a step is ONE SpMV launch, no NCCL call
        class _Arena:
            __cuda_array_interface__ = {"shape": (n_global,), "typestr": "<f8", "data": (x_arena, False), "version": 2}
        xa = torch.as_tensor(_Arena(), device=dev)
        xa.zero_()
        xa[rank * n_local:(rank + 1) * n_local] = x_full[rank * n_local:(rank + 1) * n_local]   # only the own slice
        x_ref_full, x_full = x_full, xa
    else:
        x_ref_full = x_full
    y = torch.empty(n_local, dtype=torch.float64, device=dev)

    # ---- device-resident timing ----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        ctx.spmv_device(x_full.data_ptr(), y.data_ptr())
    barrier()
    # soak: the same launch back to back for a few hundred ms (a fixed count, identical on every rank), untimed, so that
    # the clock sampler sees the GPU under exactly this load even when K steps last only a few milliseconds
    t_soak = time.time()
    for _ in range(args.soak):
        ctx.spmv_device(x_full.data_ptr(), y.data_ptr())
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        ctx.spmv_device(x_full.data_ptr(), y.data_ptr())
    e1.record()
    barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(t_begin, t_end, t_soak) if rank == 0 else None
    # correctness of the timed result: interior rows vanish for this x, boundary rows are known
    xg = x_ref_full.view(G * world, G)
    lo, hi = rank * G, (rank + 1) * G
    ref = 4 * xg[lo:hi].clone()
    ref[:, 1:] -= xg[lo:hi, :-1]
    ref[:, :-1] -= xg[lo:hi, 1:]
    if lo > 0:
        ref -= xg[lo - 1:hi - 1]
    else:
        ref[1:] -= xg[lo:hi - 1]
    if hi < G * world:
        ref -= xg[lo + 1:hi + 1]
    else:
        ref[:-1] -= xg[lo + 1:hi]
    if not torch.equal(y.view(G, G), ref):
        raise SystemExit("bench: SpMV result differs from the closed-form stencil result")
    del ref, xg
    one_launch = world > 1 and x_arena is not None

    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    nnz_total = nnz_local * world  # boundary ranks have 4096 entries fewer; weak-scaling approximation
    if world > 1:
        z = torch.tensor([nnz_local], dtype=torch.float64, device=dev)
        dist.all_reduce(z)
        nnz_total = int(z.item())
    flops = 2.0 * nnz_total
    value = flops * args.steps / (ms * 1e-3) / 1e9
    bytes_per_launch = algorithmic_bytes(nnz_local, n_local, n_local)
    kernel_ms = ms / args.steps
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- end to end: host buffers through the C ABI ---------------------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    if world == 1:
        hx = torch.empty(n_local, dtype=torch.float64).pin_memory()
        hy = torch.empty(n_local, dtype=torch.float64).pin_memory()
        hx.copy_(x_full.cpu())
        for _ in range(3):
            ctx.spmv_into(hx.numpy(), hy.numpy())
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.spmv_into(hx.numpy(), hy.numpy())
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert torch.equal(hy, y.cpu())
    else:
        hx = torch.empty(n_local, dtype=torch.float64).pin_memory()
        hy = torch.empty(n_local, dtype=torch.float64).pin_memory()
        hx.copy_(x_ref_full[rank * n_local:(rank + 1) * n_local].cpu())
        for _ in range(3):
            ctx.spmv_shard(hx.numpy(), hy.numpy())
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.spmv_shard(hx.numpy(), hy.numpy())
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert torch.equal(hy, y.cpu())
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = {"value": flops / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * n_local * world,
           "d2h_bytes_per_step": 8 * n_local * world, "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps, "host_buffers": "pinned",
           "api": "cask_b200_spmv(ctx, x_host, y_host)" if world == 1 else "cask_b200_spmv_shard(ctx, x_host_slice, y_host_slice) per rank"}
    if world == 1:
        # the same call with PAGEABLE buffers - what a caller holding a cask::Vector (std::vector<double>) sees through
        # host/include/Spmv.hpp: the copies then stage through the driver's bounce buffers
        px, py = hx.numpy().copy(), np.empty(n_local, dtype=np.float64)
        for _ in range(2):
            ctx.spmv_into(px, py)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.spmv_into(px, py)
        barrier()
        pg_s = (time.perf_counter() - t0) / e2e_steps
        assert np.array_equal(py, hy.numpy())
        e2e["pageable"] = {"value": flops / pg_s / 1e9, "ms_per_step": 1e3 * pg_s,
                           "how": "library-staged: pinned 4 MB rings filled / drained by host copy threads (hostcopy.hpp)"}
        # ... and with the staging left to the driver (what round 1 measured)
        ctx.set_option("host_staging", 0)
        ctx.spmv_into(px, py)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.spmv_into(px, py)
        barrier()
        e2e["pageable"]["driver_staged_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / 3
        ctx.set_option("host_staging", 1)
        del px, py
        # the floor under any host-buffer call: the same bytes over PCIe in both directions at once, no kernel
        try:
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            dx_probe, dy_probe = torch.empty(n_local, dtype=torch.float64, device=dev), torch.empty(n_local, dtype=torch.float64, device=dev)
            def duplex(up, dn):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    if up:
                        with torch.cuda.stream(s_up):
                            dx_probe.copy_(hx, non_blocking=True)
                    if dn:
                        with torch.cuda.stream(s_dn):
                            hy.copy_(dy_probe, non_blocking=True)
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / 5
            duplex(True, True)
            t_up, t_dn, t_both = duplex(True, False), duplex(False, True), duplex(True, True)
            e2e["pcie"] = {"h2d_gbs": 8.0 * n_local / t_up / 1e9, "d2h_gbs": 8.0 * n_local / t_dn / 1e9,
                           "duplex_ms_for_one_step": 1e3 * t_both,
                           "note": "pinned copies of one step's x and y alone, both directions at once: the floor of e2e"}
            del dx_probe, dy_probe
        except Exception as ex:  # noqa: BLE001 - informational
            e2e["pcie"] = {"error": repr(ex)[:200]}

    # ---- CG iterations / s on the 3D 27-point system (strong scaling: fixed 256^3 grid) -----------
    cg = bicg = rmat = pcg_ilu = None
    del x_full, x_ref_full, y, cols, vals, rp
    torch.cuda.empty_cache()
    # The side measurements never take the headline line down with them: a failure is recorded in place of the numbers.
    def side(fn, *a):
        try:
            return fn(*a)
        except Exception as e:  # noqa: BLE001 - reported, not swallowed
            return {"error": "%s: %s" % (type(e).__name__, e)}
        finally:
            try:
                torch.cuda.empty_cache()
            except Exception:
                pass
    if not args.no_cg:
        cg = side(bench_cg, ctx, cb, torch, dist, dev, rank, world, barrier, args.cg_maxiters, args.cg_emulate_shard)
    if not args.no_extra:
        if not args.only_rmat and not args.only_pcg_ilu:
            bicg = side(bench_bicgstab, ctx, cb, torch, dist, dev, rank, world, barrier, args.bicg_cap)
        if not args.only_bicgstab and not args.only_pcg_ilu:
            rmat = side(bench_rmat, ctx, cb, torch, dist, dev, rank, world, barrier, 25, 15, args.rmat_stripe)
        if world == 1 and not args.only_rmat and not args.only_bicgstab:
            pcg_ilu = side(bench_pcg_ilu, ctx, cb, torch, dev)

    if rank == 0:
        stored_per_launch = vd_matrix_bytes + 8 * (n_local + n_local) + 12 * stats.get("csr_nnz", 0)
        kernel = "spmv_ell_persistent_kernel<%d,false,%d>" % (stats.get("persist_ku", 2) or 2, vd_mode)
        traffic, traffic_src = ncu_traffic(kernel)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kernel,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "stored_bytes_per_launch": stored_per_launch,
                "stored_gbs": stored_per_launch / (kernel_ms * 1e-3) / 1e9,
                "stored_frac": stored_per_launch / (kernel_ms * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8tbs": achieved / 8000.0,
                "note": "achieved/frac: algorithmic bytes 12 nnz + 8 (rows + cols) (SURVEY 8d) / kernel time; stored_*: the bytes "
                        "the format really streams (10 B per stored nonzero: 16-bit x-cache positions) - frac > 1 is the "
                        "format moving fewer bytes than the algorithmic count, stored_frac is the kernel's distance from the copy peak"}
        line = {
            "metric": "fp64 SpMV GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": kernel_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(G, world, nnz_local)[0],
                       "sharding": "row stripes over ranks (Spmv.cpp:334-364)" + ("; x in the symmetric arena: ONE launch per step, whose last CTA stores the boundary rows "
                                   "into the neighbours' copies (flow-controlled by acknowledgements); no NCCL call"
                                   if one_launch else "; x halo over NCCL send/recv, interior and boundary launches" if world > 1 else ""),
                       "l2": "inputs (%.2f GB per launch) larger than L2 (126 MB); no flush needed" % (bytes_per_launch / 1e9),
                       "soak_steps": args.soak, "value_dict": vd_mode, "preprocess_s": preprocess_s},
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        details = {"config": {"global_rows": n_global, "design": {"num_pipes": 1, "cache_size": args.cache, "input_width": 16},
                              "preprocess_s": preprocess_s, "plan": stats,
                              "format": {"value_dict": vd_mode, "table_entries_per_slice": vd_entries,
                                         "stored_bytes_per_launch": stored_per_launch}}}
        # solver / R-MAT side measurements: fixed compact keys in the line, everything else in the details file
        def compact(src, keys):
            return {k: src[k] for k in keys if k in src} if src else None
        if cg:
            details["cg"] = cg
            line["cg"] = compact(cg, ("iters_per_s", "us_per_iteration_marginal", "loop_trips", "iterations_reported", "converged",
                                      "max_abs_err_vs_x_true", "gpu_launches", "peer_memory_path", "roofline", "error"))
        if pcg_ilu:
            details["pcg_ilu"] = pcg_ilu
            line["pcg_ilu"] = compact(pcg_ilu, ("iters_per_s", "loop_trips", "converged", "levels_per_triangular_solve", "error"))
            if "launched_one_by_one" in pcg_ilu:
                line["pcg_ilu"]["iters_per_s_launched_one_by_one"] = pcg_ilu["launched_one_by_one"]["iters_per_s"]
            if "level_graph" in pcg_ilu:
                line["pcg_ilu"]["iters_per_s_level_graph"] = pcg_ilu["level_graph"]["iters_per_s"]
        if bicg:
            details["bicgstab"] = bicg
            line["bicgstab"] = compact(bicg, ("iters_per_s", "iterations", "converged", "rel_residual", "max_abs_err_vs_ones",
                                              "gpu_launches", "timed", "roofline", "clocks", "error"))
        if rmat:
            details["rmat"] = rmat
            line["rmat"] = compact(rmat, ("ms_per_spmv", "gflops", "nnz", "kernel", "l2_hit_rate_on_x_pct", "max_err_all_rows_rel_to_sum_abs",
                                          "preprocess_s", "nnz_share_max_rank", "roofline", "error"))
        if world == 1 and not args.value_dict and not args.no_probe:
            probes = [("pair_dict_probe", lambda a: value_dict_probe(a, 2))]
            timed_out = False
            for key, fn in probes:
                if timed_out:  # one child already cost its full time limit: the run stays within minutes
                    details[key] = {"error": "skipped: an earlier probe exceeded its time limit"}
                    continue
                try:
                    details[key] = fn(args)
                except Exception as e:  # noqa: BLE001 - informational
                    details[key] = {"error": "%s: %s" % (type(e).__name__, e)}
                timed_out = "was killed" in str(details[key].get("error", ""))
            pr = details.get("pair_dict_probe") or {}
            line["pair_dict_probe"] = {k: pr.get(k) for k in ("value", "ms_per_step", "stored_gbs", "error") if k in pr}
            if pr.get("cg"):
                line["pair_dict_probe"]["cg_iters_per_s"] = pr["cg"].get("iters_per_s")
        if world == 1 and not args.no_cpu:
            try:
                cpu = cpu_port_baseline(G)
            except Exception as e:
                cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}
            details["cpu_baseline"] = cpu
            line["cpu_baseline"] = {k: cpu.get(k) for k in ("value", "unit", "cores", "kind", "sample")}
            if cg and "error" not in cg:
                try:
                    ccpu = cpu_reference_cg()
                except Exception as e:
                    ccpu = {"value": None, "unit": "CG iterations/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}
                details["cg"]["cpu_baseline"] = ccpu
                line["cg"]["cpu_baseline"] = {k: ccpu.get(k) for k in ("value", "unit", "cores", "kind", "row_iterations_per_s")}
        dpath = os.environ.get("CASK_B200_BENCH_DETAILS") or os.path.join(
            ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "", "bench_details_n%d.json" % world)
        try:
            with open(dpath, "w") as f:
                json.dump(dict(line, details=details), f, indent=1)
            line["details_file"] = os.path.relpath(dpath, ROOT)
        except OSError:
            pass
        # the headline stays compact (<= 4 KB): explanatory strings live in the details file, which holds the full line

        def slim(o):
            if isinstance(o, dict):
                return {k: slim(v) for k, v in o.items() if k not in ("note", "per", "how", "what", "peak_source", "window")}
            if isinstance(o, float):
                return float("%.6g" % o)
            return o
        line = slim(line)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def stored_spmv_bytes(ctx, torch, dist, dev, world, n_local):
    """Bytes one SpMV streams in the stored format, whole job: staged-ELL entries (10 / 3 / 1 B each) + gather-CSR
    entries (12 B + 4 B per row) + the rank's x slice read once + y written once, summed over the ranks."""
    st = ctx.plan_stats()
    b = ctx.value_dict()[2] + 12 * st["csr_nnz"] + 4 * st["csr_rows"] + 16 * n_local
    t = torch.tensor([float(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return float(t.item())


def bench_cg(ctx, cb, torch, dist, dev, rank, world, barrier, maxiters=2000, emulate=0):
    """BASELINE configs[3]: CG (the reference's pcg loop, identity preconditioner) on the 3D 27-point
    256^3 Poisson system, row-sharded over the ranks; b = A x_true, x_true[k] = 1 + 0.25 (k mod 4)."""
    N = CG_GRID
    kind = cb.SYNTH_POISSON3D27
    n = cb.synth_rows(kind, N)
    r0, nr = cb.shard_rows(n, world, rank)
    if emulate > 1 and world == 1:
        r0, nr = cb.shard_rows(n, emulate, emulate // 2)
    nnz = cb.synth_nnz(kind, N, r0, nr)
    rp = torch.empty(nr + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    va = torch.empty(nnz, dtype=torch.float64, device=dev)
    cb.synth_device(kind, N, r0, nr, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), torch.cuda.current_stream().cuda_stream)
    dsg = cb.design(num_pipes=1, cache_size=8192, input_width=16)
    if world > 1 or emulate > 1:
        ctx.preprocess_shard_device(dsg, n, n, r0, nr, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    else:
        ctx.preprocess_device(dsg, n, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    xt = (1.0 + 0.25 * (torch.arange(n, device=dev) % 4).double()).contiguous()
    b = torch.empty(nr, dtype=torch.float64, device=dev)
    ctx.spmv_device(xt.data_ptr(), b.data_ptr())
    ctx.synchronize()
    x = torch.zeros(nr, dtype=torch.float64, device=dev)
    ctx.cg_device(b.data_ptr(), x.data_ptr(), maxiters=20)  # warm-up
    x.zero_()
    barrier()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    conv, iters, rs, trips = ctx.cg_device(b.data_ptr(), x.data_ptr(), maxiters=maxiters, tol=1e-5)
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    launches = ctx.launch_count() - l0
    err = float((x - xt[r0:r0 + nr]).abs().max().item())
    # marginal cost of an iteration (solve set-up and the final synchronisation excluded): two capped solves
    tm, caps, marginal_us = [], (50, 250), None
    if maxiters >= caps[1] and trips >= caps[1]:
        for cap in caps:
            x.zero_()
            barrier()
            t1 = time.perf_counter()
            ctx.cg_device(b.data_ptr(), x.data_ptr(), maxiters=cap, tol=1e-5)
            barrier()
            tm.append(time.perf_counter() - t1)
        marginal_us = (tm[1] - tm[0]) / float(caps[1] - caps[0]) * 1e6
    et = torch.tensor([err], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    nnz_total = cb.synth_nnz(kind, N, 0, n)
    stored = stored_spmv_bytes(ctx, torch, dist, dev, world, nr)
    sec_per_it = (marginal_us * 1e-6) if marginal_us else dt / max(trips, 1)
    roof = roofline_of(algorithmic_bytes(nnz_total, n, n) + 72 * n, stored + 72 * n, sec_per_it, measured_peak()[0] * world)
    roof["per"] = "CG iteration (marginal cost), whole job; peak = %d x measured HBM copy peak; bytes = SpMV + 72 n (SURVEY 8d)" % world
    return {"workload": "C4: CG on 3D 27-pt Poisson %d^3 (%d rows, %d nnz), row-sharded over %d GPU(s)" % (N, n, nnz_total, world),
            "roofline": roof,
            "scaling": "strong", "iters_per_s": trips / dt, "loop_trips": trips, "iterations_reported": iters,
            "converged": conv, "rs_final": rs, "seconds": dt, "max_abs_err_vs_x_true": float(et.item()),
            "gpu_launches": int(launches), "us_per_iteration_marginal": marginal_us,
            "peer_memory_path": bool(ctx.peer_active()) if world > 1 else None,
            "traffic_bound_iters_per_s_1gpu": 1.0 / ((algorithmic_bytes(nnz_total, n, n) + 72 * n) / (measured_peak()[0] * 1e9))}


def bench_pcg_ilu(ctx, cb, torch, dev, N=64):
    """pcg<double, ILUPreconditioner> with a unit lower solve on the 3D 27-point Poisson twin N^3
    (single rank: the dependency levels of ILU(0) cross every stripe).  One launch per level, 7 (N - 1) + 1 levels per
    triangular solve; all levels run in ONE cooperative kernel behind grid barriers (default), as a CUDA graph of the
    per-level kernels (ilu_persistent = 0), or kernel by kernel (ilu_graph = 0 as well)."""
    kind = cb.SYNTH_POISSON3D27
    n = cb.synth_rows(kind, N)
    nnz = cb.synth_nnz(kind, N, 0, n)
    rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    va = torch.empty(nnz, dtype=torch.float64, device=dev)
    cb.synth_device(kind, N, 0, n, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), torch.cuda.current_stream().cuda_stream)
    xt = (1.0 + 0.25 * (torch.arange(n, device=dev) % 4).double()).contiguous()
    b = torch.empty(n, dtype=torch.float64, device=dev)
    out = {"workload": "pcg<ILU(0), unit lower solve> on 3D 27-pt Poisson %d^3 (%d rows, %d nnz), one GPU" % (N, n, nnz),
           "levels_per_triangular_solve": 7 * (N - 1) + 1}
    for mode, (persistent, graph) in (("persistent", (1, 1)), ("graph", (0, 1)), ("launches", (0, 0))):
        ctx.set_option("ilu_persistent", persistent)
        ctx.set_option("ilu_graph", graph)
        ctx.preprocess_device(cb.design(num_pipes=1, cache_size=8192, input_width=16), n, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
        ctx.spmv_device(xt.data_ptr(), b.data_ptr())
        ctx.synchronize()
        x = torch.zeros(n, dtype=torch.float64, device=dev)
        ctx.pcg_device(b.data_ptr(), x.data_ptr(), cb.PRECON_ILU_UNIT, maxiters=3)   # factorisation + graph capture outside the timing
        x.zero_()
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        conv, it, rs = ctx.pcg_device(b.data_ptr(), x.data_ptr(), cb.PRECON_ILU_UNIT, maxiters=500, tol=1e-5)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        trips = it + 2 if conv else it + 1   # `iterations` = index of the last non-converged iteration (SparseLinearSolvers.hpp:231)
        rec = {"iters_per_s": trips / dt, "loop_trips": trips, "converged": bool(conv), "seconds": dt, "rs_final": rs,
               "max_abs_err_vs_x_true": float((x - xt).abs().max().item()), "kernel_nodes_or_launches": int(ctx.launch_count() - l0)}
        if mode == "persistent":
            out.update(rec)
        elif mode == "graph":
            out["level_graph"] = rec
        else:
            out["launched_one_by_one"] = rec
    ctx.set_option("ilu_persistent", 1)
    ctx.set_option("ilu_graph", 1)
    return out


def bench_bicgstab(ctx, cb, torch, dist, dev, rank, world, barrier, cap=4000):
    """BASELINE configs[4]: BiCGStab (Eigen's loop, Jacobi preconditioner) on the nonsymmetric 3D 7-point
    convection-diffusion system, 512^3 grid (134M rows, 0.94B nnz), row-sharded; b = A 1; 40 iterations timed."""
    N = 512
    kind = cb.SYNTH_CONVDIFF3D7
    n = cb.synth_rows(kind, N)
    r0, nr = cb.shard_rows(n, world, rank)
    nnz = cb.synth_nnz(kind, N, r0, nr)
    rp = torch.empty(nr + 1, dtype=torch.int32, device=dev)
    ci = torch.empty(nnz, dtype=torch.int32, device=dev)
    va = torch.empty(nnz, dtype=torch.float64, device=dev)
    cb.synth_device(kind, N, r0, nr, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), torch.cuda.current_stream().cuda_stream)
    dsg = cb.design(num_pipes=1, cache_size=8192, input_width=16)
    if world > 1:
        ctx.preprocess_shard_device(dsg, n, n, r0, nr, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    else:
        ctx.preprocess_device(dsg, n, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    ones = torch.ones(n, dtype=torch.float64, device=dev)
    b = torch.empty(nr, dtype=torch.float64, device=dev)
    ctx.spmv_device(ones.data_ptr(), b.data_ptr())
    ctx.synchronize()
    del ones
    x = torch.zeros(nr, dtype=torch.float64, device=dev)
    ctx.bicgstab_device(b.data_ptr(), x.data_ptr(), tol=1e-10, maxit=3)  # warm-up
    barrier()
    # timed: the WHOLE solve to Eigen's stopping test ||r|| <= tol ||b|| with tol = 1e-10 (the default
    # DBL_EPSILON is unreachable in practice), iteration cap 4000
    tol = 1e-10
    l0 = ctx.launch_count()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    t_wall0 = time.time()
    t0 = time.perf_counter()
    its, err = ctx.bicgstab_device(b.data_ptr(), x.data_ptr(), tol=tol, maxit=cap)
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    launches = int(ctx.launch_count() - l0)
    e = torch.tensor([float((x - 1.0).abs().max().item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    nnz_total = cb.synth_nnz(kind, N, 0, n)
    stored = stored_spmv_bytes(ctx, torch, dist, dev, world, nr)
    # vector streams of one iteration beside the two SpMVs (8 n bytes each): p-update 6 (r, p, v, 1/d in; p, y out),
    # r0 in the first SpMV's dot 1, s-update 5 (r, v, 1/d in; s, z out), s in the second SpMV's dots 1, x/r-update 8
    # (y, z, s, t, r0, x in; x, r out) = 21; the unfused path reads t and s once more (dot2 kernel) = 22 + 1
    streams = 21 if launches <= 6 * max(its, 1) else 23
    sec_per_it = dt / max(its, 1)
    roof = roofline_of(2 * algorithmic_bytes(nnz_total, n, n) + 8 * streams * n, 2 * stored + 8 * streams * n, sec_per_it,
                       measured_peak()[0] * world)
    roof["per"] = ("BiCGStab iteration (whole solve / iterations), whole job; peak = %d x measured HBM copy peak; bytes = 2 SpMV + %d vector "
                   "streams of 8 n bytes (p-update 6, s-update 5, x/r-update 8, dot operands)" % (world, streams))
    return {"workload": "C5: BiCGStab (Jacobi) on 3D 7-pt convection-diffusion %d^3 (%d rows, %d nnz), row-sharded over %d GPU(s)"
                        % (N, n, nnz_total, world),
            "scaling": "strong", "iters_per_s": its / dt, "iterations": its, "converged": bool(err <= tol), "tol": tol,
            "rel_residual": err, "seconds": dt, "timed": "whole solve to ||r|| <= 1e-10 ||b|| (cap %d)" % cap,
            "max_abs_err_vs_ones": float(e.item()), "spmv_per_iteration": 2, "vector_streams_per_iteration": streams,
            "gpu_launches": launches, "roofline": roof, "clocks": clocks}


def bench_rmat(ctx, cb, torch, dist, dev, rank, world, barrier, scale=25, edge_factor=15, rmat_stripe=None):
    """BASELINE configs[2]: R-MAT power-law matrix, 2^25 rows, ~5e8 nnz, (a,b,c,d) = (0.57,0.19,0.19,0.05),
    duplicates merged, rows sorted by column; generated on the device with torch (bench-side synthetic data),
    row-sharded; y = A x with x gathered through L2 (irregular rows -> vector-per-row kernels)."""
    n = 1 << scale
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    E = edge_factor << scale
    keys = []
    chunk = 1 << 26
    for e0 in range(0, E, chunk):
        m = min(chunk, E - e0)
        row = torch.zeros(m, dtype=torch.int64, device=dev)
        col = torch.zeros(m, dtype=torch.int64, device=dev)
        for _ in range(scale):
            u = torch.rand(m, device=dev, generator=g)
            rb = (u >= 0.76).to(torch.int64)                       # quadrants c, d
            cbit = (((u >= 0.57) & (u < 0.76)) | (u >= 0.95)).to(torch.int64)  # quadrants b, d
            row = row * 2 + rb
            col = col * 2 + cbit
        keys.append((row << scale) | col)                          # every rank draws the same edge list (same seed)
        del row, col, u, rb, cbit
    key = torch.cat(keys)
    del keys
    # Row stripes.  One GPU: everything.  Sharded: stripes of (nearly) equal WORK, cut at multiples of 1024 rows from the
    # edge histogram - under the reference's equal-row rule (Spmv.cpp:334) rank 0 of 8 would own 0.76^3 = 44 % of this
    # matrix and bound the whole job; the library takes any contiguous partition (cask_b200_preprocess_shard_device).
    # Work = merge items of the gather kernel = nonzeros + rows (CASK_B200_RMAT_BALANCE=nnz: nonzeros only).
    # --rmat-stripe W,r (one GPU): stripe r of W as a rectangular matrix with the compact column numbering of the sparse
    # exchange (col_reorder 2) - what rank r's kernels do in a W-rank job, measurable (and profilable) on one GPU.
    emu = None
    if world == 1 and rmat_stripe:
        emu = tuple(int(v) for v in rmat_stripe.split(","))
    parts, part = (world, rank) if world > 1 else (emu if emu else (1, 0))
    balance = os.environ.get("CASK_B200_RMAT_BALANCE", "items")
    if parts > 1:
        hist = torch.bincount(key >> (scale + 10), minlength=n >> 10).double()
        if balance != "nnz":
            hist = hist + 1024.0
        cum = torch.cumsum(hist, 0)
        targets = cum[-1] * torch.arange(1, parts, device=dev, dtype=torch.float64) / parts
        cuts = (torch.searchsorted(cum, targets) + 1).clamp(max=n >> 10) << 10
        bounds = [0] + [int(c) for c in cuts.tolist()] + [n]
        for i in range(1, len(bounds)):
            bounds[i] = max(bounds[i], bounds[i - 1])
        r0, nr = bounds[part], bounds[part + 1] - bounds[part]
        del hist, cum
    else:
        r0, nr = 0, n
    keep = ((key >> scale) >= r0) & ((key >> scale) < r0 + nr)
    key = key[keep] - (r0 << scale)
    del keep
    key = torch.unique(key, sorted=True)
    nnz = int(key.numel())
    rows_l = (key >> scale)
    ci = (key & (n - 1)).to(torch.int32).contiguous()
    del key
    counts = torch.bincount(rows_l, minlength=nr)
    del rows_l
    rp = torch.zeros(nr + 1, dtype=torch.int32, device=dev)
    rp[1:] = torch.cumsum(counts, 0).to(torch.int32)
    del counts
    va = (torch.rand(nnz, device=dev, dtype=torch.float64, generator=g) * 2.0 - 1.0).contiguous()
    dsg = cb.design(num_pipes=1, cache_size=8192, input_width=16)
    t0 = time.perf_counter()
    if world > 1:
        ctx.preprocess_shard_device(dsg, n, n, r0, nr, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
    else:
        if emu:
            ctx.set_option("col_reorder", int(os.environ.get("CASK_B200_STRIPE_REORDER", "2")))
        ctx.preprocess_device(dsg, nr, n, nnz, rp.data_ptr(), ci.data_ptr(), va.data_ptr())
        if emu:
            ctx.set_option("col_reorder", int(os.environ.get("CASK_B200_COL_REORDER", "-1")))
    ctx.synchronize()
    prep = time.perf_counter() - t0
    stats = ctx.plan_stats()
    # x from its own generator: g has drawn a different number of values on every rank by now (va has the rank's nnz
    # entries), and only the OWN slice of x is an input of the sharded call - the check below needs the same x everywhere
    gx = torch.Generator(device=dev)
    gx.manual_seed(2)
    x = torch.rand(n, device=dev, dtype=torch.float64, generator=gx).contiguous()
    y = torch.empty(nr, dtype=torch.float64, device=dev)
    for _ in range(3):
        ctx.spmv_device(x.data_ptr(), y.data_ptr())
    barrier()
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctx.spmv_device(x.data_ptr(), y.data_ptr())
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    z = torch.tensor([float(nnz)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(z)
    ms, nnz_total = float(ms.item()), int(z.item())
    # spot check against torch's own CSR product on a sample of rows
    sample = torch.randint(0, nr, (4096,), device=dev, generator=g)
    lo, hi = rp[sample].long(), rp[sample + 1].long()
    ref = torch.stack([(va[a:b] * x[ci[a:b].long()]).sum() for a, b in zip(lo.tolist()[:256], hi.tolist()[:256])])
    got = y[sample[:256]]
    rel = float(((got - ref).abs() / (ref.abs() + 1e-30)).max().item())
    # ... and against a full fp64 product by torch's CSR kernel (library code, checker only): every row, hubs included
    a_t = torch.sparse_csr_tensor(rp.long(), ci.long(), va, size=(nr, n))
    y_ref = torch.mv(a_t, x)
    a_abs = torch.sparse_csr_tensor(rp.long(), ci.long(), va.abs(), size=(nr, n))
    scale_rows = torch.mv(a_abs, x.abs()).clamp_min(1e-300)
    rel_all = float(((y - y_ref).abs() / scale_rows).max().item())
    del a_t, a_abs, y_ref, scale_rows
    share = torch.tensor([float(nnz)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(share, op=dist.ReduceOp.MAX)
    stored = 12.0 * nnz_total + 4.0 * n + 16.0 * n
    reordered = torch.tensor([float(stats.get("cols_referenced", 0)) if stats.get("col_reorder") else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(reordered)
    stored += 20.0 * float(reordered.item())   # hub clustering: perm (4 B) + x read + permuted x written per referenced column
    roof = roofline_of(algorithmic_bytes(nnz_total, n, n), stored, ms * 1e-3, measured_peak()[0] * world)
    # instantiation name as ncu prints it (template arguments: fused dot, items per thread, resident CTAs per SM)
    kernel = ("spmv_csr_merge_kernel<0,%d,%d>" % (stats.get("merge_items", 7), stats.get("merge_ctas", 5))
              if stats.get("csr_kernel") == 1 else "spmv_csr_items_kernel<0,0>")
    roof["traffic"], roof["traffic_source"] = ncu_traffic(kernel)
    roof["per"] = "SpMV, whole job; peak = %d x measured HBM copy peak; the kernel is bound by the x gather, not by these bytes" % world
    l2 = None
    tj = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tj):
        with open(tj) as f:
            for rec in json.load(f).get("kernels", {}).values():
                if rec.get("kernel") == kernel and rec.get("l2_hit_rate_on_x_pct") is not None:
                    l2 = rec["l2_hit_rate_on_x_pct"]
    return {"workload": "C3: R-MAT scale %d, %d rows, %d nnz (after merging duplicates), y = A x, row-sharded over %d GPU(s)"
                        % (scale, n, nnz_total, world),
            "scaling": "strong", "ms_per_spmv": ms, "gflops": 2.0 * nnz_total / (ms * 1e-3) / 1e9, "nnz": nnz_total,
            "algorithmic_gbs": algorithmic_bytes(nnz_total, n, n) / (ms * 1e-3) / 1e9, "preprocess_s": prep,
            "kernel": kernel, "l2_hit_rate_on_x_pct": l2, "roofline": roof, "col_reorder": int(stats.get("col_reorder", 0)),
            "nnz_share_max_rank": float(share.item()) / nnz_total,
            "stripes": "one" if parts == 1 else "equal %s, cut at multiples of 1024 rows (rows of this %s: %d)" % (
                "nonzero count" if balance == "nnz" else "nonzeros + rows", "rank" if world > 1 else "emulated stripe %d of %d" % (part, parts), nr),
            "rows_this_rank": nr, "nnz_this_rank": nnz,
            "max_rel_diff_256_sampled_rows": rel, "max_err_all_rows_rel_to_sum_abs": rel_all,
            "plan": {k: stats[k] for k in ("slices_staged_ell", "slices_gather_csr", "csr_lanes_per_row", "max_row_length",
                                           "row_length_histogram", "csr_items", "csr_kernel", "col_reorder", "cols_referenced",
                                           "merge_items", "merge_ctas")}}


if __name__ == "__main__":
    main()
