#!/bin/bash
# Multi-GPU evidence session (one box, N GPUs; charged N x the box time - keep it short):
#
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash profiles/gpu_session_multi.sh r2 8'
#
# Writes under gpurun_out/<tag>_*: the 2-rank GPU tests (tests/test_gpu_dist.py), bench.py at N = 1, 2, 4 ... <gpus>
# (C2 SpMV weak scaling + C4 CG / C5 BiCGStab / C3 SpMV strong scaling in the same line), the same CG solve on the NCCL
# path (peer_mode 0) for an A/B of the peer-memory layer, and the in-situ kernel timeline of the CG loop on every rank
# (CASK_B200_TRACE; ncu cannot follow a multi-rank job) summarised by profiles/trace_summary.py.
# Every step has its own timeout; nothing here runs under a profiler, so the numbers are bench values.
TAG=${1:-multi}
GPUS=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
PY=python
PORT=29517
run() {  # run <n> <outfile> <bench args...>
  local n=$1 out=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 900 $PY bench.py --gpus 1 "$@" > $out 2> ${out%.json}.err
  else
    timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $n "$@" > $out 2> ${out%.json}.err
    PORT=$((PORT + 1))
  fi
  tail -c 400 $out; echo
}
step() { echo "== $1 ($(date +%T))"; }

step "2-rank GPU tests"
timeout 900 $PY -m pytest tests/test_gpu_dist.py -x -q > $OUT/${TAG}_pytest_dist.log 2>&1; tail -3 $OUT/${TAG}_pytest_dist.log

N=1
while [ $N -le $GPUS ]; do
  step "bench.py at N=$N"
  run $N $OUT/${TAG}_bench_n$N.json --no-cpu
  N=$((N * 2))
done

step "C4 CG at N=$GPUS on the NCCL path (peer_mode 0) for the A/B"
CASK_B200_PEER=0 run $GPUS $OUT/${TAG}_bench_n${GPUS}_nccl.json --no-cpu --no-extra --steps 50 --soak 100

step "CG kernel timeline at N=$GPUS (peer path)"
CASK_B200_TRACE=$OUT/${TAG}_trace_rank run $GPUS $OUT/${TAG}_bench_n${GPUS}_trace.json --no-cpu --no-extra --steps 20 --soak 0
$PY profiles/trace_summary.py $OUT/${TAG}_trace_rank > $OUT/${TAG}_trace_summary.txt 2>&1; head -30 $OUT/${TAG}_trace_summary.txt

step "scaling table"
$PY - <<PYEOF
import json, glob, re
rows = []
for f in sorted(glob.glob("$OUT/${TAG}_bench_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    cg = d.get("cg") or {}
    print("%-40s N=%d  SpMV %.0f GFLOP/s (%.3f ms)  e2e %.1f GFLOP/s  CG %s it/s (peer %s)  BiCGStab %s it/s  R-MAT %s ms" % (
        f.split("/")[-1], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"],
        "%.0f" % cg["iters_per_s"] if "iters_per_s" in cg else cg.get("error"), cg.get("peer_memory_path"),
        "%.1f" % d["bicgstab"]["iters_per_s"] if "iters_per_s" in (d.get("bicgstab") or {}) else (d.get("bicgstab") or {}).get("error"),
        "%.2f" % d["rmat_spmv"]["ms_per_spmv"] if "ms_per_spmv" in (d.get("rmat_spmv") or {}) else (d.get("rmat_spmv") or {}).get("error")))
PYEOF
step "done"
