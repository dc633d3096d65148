#!/bin/bash
# Multi-GPU evidence session (one box, N GPUs; charged N x the box time - keep it short):
#
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash profiles/gpu_session_multi.sh r2 8 full'
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900  -- 'bash profiles/gpu_session_multi.sh r2 2 quick'
#
# Writes under gpurun_out/<tag>_*: the multi-rank GPU tests (tests/test_gpu_dist.py: world 2 on both paths, 4 and 8 as
# the box allows), bench.py at N = 1, 2, 4 ... <gpus> with the driver's flags (C2 SpMV weak scaling + C4 CG / C5 BiCGStab /
# C3 SpMV strong scaling in the same line), and in `full` mode the reference arm at every N, the same run on the NCCL path
# (peer_mode 0) for an A/B of the peer-memory layer, and the in-situ kernel timeline of the CG loop on every rank
# (CASK_B200_TRACE; ncu cannot follow a multi-rank job) summarised by profiles/trace_summary.py.
# Every step has its own timeout; nothing here runs under a profiler, so the numbers are bench values.
TAG=${1:-multi}
GPUS=${2:-8}
MODE=${3:-full}
OUT=gpurun_out
mkdir -p $OUT
PY=python
PORT=29517
run() {  # run <n> <outfile> <bench args...>
  local n=$1 out=$2; shift 2
  if [ "$n" = 1 ]; then
    CASK_B200_BENCH_DETAILS=${out%.json}_details.json timeout 900 $PY bench.py --gpus 1 "$@" > $out 2> ${out%.json}.err
  else
    CASK_B200_BENCH_DETAILS=${out%.json}_details.json timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $n "$@" > $out 2> ${out%.json}.err
    PORT=$((PORT + 1))
  fi
  tail -c 300 ${out%.json}.err | grep -v Warning | tail -3
}
step() { echo "== $1 ($(date +%T))"; }

if [ "$MODE" != bench ]; then
step "multi-rank GPU tests"
timeout 1200 $PY -m pytest tests/test_gpu_dist.py -q -rs > $OUT/${TAG}_pytest_dist.log 2>&1; tail -6 $OUT/${TAG}_pytest_dist.log
fi

if [ "$MODE" = quick ] || [ "$MODE" = bench ]; then NS="$GPUS"; elif [ "$MODE" = scale ]; then NS="4 $GPUS"; else NS=""; N=1; while [ $N -le $GPUS ]; do NS="$NS $N"; N=$((N * 2)); done; fi
for N in $NS; do
  step "bench.py at N=$N"
  if [ "$MODE" = full ]; then run $N $OUT/${TAG}_bench_reference_n$N.json --impl reference --steps 20 --warmup 5 --no-cg; fi
  run $N $OUT/${TAG}_bench_n$N.json --steps 20 --warmup 5 --no-cpu --no-probe
done

if [ "$MODE" != quick ] && [ "$MODE" != bench ]; then
  step "N=$GPUS with the solvers' vectors NOT kept in the persisting L2 set-aside (l2_keep 0) for the A/B"
  CASK_B200_L2_KEEP=0 run $GPUS $OUT/${TAG}_bench_n${GPUS}_l2keep0.json --no-cpu --no-probe --only-bicgstab --steps 20 --warmup 5 --bicg-cap 300
  step "N=$GPUS on the NCCL path (peer_mode 0) for the A/B"
  CASK_B200_PEER=0 run $GPUS $OUT/${TAG}_bench_n${GPUS}_nccl.json --no-cpu --no-probe --only-bicgstab --steps 20 --warmup 5 --bicg-cap 300
  step "CG kernel timeline at N=$GPUS (peer path)"
  CASK_B200_TRACE=$OUT/${TAG}_trace_rank run $GPUS $OUT/${TAG}_bench_n${GPUS}_trace.json --no-cpu --no-extra --no-probe --steps 20 --soak 0
  $PY profiles/trace_summary.py $OUT/${TAG}_trace_rank > $OUT/${TAG}_trace_summary.txt 2>&1; head -30 $OUT/${TAG}_trace_summary.txt
fi

step "scaling table"
$PY - <<PYEOF | tee $OUT/${TAG}_scaling_table.md
import json, glob
rows = []
for f in sorted(glob.glob("$OUT/${TAG}_bench_n*.json")):
    if "reference" in f or f.endswith("_details.json"):
        continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    rows.append((f.split("/")[-1], d))
print("| run | N | SpMV GFLOP/s | ms/step | launches/step | e2e GFLOP/s | CG it/s (us/it) | BiCGStab it/s (its) | R-MAT ms | R-MAT max nnz share |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, d in rows:
    cg, bi, rm = d.get("cg") or {}, d.get("bicgstab") or {}, d.get("rmat") or {}
    print("| %s | %d | %.0f | %.4f | %.1f | %.1f | %s | %s | %s | %s |" % (
        name, d["n_gpus"], d["value"], d["ms_per_step"], d["gpu_launches"] / max(d["steps"], 1), d["e2e"]["value"],
        "%.0f (%s)" % (cg["iters_per_s"], "%.0f" % cg["us_per_iteration_marginal"] if cg.get("us_per_iteration_marginal") else "-") if "iters_per_s" in cg else cg.get("error"),
        "%.1f (%s)" % (bi["iters_per_s"], bi.get("iterations")) if "iters_per_s" in bi else bi.get("error"),
        "%.3f" % rm["ms_per_spmv"] if "ms_per_spmv" in rm else rm.get("error"),
        "%.3f" % rm["nnz_share_max_rank"] if "nnz_share_max_rank" in rm else "-"))
PYEOF
step "done"
