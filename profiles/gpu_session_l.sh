#!/bin/bash
# Multi-GPU session L: sparse exchange of the row-sharded gather path (dist_sparse) - tests, then C3 A/B against the
# broadcast of every slice.   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash profiles/gpu_session_l.sh r2l 2'
TAG=${1:-r2l}
GPUS=${2:-2}
MODE=${3:-full}
OUT=gpurun_out
mkdir -p $OUT
PY=python
PORT=29617
step() { echo "== $1 ($(date +%T))"; }
run() {  # run <outfile> <env...> -- <bench args...>
  local out=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" CASK_B200_BENCH_DETAILS=${out%.json}_details.json timeout 600 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS \
    --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $GPUS "$@" > $out 2> ${out%.json}.err
  PORT=$((PORT + 1))
  $PY - $out <<'PYEOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("rmat") or d.get("rmat_spmv") or {}
    print(sys.argv[1].split("/")[-1], {k: r.get(k) for k in ("ms_per_spmv", "max_err_all_rows_rel_to_sum_abs", "col_reorder", "preprocess_s", "nnz_share_max_rank", "error")})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PYEOF
  tail -c 400 ${out%.json}.err | grep -v Warning | tail -2
}
if [ "$MODE" != bench ]; then
step "multi-rank GPU tests (world 2 on both paths; 4 and 8 as the box allows)"
timeout 1200 $PY -m pytest tests/test_gpu_dist.py -q -rs -x > $OUT/${TAG}_pytest_dist.log 2>&1; tail -25 $OUT/${TAG}_pytest_dist.log
fi
step "C3 at N=$GPUS: sparse exchange, one kernel over peer memory (default)"
run $OUT/${TAG}_rmat_n${GPUS}_sparse.json CASK_B200_DIST_SPARSE=1 -- --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0
step "C3 at N=$GPUS: sparse exchange over NCCL send/recv (peer_mode 0)"
run $OUT/${TAG}_rmat_n${GPUS}_sparse_nccl.json CASK_B200_DIST_SPARSE=1 CASK_B200_PEER=0 -- --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0
step "C3 at N=$GPUS: every slice broadcast to all"
run $OUT/${TAG}_rmat_n${GPUS}_allgather.json CASK_B200_DIST_SPARSE=0 -- --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0
step "done"
