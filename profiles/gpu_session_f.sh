#!/bin/bash
# One-GPU session F: micro-benchmark with the fused replica + L2 set-aside variants, then the default bench.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
PY=python
step() { echo "== $1 ($(date +%T))"; }
step "streaming micro-benchmark"
timeout 300 profiles/micro/stream_bench | tee $OUT/${TAG}_stream_bench.txt
step "default bench"
CASK_B200_BENCH_DETAILS=$OUT/${TAG}_bench_details.json timeout 900 $PY bench.py --steps 20 --warmup 5 --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json | cut -c1-4500
tail -2 $OUT/${TAG}_bench.err
step "done"
