#!/bin/bash
# One-GPU session D: streaming micro-benchmark (design input of the vector kernels), BiCGStab with the prefetched
# fused dot, its launch list.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash profiles/gpu_session_d.sh r2d'
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
PY=python
NCU="ncu --clock-control none"
step() { echo "== $1 ($(date +%T))"; }
line() { $PY - "$@" <<'PYEOF'
import json, sys
f, keys = sys.argv[1], sys.argv[2:]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
except Exception as e:
    print(f, "unreadable:", e); sys.exit(0)
out = {}
for k in keys:
    v = d
    for part in k.split("."):
        v = v.get(part) if isinstance(v, dict) else None
    out[k] = v
print(f.split("/")[-1], out)
PYEOF
}
step "streaming micro-benchmark"
timeout 300 profiles/micro/stream_bench | tee $OUT/${TAG}_stream_bench.txt

step "BiCGStab C5: fused (prefetched dot) vs unfused, 200 iterations"
timeout 600 $PY bench.py --only-bicgstab --bicg-cap 200 --no-cg --no-cpu --no-probe --steps 20 --warmup 5 --soak 0 > $OUT/${TAG}_bicg_fused.json 2> $OUT/${TAG}_bicg.err
line $OUT/${TAG}_bicg_fused.json bicgstab.iters_per_s bicgstab.iterations bicgstab.gpu_launches bicgstab.roofline.stored_frac
CASK_B200_BICG_UNFUSED=1 timeout 600 $PY bench.py --only-bicgstab --bicg-cap 200 --no-cg --no-cpu --no-probe --steps 20 --warmup 5 --soak 0 > $OUT/${TAG}_bicg_unfused.json 2>> $OUT/${TAG}_bicg.err
line $OUT/${TAG}_bicg_unfused.json bicgstab.iters_per_s bicgstab.iterations bicgstab.gpu_launches bicgstab.roofline.stored_frac
tail -3 $OUT/${TAG}_bicg.err

step "ncu launch list: BiCGStab C5, 12 iterations"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_bicg_launches.csv \
  $PY bench.py --only-bicgstab --bicg-cap 12 --no-cg --no-cpu --steps 3 --warmup 3 > $OUT/${TAG}_bicg_launches.log 2>&1
$PY profiles/summarize_launches.py $OUT/${TAG}_bicg_launches.csv > $OUT/${TAG}_bicg_launches_summary.md 2>> $OUT/${TAG}_bicg_launches.log
head -16 $OUT/${TAG}_bicg_launches_summary.md | cut -c1-200

step "C3 default + CG default"
timeout 600 $PY bench.py --no-cpu --no-probe --steps 20 --warmup 5 --soak 200 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
line $OUT/${TAG}_bench.json value rmat.ms_per_spmv rmat.kernel cg.iters_per_s bicgstab.iters_per_s bicgstab.iterations
step "done"
