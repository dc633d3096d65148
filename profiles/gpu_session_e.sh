#!/bin/bash
# One-GPU session E: CG with the compile-time keep switch; BiCGStab rate vs solve length with clocks sampled.
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
PY=python
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
step "CG on C4"
timeout 600 $PY bench.py --no-extra --no-cpu --no-probe --steps 20 --warmup 5 --soak 200 > $OUT/${TAG}_bench_cg.json 2> $OUT/${TAG}_bench_cg.err
line $OUT/${TAG}_bench_cg.json value cg.iters_per_s cg.us_per_iteration_marginal cg.loop_trips cg.roofline.stored_frac cg.roofline.frac
tail -2 $OUT/${TAG}_bench_cg.err
step "BiCGStab on C5: rate vs solve length"
for cap in 200 600 4000; do
  timeout 600 $PY bench.py --only-bicgstab --bicg-cap $cap --no-cg --no-cpu --no-probe --steps 20 --warmup 5 --soak 0 > $OUT/${TAG}_bicg_cap$cap.json 2>> $OUT/${TAG}_bicg.err
  line $OUT/${TAG}_bicg_cap$cap.json bicgstab.iters_per_s bicgstab.iterations bicgstab.rel_residual bicgstab.clocks bicgstab.roofline.stored_frac
done
CASK_B200_BICG_UNFUSED=1 timeout 600 $PY bench.py --only-bicgstab --bicg-cap 4000 --no-cg --no-cpu --no-probe --steps 20 --warmup 5 --soak 0 > $OUT/${TAG}_bicg_unfused_cap4000.json 2>> $OUT/${TAG}_bicg.err
line $OUT/${TAG}_bicg_unfused_cap4000.json bicgstab.iters_per_s bicgstab.iterations bicgstab.rel_residual bicgstab.clocks
tail -3 $OUT/${TAG}_bicg.err
step "done"
