#!/bin/bash
# One-GPU session M: host-buffer path with pageable vectors (library-staged), then the default bench.
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
PY=python
step() { echo "== $1 ($(date +%T))"; }
step "gpu tests: SpMV suite"
timeout 900 $PY -m pytest tests/test_gpu_spmv.py tests/test_gpu_harness.py -m gpu -q -rs -x > $OUT/${TAG}_pytest_spmv.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_spmv.log
tail -12 $OUT/${TAG}_pytest_spmv.log
step "default bench (no probes, no extra solvers)"
CASK_B200_BENCH_DETAILS=$OUT/${TAG}_bench_details.json timeout 900 $PY bench.py --steps 20 --warmup 5 --no-cpu --no-probe --no-extra --no-cg > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
$PY - <<PYEOF
import json
d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"]); print(json.dumps(d["e2e"], indent=1))
PYEOF
tail -3 $OUT/${TAG}_bench.err
for t in 2 4 12; do
step "pageable e2e with $t host copy threads"
CASK_B200_HOST_COPY_THREADS=$t timeout 600 $PY bench.py --steps 10 --warmup 3 --no-cpu --no-probe --no-extra --no-cg --soak 0 > $OUT/${TAG}_bench_t$t.json 2> $OUT/${TAG}_bench_t$t.err
$PY - <<PYEOF
import json
d = json.loads(open("$OUT/${TAG}_bench_t$t.json").read().strip().splitlines()[-1])
print(d["e2e"].get("pageable"))
PYEOF
done
step "done"
