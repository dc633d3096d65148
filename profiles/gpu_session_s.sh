#!/bin/bash
# One-GPU session S: resident CTAs per SM of the merge-path gather kernel on C3 (more warps in flight against the
# 64 % LSU pipe utilisation), with and without hub clustering.
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
PY=python
step() { echo "== $1 ($(date +%T))"; }
rmat() {  # rmat <name> ENV=.. ENV=..
  local name=$1; shift
  local f=$OUT/${TAG}_rmat_${name}.json
  env "$@" timeout 300 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $f 2>> $OUT/${TAG}_rmat.err
  $PY - $f <<'PYEOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("rmat") or {}
    print(sys.argv[1].split("/")[-1], {k: r.get(k) for k in ("ms_per_spmv", "max_err_all_rows_rel_to_sum_abs", "error")})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PYEOF
}
step "gpu tests: gather path"
timeout 900 $PY -m pytest tests/test_gpu_spmv.py -m gpu -q -rs -k "merge or col_reorder" > $OUT/${TAG}_pytest_gather.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gather.log
step "C3: resident CTAs per SM x merge items"
rmat items7_ctas5 CASK_B200_MERGE_CTAS=0
rmat items7_ctas6 CASK_B200_MERGE_CTAS=6
rmat items7_ctas7 CASK_B200_MERGE_CTAS=7
rmat items5_ctas6 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=6
rmat items5_ctas7 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=7
rmat items5_ctas8 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=8
step "C3: the same with hub clustering"
rmat reorder_items7_ctas6 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_CTAS=6
rmat reorder_items5_ctas8 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=8
tail -2 $OUT/${TAG}_rmat.err | cut -c1-200
step "done"
