#!/bin/bash
# One-GPU session K: hub clustering of the gather path (col_reorder) on C3 - tests, A/B, phase diagnostics, ncu capture.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/gpu_session_k.sh r2k'
TAG=${1:-r2k}
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
rmat() {  # rmat <name> ENV=.. ENV=..
  local name=$1; shift
  local f=$OUT/${TAG}_rmat_${name}.json
  env "$@" timeout 300 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $f 2>> $OUT/${TAG}_rmat.err
  line $f rmat.ms_per_spmv rmat.max_err_all_rows_rel_to_sum_abs rmat.roofline.frac rmat.col_reorder rmat.preprocess_s
}

step "gpu tests: gather path"
timeout 900 $PY -m pytest tests/test_gpu_spmv.py -m gpu -q -rs -k "merge or col_reorder" > $OUT/${TAG}_pytest_gather.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gather.log
tail -8 $OUT/${TAG}_pytest_gather.log

step "C3: hub clustering A/B"
rmat reorder0 CASK_B200_COL_REORDER=0
rmat reorder1 CASK_B200_COL_REORDER=1
rmat reorder1_xcg1 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_XCG=1
rmat reorder1_items5 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=5
rmat reorder1_items11 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=11
step "C3: phase diagnostics (y is NOT computed in these runs: the error column is meaningless)"
rmat reorder0_streamonly CASK_B200_COL_REORDER=0 CASK_B200_MERGE_DIAG=2
rmat reorder0_nogather CASK_B200_COL_REORDER=0 CASK_B200_MERGE_DIAG=4
rmat reorder0_streamonly_nogather CASK_B200_COL_REORDER=0 CASK_B200_MERGE_DIAG=6
rmat reorder1_streamonly CASK_B200_COL_REORDER=1 CASK_B200_MERGE_DIAG=2
tail -3 $OUT/${TAG}_rmat.err

step "ncu launch list: C3 with hub clustering (permute + merge + fix-up shares)"
CASK_B200_COL_REORDER=1 timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:"permute_x|spmv_csr_merge" -c 60 --csv --log-file $OUT/${TAG}_rmat_launches.csv \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe > $OUT/${TAG}_rmat_launches.log 2>&1
$PY profiles/summarize_launches.py $OUT/${TAG}_rmat_launches.csv > $OUT/${TAG}_rmat_launches_summary.md 2>&1; head -12 $OUT/${TAG}_rmat_launches_summary.md

step "ncu --set full: merge-path gather SpMV on C3 with hub clustering"
CASK_B200_COL_REORDER=1 timeout 900 $NCU --set full --import-source on -k regex:"spmv_csr_merge_kernel|permute_x" -s 6 -c 2 -f -o $OUT/${TAG}_spmv_merge_rmat_reorder \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe > $OUT/${TAG}_ncu_rmat.log 2>&1
ncu -i $OUT/${TAG}_spmv_merge_rmat_reorder.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py > $OUT/${TAG}_spmv_merge_rmat_reorder_ncu.md
head -60 $OUT/${TAG}_spmv_merge_rmat_reorder_ncu.md
step "done"
