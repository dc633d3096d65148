TAG=r2s
OUT=gpurun_out
PY=python
rmat() {
  local name=$1; shift
  local f=$OUT/${TAG}_rmat_${name}.json
  env "$@" timeout 300 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $f 2>> $OUT/${TAG}_rmat.err
  $PY - $f <<'PYEOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("rmat") or {}
    print(sys.argv[1].split("/")[-1], {k: r.get(k) for k in ("ms_per_spmv", "max_err_all_rows_rel_to_sum_abs", "preprocess_s", "error")})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PYEOF
}
rmat reorder_items5_ctas5 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=0
rmat reorder_items5_ctas6 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=6
rmat reorder_items5_ctas7 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=7
rmat reorder_items5_ctas8_b CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=5 CASK_B200_MERGE_CTAS=8
rmat reorder_items7_ctas5 CASK_B200_COL_REORDER=1 CASK_B200_MERGE_ITEMS=7 CASK_B200_MERGE_CTAS=0
