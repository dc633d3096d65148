#!/bin/bash
# One-GPU session T: automatic hub clustering of large skewed gather plans - tests, C3 with default options, ncu capture.
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
PY=python
NCU="ncu --clock-control none"
step() { echo "== $1 ($(date +%T))"; }
step "gpu tests: SpMV suite"
timeout 900 $PY -m pytest tests/test_gpu_spmv.py -m gpu -q -rs > $OUT/${TAG}_pytest_spmv.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_spmv.log
step "C3 with default options (automatic rule), and with clustering off"
for v in "auto" "off"; do
  f=$OUT/${TAG}_rmat_$v.json
  if [ $v = off ]; then export CASK_B200_COL_REORDER=0; else unset CASK_B200_COL_REORDER; fi
  CASK_B200_BENCH_DETAILS=$OUT/${TAG}_rmat_${v}_details.json timeout 300 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $f 2>> $OUT/${TAG}_rmat.err
  $PY -c "
import json; d=json.load(open('$OUT/${TAG}_rmat_${v}_details.json')); r=d['details']['rmat']
print('$v', r['ms_per_spmv'], r['max_err_all_rows_rel_to_sum_abs'], r['preprocess_s'], r['plan'], r['roofline']['frac'], r['roofline']['stored_frac'])"
done
unset CASK_B200_COL_REORDER
step "ncu --set full: C3 default (hub clustering, 5 items, 8 CTAs/SM)"
timeout 900 $NCU --set full --import-source on -k regex:"spmv_csr_merge_kernel|permute_x" -s 6 -c 2 -f -o $OUT/${TAG}_spmv_merge_rmat_auto \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe > $OUT/${TAG}_ncu_rmat.log 2>&1
cp profiles/roofline_traffic.json $OUT/${TAG}_roofline_traffic.json
ncu -i $OUT/${TAG}_spmv_merge_rmat_auto.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $OUT/${TAG}_roofline_traffic.json \
  --workload "C3: R-MAT scale 25, y = A x (hub-clustered columns, merge-path tiles of 5 items per thread)" --source profiles/${TAG}_spmv_merge_rmat_auto_ncu.md > $OUT/${TAG}_spmv_merge_rmat_auto_ncu.md
head -44 $OUT/${TAG}_spmv_merge_rmat_auto_ncu.md
step "done"
