#!/bin/bash
# One-GPU session U: a rank's stripe of the 8-rank C3 job (compact numbering) with the 5-item / 8-CTA kernel configuration.
OUT=gpurun_out; PY=python; TAG=r2u
for r in 0 7; do for cfg in "7 0" "5 8" "5 7"; do
  set -- $cfg
  f=$OUT/${TAG}_stripe_r${r}_items$1_ctas$2.json
  CASK_B200_MERGE_ITEMS=$1 CASK_B200_MERGE_CTAS=$2 timeout 300 $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe --rmat-stripe 8,$r > $f 2>> $OUT/${TAG}.err
  $PY -c "
import json; d=json.loads(open('$f').read().strip().splitlines()[-1])['rmat']; print('stripe $r items $1 ctas $2:', d.get('ms_per_spmv'), d.get('max_err_all_rows_rel_to_sum_abs'), d.get('error'))"
done; done
