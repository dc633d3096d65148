#!/bin/bash
# One-GPU session U2: would hub clustering INSIDE a rank's compact numbering pay in the sharded C3 job?  Stripes of the
# 8- and 2-rank jobs run alone with compact numbering (2, what dist.cu builds) and with hub clustering (1) + 5 x 8 tiles.
OUT=gpurun_out; PY=python; TAG=r2u2
for w in 8 2; do for r in 0 $((w-1)); do for cfg in "2 7 0" "1 5 8" "1 7 0"; do
  set -- $cfg
  f=$OUT/${TAG}_w${w}_r${r}_reorder$1_items$2_ctas$3.json
  CASK_B200_STRIPE_REORDER=$1 CASK_B200_MERGE_ITEMS=$2 CASK_B200_MERGE_CTAS=$3 timeout 300 $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe --rmat-stripe $w,$r > $f 2>> $OUT/${TAG}.err
  $PY -c "
import json; d=json.loads(open('$f').read().strip().splitlines()[-1])['rmat']; print('world $w stripe $r reorder $1 items $2 ctas $3:', d.get('ms_per_spmv'), d.get('max_err_all_rows_rel_to_sum_abs'), d.get('error'))"
done; done; done
