#!/bin/bash
# 8-GPU check of the sparse exchange with hub-ordered segments: C3 only, default options.
OUT=gpurun_out; TAG=${1:-r2l8c}
CASK_B200_BENCH_DETAILS=$OUT/${TAG}_rmat_n8_details.json timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 8 --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $OUT/${TAG}_rmat_n8.json 2> $OUT/${TAG}_rmat_n8.err
python - <<PYEOF
import json
d = json.load(open("$OUT/${TAG}_rmat_n8_details.json")); r = d["details"]["rmat"]
print({k: r.get(k) for k in ("ms_per_spmv", "max_err_all_rows_rel_to_sum_abs", "preprocess_s", "nnz_share_max_rank", "col_reorder")}, r["plan"])
PYEOF
tail -c 300 $OUT/${TAG}_rmat_n8.err | grep -v Warning | tail -2
