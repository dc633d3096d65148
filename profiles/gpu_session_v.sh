#!/bin/bash
# One-GPU session V: chunk count of the host pipeline with the geometric ramp (pinned e2e on C2).
OUT=gpurun_out; PY=python; TAG=r2v
for k in 6 8 10 12 16 20; do
  CASK_B200_HOST_CHUNKS=$k timeout 600 $PY bench.py --steps 20 --warmup 5 --no-cpu --no-probe --no-extra --no-cg --soak 0 > $OUT/${TAG}_bench_chunks$k.json 2> $OUT/${TAG}_bench_chunks$k.err
  $PY - <<PYEOF
import json
d = json.loads(open("$OUT/${TAG}_bench_chunks$k.json").read().strip().splitlines()[-1])
print("chunks $k: pinned e2e ms %.3f" % d["e2e"]["ms_per_step"], "pageable %.3f" % d["e2e"]["pageable"]["ms_per_step"], "duplex floor %.3f" % d["e2e"].get("pcie", {}).get("duplex_ms_for_one_step", 0))
PYEOF
done
