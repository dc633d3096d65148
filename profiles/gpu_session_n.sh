#!/bin/bash
# One-GPU session N: what each rank of an 8-rank C3 job does (stripe emulation with the sparse exchange's numbering),
# host pipeline chunk sweep, ILU-preconditioned CG with the level graph, new tests.
TAG=${1:-r2n}
OUT=gpurun_out
mkdir -p $OUT
PY=python
NCU="ncu --clock-control none"
step() { echo "== $1 ($(date +%T))"; }
step "gpu tests: gather options, preconditioners"
timeout 900 $PY -m pytest tests/test_gpu_spmv.py tests/test_gpu_w_precond.py -m gpu -q -rs -k "col_reorder or merge or pageable or ilu or pcg" > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log; tail -6 $OUT/${TAG}_pytest.log
step "C3 stripes of an 8-rank job, one at a time (ncu launch list: permute, merge, fix-up)"
for bal in items nnz; do for r in 0 3 7; do
  f=$OUT/${TAG}_stripe_${bal}_r$r
  CASK_B200_RMAT_BALANCE=$bal timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:"permute_x|spmv_csr_merge|spmv_ell" -c 40 --csv --log-file $f.csv \
    $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe --rmat-stripe 8,$r > $f.json 2> $f.err
  echo "balance=$bal stripe $r of 8:"; $PY profiles/summarize_launches.py $f.csv | head -6
  $PY - $f.json <<'PYEOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d.get("rmat") or {}
    print({k: r.get(k) for k in ("ms_per_spmv", "max_err_all_rows_rel_to_sum_abs", "nnz", "error")})
except Exception as e:
    print("unreadable", e)
PYEOF
done; done
step "host pipeline chunks (pinned e2e)"
for k in 16 32 64; do
  CASK_B200_HOST_CHUNKS=$k timeout 600 $PY bench.py --steps 20 --warmup 3 --no-cpu --no-probe --no-extra --no-cg --soak 0 > $OUT/${TAG}_bench_chunks$k.json 2> $OUT/${TAG}_bench_chunks$k.err
  $PY - <<PYEOF
import json
d = json.loads(open("$OUT/${TAG}_bench_chunks$k.json").read().strip().splitlines()[-1])
print("chunks $k: pinned e2e ms", d["e2e"]["ms_per_step"], "pageable", d["e2e"]["pageable"]["ms_per_step"], "duplex floor", d["e2e"].get("pcie", {}).get("duplex_ms_for_one_step"))
PYEOF
done
step "pcg<ILU, unit lower> on the 64^3 twin: level graph vs launches"
timeout 600 $PY bench.py --steps 5 --warmup 3 --no-cpu --no-probe --no-cg --only-pcg-ilu --soak 0 > $OUT/${TAG}_bench_pcg_ilu.json 2> $OUT/${TAG}_bench_pcg_ilu.err
$PY - <<PYEOF
import json
d = json.loads(open("$OUT/${TAG}_bench_pcg_ilu.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("pcg_ilu"), indent=1))
PYEOF
tail -2 $OUT/${TAG}_bench_pcg_ilu.err
step "done"
