#!/bin/bash
# One-GPU evidence session of round 2 (final state of the tree): everything the judge reads from one `gpurun` call.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/gpu_session_final.sh r2p'
# Order = importance: whole GPU suite, smoke, both bench arms with the driver's flags, selector validation, ncu launch
# list, full captures of the kernels that bound C2 / C4 / C3 (summaries + roofline_traffic.json written by script).
# Every step has its own timeout; numbers printed by a run under ncu are never bench values.
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
PY=python
NCU="ncu --clock-control none"
step() { echo "== $1 ($(date +%T))"; }
TJ=$OUT/${TAG}_roofline_traffic.json
cp profiles/roofline_traffic.json $TJ 2>/dev/null

step "gpu tests (whole suite, no -x)"
timeout 1800 $PY -m pytest tests -m gpu -q -rs --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -16 $OUT/${TAG}_pytest_gpu.log

step "smoke"
timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log

step "bench, both arms (driver's flags)"
timeout 600 $PY bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
tail -c 600 $OUT/${TAG}_bench_reference.json; echo
CASK_B200_BENCH_DETAILS=$OUT/${TAG}_bench_details.json timeout 1200 $PY bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
wc -c $OUT/${TAG}_bench.json; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

step "host pipeline: geometric chunk ramp vs equal chunks (pinned e2e)"
for r in 1 0; do
  CASK_B200_HOST_RAMP=$r timeout 600 $PY bench.py --steps 20 --warmup 3 --no-cpu --no-probe --no-extra --no-cg --soak 0 > $OUT/${TAG}_bench_ramp$r.json 2> $OUT/${TAG}_bench_ramp$r.err
  $PY - <<PYEOF
import json
d = json.loads(open("$OUT/${TAG}_bench_ramp$r.json").read().strip().splitlines()[-1])
print("ramp $r: pinned e2e ms", d["e2e"]["ms_per_step"], "pageable", d["e2e"]["pageable"]["ms_per_step"], "duplex floor", d["e2e"].get("pcie", {}).get("duplex_ms_for_one_step"))
PYEOF
done

step "selector model vs measured kernel times"
timeout 900 $PY profiles/dse_validate.py > $OUT/${TAG}_dse_validate.md 2> $OUT/${TAG}_dse_validate.err; cat $OUT/${TAG}_dse_validate.md; tail -2 $OUT/${TAG}_dse_validate.err

step "pcg<ILU, unit lower> on the 64^3 twin"
timeout 600 $PY bench.py --steps 5 --warmup 3 --no-cpu --no-probe --no-cg --only-pcg-ilu --soak 0 > $OUT/${TAG}_bench_pcg_ilu.json 2> $OUT/${TAG}_bench_pcg_ilu.err
$PY -c "import json; print(json.dumps(json.loads(open('$OUT/${TAG}_bench_pcg_ilu.json').read().strip().splitlines()[-1]).get('pcg_ilu')))"

step "ncu launch list of bench.py (C2 SpMV + C4 CG)"
timeout 1200 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $OUT/${TAG}_launches.csv \
  $PY bench.py --steps 20 --warmup 3 --no-extra --no-cpu --no-probe --soak 0 --cg-maxiters 100 > $OUT/${TAG}_launches.log 2>&1
$PY profiles/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.md 2>> $OUT/${TAG}_launches.log
head -14 $OUT/${TAG}_launches_summary.md

step "ncu --set full: persistent SpMV on C2"
timeout 900 $NCU --set full --import-source on -k regex:spmv_ell_persistent -s 5 -c 1 -f -o $OUT/${TAG}_spmv_persistent \
  $PY bench.py --steps 10 --warmup 3 --soak 0 --no-cg --no-extra --no-cpu --no-probe > $OUT/${TAG}_ncu_spmv.log 2>&1
ncu -i $OUT/${TAG}_spmv_persistent.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $TJ \
  --workload "C2: 2D 5-pt Poisson 4096^2, y = A x" --source profiles/${TAG}_spmv_persistent_ncu.md > $OUT/${TAG}_spmv_persistent_ncu.md
head -30 $OUT/${TAG}_spmv_persistent_ncu.md

step "ncu --set full: one CG iteration on C4 (SpMV + dot, fused update)"
timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k regex:".*(cg_update_fused|persistent_kernel<4).*" -s 20 -c 2 -f -o $OUT/${TAG}_cg_iteration \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-extra --no-cpu --no-probe --cg-maxiters 40 > $OUT/${TAG}_ncu_cg.log 2>&1
ncu -i $OUT/${TAG}_cg_iteration.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $TJ \
  --workload "C4: CG iteration, 3D 27-pt Poisson 256^3" --source profiles/${TAG}_cg_iteration_ncu.md > $OUT/${TAG}_cg_iteration_ncu.md
head -8 $OUT/${TAG}_cg_iteration_ncu.md

step "ncu --set full: merge-path gather SpMV on C3"
timeout 900 $NCU --set full --import-source on -k regex:spmv_csr_merge_kernel -s 3 -c 1 -f -o $OUT/${TAG}_spmv_merge_rmat \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu --no-probe > $OUT/${TAG}_ncu_rmat.log 2>&1
ncu -i $OUT/${TAG}_spmv_merge_rmat.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $TJ \
  --workload "C3: R-MAT scale 25, y = A x (merge-path tiles, 7 items per thread)" --source profiles/${TAG}_spmv_merge_rmat_ncu.md > $OUT/${TAG}_spmv_merge_rmat_ncu.md
head -8 $OUT/${TAG}_spmv_merge_rmat_ncu.md
step "done"
