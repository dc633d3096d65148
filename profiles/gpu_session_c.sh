#!/bin/bash
# One-GPU session C: whole GPU suite, default bench, A/Bs of this round's kernel changes, captures.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/gpu_session_c.sh r2c'
TAG=${1:-r2c}
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

step "gpu tests (whole suite, no -x)"
timeout 1800 $PY -m pytest tests -m gpu -q -rs --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -14 $OUT/${TAG}_pytest_gpu.log

step "bench, both arms (driver's flags)"
timeout 600 $PY bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
line $OUT/${TAG}_bench_reference.json value ms_per_step cpu_baseline.cores other_cpu.gflops
CASK_B200_BENCH_DETAILS=$OUT/${TAG}_bench_details.json timeout 900 $PY bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
wc -c $OUT/${TAG}_bench.json; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err

step "A/B: vector kernels with one contiguous chunk per CTA (round-1 map) - CG on C4"
CASK_B200_VEC_RR=0 timeout 600 $PY bench.py --no-extra --no-cpu --no-probe --steps 20 --warmup 5 --soak 200 > $OUT/${TAG}_bench_vecrr0.json 2> $OUT/${TAG}_bench_vecrr0.err
line $OUT/${TAG}_bench_vecrr0.json cg.iters_per_s cg.us_per_iteration_marginal cg.roofline.stored_frac
line $OUT/${TAG}_bench.json cg.iters_per_s cg.us_per_iteration_marginal cg.roofline.stored_frac

step "A/B: BiCGStab on C5, unfused kernels"
CASK_B200_BICG_UNFUSED=1 timeout 600 $PY bench.py --only-bicgstab --no-cg --no-cpu --no-probe --steps 20 --warmup 5 --soak 0 > $OUT/${TAG}_bench_bicg_unfused.json 2> $OUT/${TAG}_bench_bicg_unfused.err
line $OUT/${TAG}_bench_bicg_unfused.json bicgstab.iters_per_s bicgstab.iterations bicgstab.rel_residual bicgstab.gpu_launches bicgstab.roofline.stored_frac
CASK_B200_BICG_UNFUSED=1 CASK_B200_VEC_RR=0 timeout 600 $PY bench.py --only-bicgstab --no-cg --no-cpu --no-probe --steps 20 --warmup 5 --soak 0 > $OUT/${TAG}_bench_bicg_round1.json 2> $OUT/${TAG}_bench_bicg_round1.err
line $OUT/${TAG}_bench_bicg_round1.json bicgstab.iters_per_s bicgstab.iterations bicgstab.rel_residual bicgstab.gpu_launches
line $OUT/${TAG}_bench.json bicgstab.iters_per_s bicgstab.iterations bicgstab.rel_residual bicgstab.gpu_launches bicgstab.roofline.stored_frac

step "C3 sweep: merge items per thread x gathers past L1"
for it in 5 7 11 17; do for cg in 0 1; do
  f=$OUT/${TAG}_rmat_items${it}_xcg${cg}.json
  CASK_B200_CSR_KERNEL=1 CASK_B200_MERGE_ITEMS=$it CASK_B200_MERGE_XCG=$cg timeout 300 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $f 2>> $OUT/${TAG}_rmat.err
  line $f rmat.ms_per_spmv rmat.max_err_all_rows_rel_to_sum_abs rmat.roofline.frac
done; done
CASK_B200_CSR_KERNEL=0 timeout 300 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $OUT/${TAG}_rmat_rowgroups.json 2>> $OUT/${TAG}_rmat.err
line $OUT/${TAG}_rmat_rowgroups.json rmat.ms_per_spmv rmat.kernel
tail -3 $OUT/${TAG}_rmat.err

step "ncu --set full: merge-path gather SpMV on C3 (default configuration)"
timeout 900 $NCU --set full --import-source on -k regex:spmv_csr_merge_kernel -s 3 -c 1 -f -o $OUT/${TAG}_spmv_merge_rmat \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu > $OUT/${TAG}_ncu_rmat.log 2>&1
ncu -i $OUT/${TAG}_spmv_merge_rmat.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py > $OUT/${TAG}_spmv_merge_rmat_ncu.md
head -34 $OUT/${TAG}_spmv_merge_rmat_ncu.md

step "ncu --set full: one CG iteration on C4 (SpMV + dot, fused update)"
timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k regex:".*(cg_update_fused|persistent_kernel<4).*" -s 20 -c 2 -f -o $OUT/${TAG}_cg_iteration \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-extra --no-cpu --cg-maxiters 40 > $OUT/${TAG}_ncu_cg.log 2>&1
ncu -i $OUT/${TAG}_cg_iteration.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py > $OUT/${TAG}_cg_iteration_ncu.md
head -30 $OUT/${TAG}_cg_iteration_ncu.md
step "done"
