#!/bin/bash
# Short one-GPU session: the merge-path gather kernel (C3) and the CG pair on C4 under ncu.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/gpu_session_b.sh r2b'
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
PY=python
NCU="ncu --clock-control none"
step() { echo "== $1 ($(date +%T))"; }

step "gpu tests: SpMV suites + host tools"
timeout 900 $PY -m pytest tests/test_gpu_spmv.py tests/test_gpu_z_host_tools.py tests/test_gpu_solvers.py -m gpu -q -rs > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log

step "C3: merge-path tiles (default) vs row-group items"
for v in "CASK_B200_CSR_KERNEL=1" "CASK_B200_CSR_KERNEL=0"; do
  f=$OUT/${TAG}_rmat_$(echo $v | tr ' =' '__').json
  env $v timeout 600 $PY bench.py --only-rmat --no-cg --no-cpu --no-probe --steps 20 --warmup 3 --soak 0 > $f 2>> $OUT/${TAG}_rmat.err
  $PY -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1])['rmat']; print('$v', d)" 2>&1 | tail -2
done
tail -5 $OUT/${TAG}_rmat.err

step "ncu --set full: merge-path gather SpMV on C3 (R-MAT)"
timeout 1500 $NCU --set full --import-source on -k regex:spmv_csr_merge_kernel -s 3 -c 1 -f -o $OUT/${TAG}_spmv_merge_rmat \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-rmat --no-cpu > $OUT/${TAG}_ncu_rmat.log 2>&1
ncu -i $OUT/${TAG}_spmv_merge_rmat.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py > $OUT/${TAG}_spmv_merge_rmat_ncu.md
head -40 $OUT/${TAG}_spmv_merge_rmat_ncu.md

step "ncu --set full: one CG iteration on C4 (SpMV + dot, fused update)"
timeout 1200 $NCU --set full --import-source on --kernel-name-base demangled -k regex:".*(cg_update_fused|persistent_kernel<4).*" -s 20 -c 2 -f -o $OUT/${TAG}_cg_iteration \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-extra --no-cpu --cg-maxiters 40 > $OUT/${TAG}_ncu_cg.log 2>&1
ncu -i $OUT/${TAG}_cg_iteration.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py > $OUT/${TAG}_cg_iteration_ncu.md
head -60 $OUT/${TAG}_cg_iteration_ncu.md
tail -3 $OUT/${TAG}_ncu_cg.log
step "done"
