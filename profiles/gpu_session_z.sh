#!/bin/bash
# One-GPU session Z: the two captures still missing from the record - the fused SpMV + dot of the CG loop on C4
# (spmv_ell_persistent_kernel<4,true,0>) and the three vector kernels of a BiCGStab iteration on C5.
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
PY=python
NCU="ncu --clock-control none"
TJ=$OUT/${TAG}_roofline_traffic.json
cp profiles/roofline_traffic.json $TJ
echo "== ncu --set full: SpMV + dot of the CG loop on C4"
timeout 900 $NCU --set full --import-source on -k regex:spmv_ell_persistent -s 40 -c 2 -f -o $OUT/${TAG}_cg_spmv \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-extra --no-cpu --no-probe --cg-maxiters 40 > $OUT/${TAG}_ncu_cg_spmv.log 2>&1
ncu -i $OUT/${TAG}_cg_spmv.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $TJ \
  --workload "C4: CG iteration, 3D 27-pt Poisson 256^3 (SpMV fused with p.Ap)" --source profiles/${TAG}_cg_spmv_ncu.md > $OUT/${TAG}_cg_spmv_ncu.md
head -34 $OUT/${TAG}_cg_spmv_ncu.md
echo "== ncu --set full: vector kernels of a BiCGStab iteration on C5"
timeout 900 $NCU --set full --import-source on -k regex:"bicg_(p|s|xr)_kernel" -s 6 -c 3 -f -o $OUT/${TAG}_bicg_vec \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-cg --only-bicgstab --no-cpu --no-probe --bicg-cap 12 > $OUT/${TAG}_ncu_bicg.log 2>&1
ncu -i $OUT/${TAG}_bicg_vec.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $TJ \
  --workload "C5: BiCGStab iteration, 3D 7-pt convection-diffusion 512^3" --source profiles/${TAG}_bicg_vec_ncu.md > $OUT/${TAG}_bicg_vec_ncu.md
head -8 $OUT/${TAG}_bicg_vec_ncu.md
tail -3 $OUT/${TAG}_ncu_bicg.log | cut -c1-300
