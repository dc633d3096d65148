TAG=r2z; OUT=gpurun_out; PY=python
TJ=$OUT/${TAG}_roofline_traffic.json
timeout 900 ncu --clock-control none --set full --import-source on -k regex:spmv_ell_persistent -s 262 -c 2 -f -o $OUT/${TAG}_cg_spmv \
  $PY bench.py --steps 3 --warmup 3 --soak 0 --no-extra --no-cpu --no-probe --cg-maxiters 40 > $OUT/${TAG}_ncu_cg_spmv.log 2>&1
ncu -i $OUT/${TAG}_cg_spmv.ncu-rep --page raw --csv 2>/dev/null | $PY profiles/summarize_ncu.py --traffic-json $TJ \
  --workload "C4: CG iteration, 3D 27-pt Poisson 256^3 (SpMV fused with p.Ap)" --source profiles/${TAG}_cg_spmv_ncu.md > $OUT/${TAG}_cg_spmv_ncu.md
head -32 $OUT/${TAG}_cg_spmv_ncu.md
