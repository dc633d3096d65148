#!/bin/bash
# One-GPU session W: ILU application as one cooperative level kernel - tests, then pcg<ILU, unit lower> on the 64^3 twin.
OUT=gpurun_out; PY=python; TAG=r2w
timeout 900 $PY -m pytest tests/test_gpu_w_precond.py tests/test_gpu_z_host_tools.py -m gpu -q -rs > $OUT/${TAG}_pytest_precond.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_precond.log
timeout 600 $PY bench.py --steps 5 --warmup 3 --no-cpu --no-probe --no-cg --only-pcg-ilu --soak 0 > $OUT/${TAG}_bench_pcg_ilu.json 2> $OUT/${TAG}_bench_pcg_ilu.err
$PY -c "import json; print(json.dumps(json.loads(open('$OUT/${TAG}_bench_pcg_ilu.json').read().strip().splitlines()[-1]).get('pcg_ilu')))"
tail -2 $OUT/${TAG}_bench_pcg_ilu.err | cut -c1-300
