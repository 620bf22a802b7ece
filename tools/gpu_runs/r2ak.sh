#!/bin/bash
# inscribed-circle rejection in the large-slab extremes: parity suite + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ak_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ak_pytest_gpu.log | cut -c1-300
run() { python bench.py --envs $1 --steps $2 --warmup $3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})"; }
echo "== 16384 x20"; run 16384 20 3
echo "== 131072 x100"; run 131072 100 10
