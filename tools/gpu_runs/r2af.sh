#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2af_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2af_pytest_gpu.log | cut -c1-300
bash tools/gpu_runs/bench3.sh
for c in 12 14; do echo "== solve CTAs/SM $c"; SO101_SOLVE_CTAS=$c python bench.py --envs 131072 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; done
