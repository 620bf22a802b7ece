#!/bin/bash
# cap walk in feature_seq: full GPU suite, then the bench at three sizes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ac_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2ac_pytest_gpu.log | cut -c1-300
for n in 16384 131072; do
  echo "== envs=$n"
  timeout 600 python bench.py --envs $n --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 600 python bench.py --envs 131072 --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('100 steps', d['value'], d['ms_per_step'], d['e2e']['value'])"
