#!/bin/bash
# the two-launch narrow phase: parity tests forced through it, then the bench with and without it
mkdir -p gpurun_out
SO101_NARROW_SPLIT=1 timeout 900 python -m pytest tests/test_scene_gpu.py tests/test_analytic_gpu.py tests/test_twoarm_gpu.py -m gpu -q > gpurun_out/r2aa_split_tests.log 2>&1; echo "split tests rc=$?"
tail -15 gpurun_out/r2aa_split_tests.log | cut -c1-300
for s in 32768 100000000; do
  for n in 131072 65536; do
    echo "== split_min=$s envs=$n"
    SO101_NARROW_SPLIT=$s timeout 600 python bench.py --envs $n --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('kernels_ms'))"
  done
done
