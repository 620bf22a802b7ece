#!/bin/bash
# EPA state-machine kernel: parity tests forced through the two-launch narrow phase, then the bench over refill thresholds
mkdir -p gpurun_out
SO101_NARROW_SPLIT=1 timeout 900 python -m pytest tests/test_scene_gpu.py tests/test_analytic_gpu.py tests/test_twoarm_gpu.py tests/test_placement_gpu.py -m gpu -q > gpurun_out/r2ad_split_tests.log 2>&1; echo "split tests rc=$?"
tail -6 gpurun_out/r2ad_split_tests.log | cut -c1-300
run() { python bench.py --envs $1 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
for r in 0 4 12 20 28; do echo "== 131072 refill=$r"; SO101_EPA_REFILL=$r run 131072; done
for r in 0 12; do echo "== 16384 split forced refill=$r"; SO101_NARROW_SPLIT=1 SO101_EPA_REFILL=$r run 16384; done
echo "== 16384 fused"; run 16384
