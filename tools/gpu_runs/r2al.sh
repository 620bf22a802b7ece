#!/bin/bash
# solver code-size experiments: parity subset + bench late-rollout
timeout 900 python -m pytest tests/test_scene_gpu.py tests/test_analytic_gpu.py tests/test_twoarm_gpu.py -m gpu -q -x 2>&1 | tail -2
python bench.py --envs 131072 --steps 100 --warmup 10 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})"
