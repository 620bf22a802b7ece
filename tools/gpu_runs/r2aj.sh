#!/bin/bash
# manifold kernel: pairs from a cursor in warp-sized chunks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scene_gpu.py -m gpu -q -k "two_launch" 2>&1 | tail -3
run() { python bench.py --envs $1 --steps $2 --warmup $3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})"; }
echo "== 131072 x20"; run 131072 20 3
echo "== 131072 x100"; run 131072 100 10
