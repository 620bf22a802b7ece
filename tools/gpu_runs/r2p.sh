#!/bin/bash
tag=${1:-r2p}
mkdir -p gpurun_out
for cfg in "131072 1" "131072 2" "131072 4" "131072 8" "262144 2" "262144 4" "65536 2" "32768 2"; do
  set -- $cfg
  SO101_GROUPS=$2 timeout 900 python bench.py --envs $1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_e$1_g$2.json 2> gpurun_out/${tag}_e$1_g$2.err
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/${tag}_e$1_g$2.json').read().strip().splitlines()[-1])
  print('envs $1 groups $2', round(d['value']), round(d['ms_per_step'],1), {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})
except Exception as e: print('envs $1 groups $2 failed', e)
PY
done
