#!/bin/bash
tag=${1:-r2r}
mkdir -p gpurun_out
for c in 16 8 4 2; do
  SO101_SEQ_CTAS=$c timeout 900 python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_seq$c.json 2> gpurun_out/${tag}_seq$c.err
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/${tag}_seq$c.json').read().strip().splitlines()[-1])
  print('seq ctas/SM $c', round(d['value']), round(d['ms_per_step'],1), {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})
except Exception as e: print('seq $c failed', e)
PY
done
