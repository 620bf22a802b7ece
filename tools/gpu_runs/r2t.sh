#!/bin/bash
tag=${1:-r2t}
mkdir -p gpurun_out
for v in base minctas16; do for c in 16 24; do
  lib=""; [ $v != base ] && lib="variants_tmp/libso101_$v.so"
  SO101_B200_LIB=$lib SO101_SEQ_CTAS=$c timeout 900 python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_${v}_$c.json 2> gpurun_out/${tag}_${v}_$c.err
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/${tag}_${v}_$c.json').read().strip().splitlines()[-1])
  print('$v grid cap $c', round(d['value']), round(d['ms_per_step'],1), {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})
except Exception as e: print('$v $c failed', e)
PY
done; done
