#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|f32 scene|final" gpurun_out/${tag}_pytest_gpu.log | cut -c1-500
for g in 1 2 3 4 6; do
  SO101_GROUPS=$g timeout 300 python bench.py --no-cpu-baseline --no-secondary --no-steady --steps 30 --warmup 10 > gpurun_out/${tag}_groups$g.json 2> gpurun_out/${tag}_groups$g.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_groups$g.json').read().strip().splitlines()[-1])
print('groups $g', round(d['value']), round(d['ms_per_step'],2), {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})
PY
done
