#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
for conn in 8 32; do for gr in 1 0; do for g in 2 4 8; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn SO101_GRAPH=$gr SO101_GROUPS=$g timeout 300 python bench.py --no-cpu-baseline --no-secondary --no-steady --steps 30 --warmup 10 > gpurun_out/${tag}_c${conn}_gr${gr}_g$g.json 2> gpurun_out/${tag}_c${conn}_gr${gr}_g$g.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_c${conn}_gr${gr}_g$g.json').read().strip().splitlines()[-1])
print('conn $conn graph $gr groups $g', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
PY
done; done; done
