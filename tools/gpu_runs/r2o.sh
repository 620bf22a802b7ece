#!/bin/bash
tag=${1:-r2o}
mkdir -p gpurun_out
for wl in banana16384 handover8192; do
timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-secondary --no-steady --steps 100 --warmup 10 > gpurun_out/${tag}_$wl.json 2> gpurun_out/${tag}_$wl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_$wl.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']), d['ms_per_step'], 'dropped', d['contacts_dropped'], d['dropped_by_buffer_since_load'], 'diverged', d['diverged'])
PY
done
