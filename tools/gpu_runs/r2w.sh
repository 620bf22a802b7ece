#!/bin/bash
tag=${1:-r2w}
for nc in 1 0; do for e in 16384 131072; do
SO101_NO_CARVEOUT=$nc python bench.py --workload banana16384 --envs $e --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no_carveout=$nc', d['config']['envs_per_gpu'], round(d['value']), round(d['ms_per_step'],1), {k:round(v['us_per_launch']) for k,v in d['kernels'].items()})"
done; done
