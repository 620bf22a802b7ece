#!/bin/bash
tag=${1:-r2q}
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python bench.py --cpu-seconds 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
echo "bench wall seconds: $SECONDS"; python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','graph_launches','launches_per_step','diverged','contacts_dropped','dropped_by_buffer_since_load')})
print('config', d['config']['workload'], d['config']['envs_per_gpu'])
print('steady', d.get('steady_state')); print('roofline', {k:d['roofline'][k] for k in ('kernel','us_per_launch','frac','achieved','traffic','whole_step_frac')})
print({k:(round(v['us_per_launch'],1), round(v['share_of_kernel_time'],3)) for k,v in d['kernels'].items()})
print('cpu', d.get('cpu_baseline'))
for k,v in d['other_workloads'].items(): print(k, round(v['value']), round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value']))
PY
