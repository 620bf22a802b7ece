#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --cpu-seconds 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/${tag}_smoke.log; tail -3 gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','graph_launches','launches_per_step','diverged','contacts_dropped')})
print('steady', d.get('steady_state')); print('roofline', d['roofline']['kernel'], d['roofline']['us_per_launch'], d['roofline']['frac'])
print({k:(round(v['us_per_launch'],1), round(v['share_of_kernel_time'],3)) for k,v in d['kernels'].items()})
print('cpu', d.get('cpu_baseline'))
PY
