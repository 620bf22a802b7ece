#!/bin/bash
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python tools/exp_arm_precision.py > gpurun_out/${tag}_arm_precision.jsonl 2> gpurun_out/${tag}_arm_precision.err; echo "armexp rc=$?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
grep -E "passed|failed|FAILED|rest penetration|sliding|f32 arm|f32 scene|final" gpurun_out/${tag}_pytest_gpu.log | cut -c1-700
tail -2 gpurun_out/${tag}_smoke.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','graph_launches','launches_per_step','diverged','contacts_dropped')})
print('steady', d.get('steady_state')); print('roofline', d['roofline']['kernel'], d['roofline']['us_per_launch'], d['roofline']['frac'])
print({k:(round(v['us_per_launch'],1), round(v['share_of_kernel_time'],3)) for k,v in d['kernels'].items()})
print('arm', d.get('other_workloads'))
PY
