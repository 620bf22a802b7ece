#!/bin/bash
tag=r2final
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/${tag}_bench_banana131072.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_banana131072.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['steady_state']['value'], {k: round(v['value']) for k, v in d['other_workloads'].items()}, d['cpu_baseline']['value']); print(d['roofline']); print({k:(round(v['share_of_kernel_time'],3), round(v['us_per_launch'])) for k,v in d['kernels'].items()})"
