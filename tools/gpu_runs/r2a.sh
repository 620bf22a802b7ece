#!/bin/bash
# Round-2 session A: parity tests, smoke, arm precision experiment, bench (all legs), sanitizer runs.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python tools/exp_arm_precision.py > gpurun_out/${tag}_arm_precision.jsonl 2> gpurun_out/${tag}_arm_precision.err; echo "armexp rc=$?"
timeout 900 python bench.py --cpu-seconds 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
SO101_GRAPH=0 timeout 600 python bench.py --no-cpu-baseline --no-secondary --no-steady --steps 30 --warmup 5 > gpurun_out/${tag}_bench_nograph.json 2> gpurun_out/${tag}_bench_nograph.err; echo "bench nograph rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-secondary --no-steady --steps 30 --warmup 5 > gpurun_out/${tag}_bench_graph30.json 2> gpurun_out/${tag}_bench_graph30.err; echo "bench graph30 rc=$?"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_$tool.txt python tools/sanitize_probe.py 2 > gpurun_out/${tag}_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
done
tail -3 gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_arm_precision.jsonl; cut -c1-600 gpurun_out/${tag}_bench.json
for tool in memcheck racecheck; do tail -3 gpurun_out/${tag}_sanitizer_$tool.txt; done
