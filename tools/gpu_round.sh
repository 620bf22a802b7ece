#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list.  Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/${tag}_bench_arm4096.json 2> gpurun_out/${tag}_bench_arm4096.err; echo "bench arm rc=$?"
timeout 600 python bench.py --workload banana16384 --envs 2048 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_banana2048.json 2> gpurun_out/${tag}_bench_banana2048.err; echo "bench banana2048 rc=$?"
timeout 900 python bench.py --workload banana16384 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_banana16384.json 2> gpurun_out/${tag}_bench_banana16384.err; echo "bench banana16384 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_arm.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_arm.log 2>&1; echo "ncu arm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_banana.csv python bench.py --workload banana16384 --envs 2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_banana.log 2>&1; echo "ncu banana rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_bench_*.json
