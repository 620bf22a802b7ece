#!/bin/bash
# ncu launch list of bench.py itself (the timed steps of the default workload, graph replay: ncu profiles the graph's kernel nodes)
mkdir -p gpurun_out
timeout 700 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2final_launches_bench_py.csv python bench.py --steps 1 --warmup 20 --no-cpu-baseline --no-secondary --no-steady --profiler-range > gpurun_out/r2final_ncu_bench_py.log 2>&1; echo "rc=$?"
grep -c scene_ gpurun_out/r2final_launches_bench_py.csv; tail -2 gpurun_out/r2final_ncu_bench_py.log | cut -c1-300
