#!/bin/bash
tag=${1:-r01e}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scene_narrow -s 300 -c 1 -o gpurun_out/${tag}_narrow python bench.py --workload banana16384 --envs 4096 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_narrow.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scene_solve -s 300 -c 1 -o gpurun_out/${tag}_solve python bench.py --workload banana16384 --envs 4096 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_solve.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 60 --csv --log-file gpurun_out/${tag}_launches_banana.csv python bench.py --workload banana16384 --envs 16384 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/${tag}_ncu_narrow.log
