#!/bin/bash
tag=${1:-r2y}
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_$tool.txt python tools/sanitize_probe.py 2 > gpurun_out/${tag}_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/${tag}_sanitizer_$tool.txt; tail -4 gpurun_out/${tag}_sanitizer_$tool.log | cut -c1-200
done
ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches_banana131072.csv python tools/ncu_target.py 131072 30 1 > gpurun_out/${tag}_ncu_l.log 2>&1; echo "launch list rc=$?"
