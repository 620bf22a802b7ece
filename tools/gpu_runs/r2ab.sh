#!/bin/bash
# where a control step goes at 131072 envs late in a random-action rollout (the regime the 100-step bench averages over)
mkdir -p gpurun_out
for pre in 20 110; do
  timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2ab_launches_131072_after${pre}.csv python tools/ncu_target.py 131072 $pre 1 > gpurun_out/r2ab_ncu_after${pre}.log 2>&1; echo "launch list after $pre rc=$?"
done
for pre in 20 110; do
  timeout 600 python tools/profile_stages.py 32768 5 f32 $pre > gpurun_out/r2ab_stages_32768_after${pre}.json 2> gpurun_out/r2ab_stages_err.log; echo "stages after $pre rc=$?"
done
