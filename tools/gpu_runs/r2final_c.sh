#!/bin/bash
# --set full captures of the solver's larger tiers and of the manifold kernel (pairs from a cursor) at 131072 envs
tag=r2final
mkdir -p gpurun_out /tmp/ncu
for spec in scene_solve_tier:2 scene_narrow_split:1; do
  k=${spec%%:*}; c=${spec##*:}
  timeout 900 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:$k -c $c -o /tmp/ncu/${tag}_$k python tools/ncu_target.py 131072 20 1 > gpurun_out/${tag}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_summary.py /tmp/ncu/${tag}_$k.ncu-rep gpurun_out/${tag}_ncu131072_$k.txt > /dev/null 2>&1
  python tools/ncu_hotlines.py /tmp/ncu/${tag}_$k.ncu-rep $k so101_sim_b200/csrc/_obj/scene_kernel_f32.o 25 >> gpurun_out/${tag}_ncu131072_$k.txt 2>&1
done
rm -rf /tmp/ncu
grep -E "Kernel Name|gpu__time_duration|dram__bytes|warps_active" gpurun_out/${tag}_ncu131072_scene_solve_tier.txt gpurun_out/${tag}_ncu131072_scene_narrow_split.txt | cut -c1-220
