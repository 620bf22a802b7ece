#!/bin/bash
tag=${1:-r2v}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --profile-from-start off --clock-control none"
for spec in scene_kindyn:2 scene_gjk:2 scene_narrow_seq:2 scene_solve_kernel:2; do
  k=${spec%%:*}; c=${spec##*:}
  timeout 900 $NCU --set full --import-source on -k regex:$k -c $c -o /tmp/ncu/${tag}_$k python tools/ncu_target.py 131072 20 1 > gpurun_out/${tag}_ncu131072_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_summary.py /tmp/ncu/${tag}_$k.ncu-rep gpurun_out/${tag}_ncu131072_$k.txt > /dev/null 2>&1
  python tools/ncu_hotlines.py /tmp/ncu/${tag}_$k.ncu-rep $k so101_sim_b200/csrc/_obj/scene_kernel_f32.o 25 >> gpurun_out/${tag}_ncu131072_$k.txt 2>&1
done
rm -rf /tmp/ncu
head -40 gpurun_out/${tag}_ncu131072_scene_kindyn.txt | cut -c1-200
