#!/bin/bash
# per-function / per-line attribution of ONE launch of each main scene kernel at 131072 envs (and the tier kernels at 16384)
tag=${1:-r2z}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --profile-from-start off --clock-control none --set full --import-source on"
for k in scene_narrow_seq scene_solve_kernel scene_gjk; do
  timeout 900 $NCU -k regex:$k -c 1 -o /tmp/ncu/${tag}_$k python tools/ncu_target.py 131072 20 1 > gpurun_out/${tag}_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_hotlines.py /tmp/ncu/${tag}_$k.ncu-rep $k so101_sim_b200/csrc/_obj/scene_kernel_f32.o 30 > gpurun_out/${tag}_hotlines131072_$k.txt 2>&1
done
rm -rf /tmp/ncu; head -12 gpurun_out/${tag}_hotlines131072_scene_solve_kernel.txt | cut -c1-160
