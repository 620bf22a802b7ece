#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
SO101_EPA_REFILL=20 timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2ae_launches_131072_after20.csv python tools/ncu_target.py 131072 20 1 > gpurun_out/r2ae_ncu.log 2>&1; echo "launch list rc=$?"
for k in scene_epa_kernel; do
  SO101_EPA_REFILL=20 timeout 900 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:$k -c 1 -o /tmp/ncu/r2ae_$k python tools/ncu_target.py 131072 20 1 > gpurun_out/r2ae_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_summary.py /tmp/ncu/r2ae_$k.ncu-rep gpurun_out/r2ae_ncu131072_$k.txt > /dev/null 2>&1
  python tools/ncu_hotlines.py /tmp/ncu/r2ae_$k.ncu-rep $k so101_sim_b200/csrc/_obj/scene_kernel_f32.o 30 >> gpurun_out/r2ae_ncu131072_$k.txt 2>&1
done
SO101_EPA_REFILL=0 timeout 900 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:scene_narrow_split -c 1 -o /tmp/ncu/r2ae_split0 python tools/ncu_target.py 131072 20 1 > gpurun_out/r2ae_ncu_split0.log 2>&1; echo "ncu split0 rc=$?"
python tools/ncu_summary.py /tmp/ncu/r2ae_split0.ncu-rep gpurun_out/r2ae_ncu131072_split0.txt > /dev/null 2>&1
python tools/ncu_hotlines.py /tmp/ncu/r2ae_split0.ncu-rep scene_narrow_split so101_sim_b200/csrc/_obj/scene_kernel_f32.o 30 >> gpurun_out/r2ae_ncu131072_split0.txt 2>&1
rm -rf /tmp/ncu
