#!/bin/bash
# ncu evidence for the round: launch list of 2 steady-state control steps + --set full captures of every scene kernel on launches
# that have work (the profiled region starts after 30 random-action steps).  Reports are summarised on the box (text only comes back).
tag=${1:-r2n}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --profile-from-start off --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches_banana16384.csv python tools/ncu_target.py 16384 30 2 > gpurun_out/${tag}_ncu_l.log 2>&1; echo "launch list rc=$?"
for spec in scene_narrow_seq:2 scene_solve_kernel:2 scene_solve_tier:8 scene_gjk:2 scene_broad:2 scene_kindyn:2; do
  k=${spec%%:*}; c=${spec##*:}
  timeout 900 $NCU --set full --import-source on -k regex:$k -c $c -o /tmp/ncu/${tag}_$k python tools/ncu_target.py 16384 30 1 > gpurun_out/${tag}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_summary.py /tmp/ncu/${tag}_$k.ncu-rep gpurun_out/${tag}_ncu_$k.txt > /dev/null 2>&1
  # per-function / per-line attribution of the LONGEST launch of the capture (tier kernels: one per tier)
  python - <<PY >> gpurun_out/${tag}_ncu_$k.txt 2>&1
import csv, subprocess
rep = '/tmp/ncu/${tag}_$k.ncu-rep'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; iname, idur = hdr.index('Kernel Name'), hdr.index('gpu__time_duration.sum')
best = {}
for n, r in enumerate(rows[2:]):
  key = r[iname][:80]
  d = float(r[idur].replace(',', ''))
  if key not in best or d > best[key][1]: best[key] = (n, d)
for key, (n, d) in best.items():
  print(f'\\n=== hot lines of launch {n} ({d} {rows[1][idur]}) of {key}')
  sub = f'/tmp/ncu/one_{n}.ncu-rep'
  subprocess.run(['ncu', '-i', rep, '--launch-skip', str(n), '--launch-count', '1', '-o', sub[:-8], '-f'], capture_output=True, text=True)
  import os
  src = sub if os.path.exists(sub) else rep
  kn = '$k'
  if 'tier' in kn: kn = 'scene_solve_tier_kernelIfLi64' if 'Li64' in key or '64, 80' in key or '<float, 64' in key else 'scene_solve_tier_kernelIfLi128'
  print(subprocess.run(['python', 'tools/ncu_hotlines.py', src, kn, 'so101_sim_b200/csrc/_obj/scene_kernel_f32.o', '28'], capture_output=True, text=True).stdout[-6000:])
PY
done
rm -rf /tmp/ncu
ls -la gpurun_out/${tag}_*; head -c 1500 gpurun_out/${tag}_ncu_scene_solve_tier.txt
