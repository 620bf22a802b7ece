#!/bin/bash
# One GPU session: parity tests, smoke, bench line, ncu launch list + full capture of the main kernels at the bench config.
# Usage (under gpurun): bash tools/gpu_runs/round.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "bench ref rc=$?"
timeout 300 python tools/profile_stages.py 16384 5 f32 30 > gpurun_out/${tag}_stages_random_regime.json 2> gpurun_out/${tag}_stages.err; echo "stages rc=$?"
# launch list of the same command as the bench (fewer steps): settle 50 steps x 166 launches (2 pipeline groups x 83) + warm-up 20 x 166 -> skip 11700, take 2 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11700 -c 340 --csv --log-file gpurun_out/${tag}_launches_banana16384.csv python bench.py --steps 4 --warmup 20 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu launches rc=$?"
for k in scene_narrow_seq scene_solve_kernel scene_solve_tier scene_gjk scene_broad scene_kindyn; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1500 -c 1 -o gpurun_out/${tag}_${k} python bench.py --steps 2 --warmup 25 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
tail -2 gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_smoke.log; cut -c1-400 gpurun_out/${tag}_bench.json; cut -c1-300 gpurun_out/${tag}_bench_reference.json
