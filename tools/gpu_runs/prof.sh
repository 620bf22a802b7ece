#!/bin/bash
# GPU session: stage profile + ncu full capture of the scene kernel.  Usage (under gpurun): bash tools/gpu_runs/prof.sh <tag>
tag=${1:-r01b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/profile_stages.py 2048 10 f32 > gpurun_out/${tag}_stages_f32_2048.json 2> gpurun_out/${tag}_stages.err; echo "stages rc=$?"
timeout 300 python tools/profile_stages.py 16384 5 f32 > gpurun_out/${tag}_stages_f32_16384.json 2>> gpurun_out/${tag}_stages.err; echo "stages rc=$?"
timeout 300 python tools/profile_stages.py 1024 5 f64 > gpurun_out/${tag}_stages_f64_1024.json 2>> gpurun_out/${tag}_stages.err; echo "stages rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scene_step -s 8 -c 1 -o gpurun_out/${tag}_scene python bench.py --workload banana16384 --envs 1184 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_scene.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_stages_*.json; tail -3 gpurun_out/${tag}_stages.err
