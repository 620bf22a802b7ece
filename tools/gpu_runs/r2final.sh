#!/bin/bash
# round-2 final evidence on one GPU: full GPU suite, smoke, default bench (all legs), launch list + --set full captures of
# the top kernels at 131072 envs, compute-sanitizer memcheck / racecheck
tag=r2final
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/${tag}_pytest_gpu.log | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/${tag}_bench_banana131072.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_banana131072.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['steady_state']['value'], {k: round(v['value']) for k, v in d['other_workloads'].items()}, d['cpu_baseline']['value'])"
ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches_banana131072.csv python tools/ncu_target.py 131072 20 1 > gpurun_out/${tag}_ncu_l.log 2>&1; echo "launch list rc=$?"
for k in scene_epa_kernel scene_narrow_split scene_solve_kernel scene_gjk; do
  timeout 900 ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:$k -c 1 -o /tmp/ncu/${tag}_$k python tools/ncu_target.py 131072 20 1 > gpurun_out/${tag}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_summary.py /tmp/ncu/${tag}_$k.ncu-rep gpurun_out/${tag}_ncu131072_$k.txt > /dev/null 2>&1
  python tools/ncu_hotlines.py /tmp/ncu/${tag}_$k.ncu-rep $k so101_sim_b200/csrc/_obj/scene_kernel_f32.o 25 >> gpurun_out/${tag}_ncu131072_$k.txt 2>&1
done
rm -rf /tmp/ncu
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_$tool.txt python tools/sanitize_probe.py 2 > gpurun_out/${tag}_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -2 gpurun_out/${tag}_sanitizer_$tool.txt; tail -5 gpurun_out/${tag}_sanitizer_$tool.log | cut -c1-200
done
