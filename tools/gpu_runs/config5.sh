#!/bin/bash
# BASELINE config 5: 131072 envs in total on N GPUs (strong scaling: 131072 / N per GPU), NCCL gather of episode statistics only.
n=${1:-1}; tag=${2:-r2s}
mkdir -p gpurun_out
per=$((131072 / n))
if [ "$n" = "1" ]; then
  timeout 1500 python bench.py --envs $per --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_config5_${n}gpu.json 2> gpurun_out/${tag}_config5_${n}gpu.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --envs $per --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_config5_${n}gpu.json 2> gpurun_out/${tag}_config5_${n}gpu.err
fi
echo "rc=$?"; tail -2 gpurun_out/${tag}_config5_${n}gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_config5_${n}gpu.json').read().strip().splitlines()[-1])
print('config5 n=$n envs/gpu=$per', {k:d[k] for k in ('value','ms_per_step','n_gpus','diverged','contacts_dropped')}, 'e2e', d['e2e']['value'])
PY
