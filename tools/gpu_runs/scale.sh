#!/bin/bash
# Scaling evidence on N GPUs of one box: weak (131072 envs per GPU, the bench default) and strong (BASELINE config 5 as written:
# 131072 envs in total, 131072 / N per GPU).  No per-step collective; NCCL all_gather of episode statistics only.
n=${1:-2}; tag=${2:-r2x}; which=${3:-both}
mkdir -p gpurun_out
run() {  # name envs_per_gpu
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --envs $2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_$1_${n}gpu.json 2> gpurun_out/${tag}_$1_${n}gpu.err
  echo "$1 rc=$?"; python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/${tag}_$1_${n}gpu.json').read().strip().splitlines()[-1])
  print('$1 n=$n envs/gpu=$2', {k:d[k] for k in ('value','ms_per_step','n_gpus','diverged','contacts_dropped')}, 'e2e', round(d['e2e']['value']))
except Exception as e:
  print('$1 failed', e); print(open('gpurun_out/${tag}_$1_${n}gpu.err').read()[-1500:])
PY
}
run weak 131072
if [ "$which" = both ]; then run strong $((131072 / n)); fi
