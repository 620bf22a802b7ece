#!/bin/bash
# strong-scaling leg alone (BASELINE config 5 as written: 131072 envs in total)
n=${1:-8}; tag=${2:-r2final}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --envs $((131072 / n)) --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-steady > gpurun_out/${tag}_strong_${n}gpu.json 2> gpurun_out/${tag}_strong_${n}gpu.err
echo "strong rc=$?"; python -c "import json; d=json.loads(open('gpurun_out/${tag}_strong_${n}gpu.json').read().strip().splitlines()[-1]); print('strong n=$n', d['value'], d['ms_per_step'], d['e2e']['value'])"
