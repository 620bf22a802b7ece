#!/bin/bash
# quick A/B: bench at 16384 and 131072 envs (20 steps) and 131072 envs (100 steps)
run() { python bench.py --envs $1 --steps $2 --warmup $3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
echo "== 16384 x20"; run 16384 20 3
echo "== 131072 x20"; run 131072 20 3
echo "== 131072 x100"; run 131072 100 10
