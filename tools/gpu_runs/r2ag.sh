#!/bin/bash
run() { python bench.py --envs $1 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-steady 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; }
for c in 12 24 36; do for n in 16384 131072; do echo "== SEQ_CTAS=$c envs=$n"; SO101_SEQ_CTAS=$c run $n; done; done
