#!/usr/bin/env bash
cd /root/repo
for i in 1 2 3; do python tools/run_cfg.py cfg3 cfg4 2>&1 | grep -o "^[a-z0-9_]* \|\"ms\": [0-9.]*" | tr "\n" " "; echo; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --configs cfg5_shard,mid256 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', d['ms_per_step'], d['value'], d['e2e']['value'], d['configs']['cfg5_shard']['ms'], d['configs']['mid256']['ms'])"
