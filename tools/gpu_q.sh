#!/usr/bin/env bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --configs cfg4,cfg5_shard 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', d['ms_per_step'], d['value'], d['e2e']['value'], d['configs']['cfg4']['ms'], d['configs']['cfg5_shard']['ms'])"
