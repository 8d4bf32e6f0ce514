#!/usr/bin/env bash
cd /root/repo
for s in 86 96 104 112 120 128 148; do echo "SHARED_SMS=$s"; HTN_RNG_SHARED_SMS=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; done
