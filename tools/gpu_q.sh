#!/usr/bin/env bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do python tools/run_cfg.py cfg4 cfg5_shard 2>&1 | grep -o "\"ms\": [0-9.]*" | tr "\n" " "; echo; done
