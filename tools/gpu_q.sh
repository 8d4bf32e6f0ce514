#!/usr/bin/env bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "cfg4 or full_size or blocked or structured" 2>&1 | tail -3
for i in 1 2; do python tools/run_cfg.py cfg4 2>&1 | grep -o "\"ms\": [0-9.]*"; HTN_INV_PAR=0 python tools/run_cfg.py cfg4 2>&1 | grep -o "\"ms\": [0-9.]*"; done
