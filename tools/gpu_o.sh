#!/usr/bin/env bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x -k "structured or sp_ or cfg3 or cfg4 or full_size or edge_cases or thread or paper or colmajor or wrapper or single or k2 or potrs or blocked or large" 2>&1 | tail -4
python tools/run_cfg.py cfg3 cfg4 2>&1 | cut -c1-200
HTN_TRTRI=0 python tools/run_cfg.py cfg3 2>&1 | cut -c1-200
