#!/usr/bin/env bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "mid_size" 2>&1 | tail -3
for p in 1 0; do echo "--- rb7 $p"; HTN_MID_PROJ_RB7=$p timeout 300 python tools/mid_time.py 2>&1 | tail -4; done
