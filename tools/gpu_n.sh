#!/usr/bin/env bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "potrf or full_size or blocked or cfg" 2>&1 | tail -3
python tools/potrf_bench.py | tail -5
echo "--- HTN_POTRF_PU=0"; HTN_POTRF_PU=0 python tools/potrf_bench.py | tail -5
echo "--- LA from 768"; HTN_POTRF_LA_MIN=768 python tools/potrf_bench.py | tail -5
