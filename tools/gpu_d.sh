#!/usr/bin/env bash
cd /root/repo
python -m pytest tests -m gpu -q -x -k "potrf or golden or fresh or error_codes" 2>&1 | tail -4
echo "--- v2"; python tools/potrf_bench.py 2>&1 | tail -5
echo "--- v1"; HTN_POTF2_V2=0 python tools/potrf_bench.py 2>&1 | tail -5
python tools/potrf_small_probe.py 2>&1 | tail -6
