#!/usr/bin/env bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "mid_size or edge_cases" 2>&1 | tail -15
timeout 300 python tools/mid_time.py 2>&1 | tail -5
