#!/usr/bin/env bash
cd /root/repo
timeout 300 python tools/mid_time.py 2>&1 | tail -5
echo "--- HTN_MID_PROJ=2"
HTN_MID_PROJ=2 timeout 300 python tools/mid_time.py 2>&1 | tail -5
