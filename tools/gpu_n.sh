#!/usr/bin/env bash
cd /root/repo
for w in 512 2048 4096 16384; do echo "--- W=$w (LA off)"; HTN_POTRF_LA=0 HTN_POTRF_W=$w python tools/potrf_bench.py | tail -4; done
