#!/usr/bin/env bash
mkdir -p gpurun_out
python tools/run_small.py 262144 && ncu --set full --import-source on --clock-control none -k regex:ht_warp_kernel -c 1 -o gpurun_out/small_warp_v3 -f python tools/run_small.py 262144 > gpurun_out/ncu_small_warp.log 2>&1
tail -3 gpurun_out/ncu_small_warp.log
