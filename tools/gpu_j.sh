#!/usr/bin/env bash
cd /root/repo
python tools/mid_one.py > gpurun_out/mid_one.log 2>&1 || { tail -5 gpurun_out/mid_one.log; exit 1; }
# second call's kernels: skip the first call (1 fill + 4 U + 4 potf + 3 R + 1 project = 13 launches)
ncu --set full --import-source on --clock-control none -k regex:'chol_rows|project2|potf64' --launch-skip 12 --launch-count 12 -o gpurun_out/mid_full -f python tools/mid_one.py > gpurun_out/mid_ncu_full.log 2>&1
tail -2 gpurun_out/mid_ncu_full.log
