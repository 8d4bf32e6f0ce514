#!/usr/bin/env bash
cd /root/repo
python tools/mid_one.py > gpurun_out/mid_one.log 2>&1 || { tail -5 gpurun_out/mid_one.log; exit 1; }
HTN_MID_SPLIT=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/mid_launches.csv python tools/mid_one.py > gpurun_out/mid_ncu.log 2>&1
HTN_MID_SPLIT=0 ncu --set full --import-source on --clock-control none -k regex:'chol_rows|project|potf64' --launch-skip 12 --launch-count 12 -o gpurun_out/mid_full -f python tools/mid_one.py > gpurun_out/mid_ncu_full.log 2>&1
tail -2 gpurun_out/mid_ncu_full.log
