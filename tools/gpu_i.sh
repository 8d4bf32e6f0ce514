#!/usr/bin/env bash
cd /root/repo
python tools/mid_one.py > gpurun_out/mid_one.log 2>&1 || { tail -5 gpurun_out/mid_one.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/mid_launches.csv python tools/mid_one.py > gpurun_out/mid_ncu.log 2>&1
tail -2 gpurun_out/mid_one.log
