#!/usr/bin/env bash
cd /root/repo
python tools/run_small.py 1048576 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/small_launches.csv python tools/run_small.py 1048576 > gpurun_out/small_ncu.log 2>&1
