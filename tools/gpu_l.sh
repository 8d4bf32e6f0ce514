#!/usr/bin/env bash
cd /root/repo
for c in cfg3 cfg4; do
python tools/cfg3_one.py $c > gpurun_out/${c}_one.log 2>&1 || { tail -5 gpurun_out/${c}_one.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${c}_launches.csv python tools/cfg3_one.py $c > gpurun_out/${c}_ncu.log 2>&1
done
