#!/usr/bin/env bash
cd /root/repo; mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; echo "ref rc=$?"
python tools/single_latency.py 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_plain_small.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_ncu.log 2>&1
echo "ncu rc=$?"
