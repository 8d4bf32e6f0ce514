#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "warp_per_problem or independent or small_problems" 2>&1 | tail -30 > gpurun_out/r2b_tests.txt
tail -4 gpurun_out/r2b_tests.txt
HTN_SMALL_PROF=1 python tools/run_small.py 262144 2>&1 | tail -2
python tools/run_cfg.py cfg2 2>&1 | tail -1 | cut -c1-330 | tee gpurun_out/r2b_cfg2.txt
