#!/usr/bin/env bash
# round-2 GPU call A: full GPU test suite, then the bench line with the configs object
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_tests.txt
echo "tests rc=$?" >> gpurun_out/r2a_tests.txt
tail -5 gpurun_out/r2a_tests.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
