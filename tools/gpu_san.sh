#!/usr/bin/env bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "warp_per_problem or ziggurat_decisions or single_sample_api or independent_golden" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2_memcheck.log
