#!/usr/bin/env bash
# Run the GPU parity suite group by group, each in its own process (a CUDA fault in one group must
# not poison the others), logs under gpurun_out/.  Usage (on the GPU box): tools/gpu_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
fail=0
for grp in "full_size" "raw_streams or ziggurat" "gemm" "potrf" "hyperplane" "structured" "error_codes or single_sample or reference_properties or device_pointer or numpy_generator or colmajor or triangle or paper_sign or edge_cases or independent_problems or device_mode"; do
  tag=$(echo "$grp" | tr ' ' '_' | cut -c1-24)
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" -x --no-header -p no:cacheprovider > gpurun_out/test_$tag.log 2>&1
  rc=$?
  echo "== [$grp] rc=$rc: $(tail -1 gpurun_out/test_$tag.log)"
  if [ $rc -ne 0 ]; then fail=1; grep -E "^(FAILED|ERROR|E  )" gpurun_out/test_$tag.log | head -12; fi
done
exit $fail
