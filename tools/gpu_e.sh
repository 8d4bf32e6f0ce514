#!/usr/bin/env bash
cd /root/repo
bash tools/gpu_f.sh 2>&1 | tail -12
HTN_RNG_COOP=100000 python -m pytest tests -m gpu -q -x -k "raw_streams or ziggurat or single_sample or numpy_generator or structured_golden or failed_cov or hyperplane_golden or large_batch" 2>&1 | tail -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/run_cfg.py cfg3 cfg4 2>&1 | tail -2 | cut -c1-300
