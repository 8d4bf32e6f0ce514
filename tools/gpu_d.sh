#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "potrf or cfg5 or cfg4" 2>&1 | tail -4
echo "--- default"; python tools/potrf_bench.py 2>&1 | tail -6
echo "--- SLA off"; HTN_POTRF_SLA=1000000 python tools/potrf_bench.py 2>&1 | tail -3
echo "--- SLA from 2048, LA off"; HTN_POTRF_LA=1535 HTN_POTRF_SLA=2048 python tools/potrf_bench.py 2>&1 | tail -5
python tools/run_cfg.py cfg4 cfg5_shard 2>&1 | tail -2 | cut -c1-420
