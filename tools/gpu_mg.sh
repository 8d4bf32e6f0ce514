#!/usr/bin/env bash
# multi-GPU bench: N from $1
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "rc=$?"
tail -c 4500 gpurun_out/r2_bench_${N}gpu.json
tail -5 gpurun_out/r2_bench_${N}gpu.err
