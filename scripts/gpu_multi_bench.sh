#!/bin/bash
# multi-GPU bench: gpurun --gpus N -- "bash scripts/gpu_multi_bench.sh N"  (no reference arm: that is rank 0 alone on the CPU, measured in the 1-GPU validation)
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02_t_bench_$N.err | tee gpurun_out/r02_t_bench_$N.json | cut -c1-300
tail -3 gpurun_out/r02_t_bench_$N.err
