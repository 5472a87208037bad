#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_e_bench_$N.err | tee gpurun_out/r02_e_bench_$N.json
tail -5 gpurun_out/r02_e_bench_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>>gpurun_out/r02_e_bench_$N.err | tee gpurun_out/r02_e_bench_ref_$N.json
