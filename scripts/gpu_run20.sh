set -x
mkdir -p gpurun_out
./scripts/micro/fp64_pipe | tee gpurun_out/r20_fp64_pipe.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r20_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r20_bench.err | tee gpurun_out/r20_bench.json
