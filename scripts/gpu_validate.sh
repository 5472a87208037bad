#!/bin/bash
# The round's validation pass on a B200 box (run through gpurun): GPU parity tests, smoke, the bench line with its CPU
# baseline, the reference arm, the ncu launch list of the bench command and the per-config table.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash scripts/gpu_validate.sh [tag]'
set -x
tag=${1:-validate}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 1500 python scripts/bench_configs.py c1 c2 c3 c4 c5 2>&1 | tee gpurun_out/${tag}_configs.jsonl
