set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r22_pytest.log
for v in 0 1200 2500 3500 5000; do
  CDK_LW_STAGGER=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --simple-data 2>>gpurun_out/r22_bench.err | tee -a gpurun_out/r22_variants.jsonl
done
CDK_LW_STAGGER=2500 CDK_LW_SYNC=32 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --simple-data 2>>gpurun_out/r22_bench.err | tee -a gpurun_out/r22_variants.jsonl
timeout 300 python scripts/trace_lw.py --out gpurun_out/r22_trace_lw.json
