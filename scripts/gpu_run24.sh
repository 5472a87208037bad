set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r24_pytest.log
timeout 300 python scripts/trace_lw.py --out gpurun_out/r24_trace_token.json
CDK_LW_TOKEN=0 timeout 300 python scripts/trace_lw.py --out gpurun_out/r24_trace_notoken.json
timeout 300 python scripts/trace_lw.py --no-outputs --out gpurun_out/r24_trace_token_llonly.json
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r24_bench.err | tee gpurun_out/r24_bench.json
