set -x
mkdir -p gpurun_out
CDK_LW_TOKEN=1 timeout 300 python scripts/trace_lw.py --out gpurun_out/r26_trace_token1.json
CDK_LW_TOKEN=3 timeout 300 python scripts/trace_lw.py --out gpurun_out/r26_trace_token3.json
