set -x
mkdir -p gpurun_out
CDK_LW_TOKEN=2 timeout 300 python scripts/trace_lw.py --out gpurun_out/r25_trace_token2.json
CDK_LW_TOKEN=2 timeout 300 python scripts/trace_lw.py --no-outputs --out gpurun_out/r25_trace_token2_llonly.json
