set -x
mkdir -p gpurun_out
CDK_LW_TOKEN=2 timeout 300 python scripts/trace_lw.py --out gpurun_out/r27_trace_permits2.json
CDK_LW_TOKEN=3 timeout 300 python scripts/trace_lw.py --out gpurun_out/r27_trace_permits3.json
CDK_LW_TOKEN=2 timeout 300 python scripts/trace_lw.py --no-outputs --out gpurun_out/r27_trace_permits2_llonly.json
