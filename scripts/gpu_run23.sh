set -x
mkdir -p gpurun_out
./scripts/micro/rk4_pipe | tee gpurun_out/r23_rk4_pipe.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r23_pytest.log
timeout 300 python scripts/trace_lw.py --out gpurun_out/r23_trace_full.json
timeout 300 python scripts/trace_lw.py --no-outputs --out gpurun_out/r23_trace_llonly.json
CDK_EKF_TMA=0 timeout 300 python scripts/trace_lw.py --out gpurun_out/r23_trace_notma.json
timeout 1500 python scripts/bench_configs.py c2 --scale 0.25 2>&1 | tee gpurun_out/r23_configs_c2.jsonl
