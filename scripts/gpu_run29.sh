set -x
mkdir -p gpurun_out
CDK_EKF_MODE=pool timeout 900 python -m pytest tests -m gpu -x -q -k "ekf or batched or nonfinite or ll_sum or chunked" 2>&1 | tail -15 | tee gpurun_out/r29_pytest_pool.log
for t in 16 8 24; do
CDK_EKF_MODE=pool CDK_POOL_T=$t timeout 300 python scripts/trace_lw.py --out gpurun_out/r29_trace_pool_t$t.json
done
CDK_EKF_MODE=pool timeout 300 python scripts/trace_lw.py --no-outputs --out gpurun_out/r29_trace_pool_llonly.json
