set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r21_pytest.log
for v in "14 0" "7 0" "14 16" "14 4"; do
  set -- $v
  CDK_LW_WPC=$1 CDK_LW_SYNC=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --simple-data 2>>gpurun_out/r21_bench.err | tee -a gpurun_out/r21_variants.jsonl
done
timeout 300 python scripts/trace_lw.py --out gpurun_out/r21_trace_lw.json
CDK_LW_SYNC=16 timeout 300 python scripts/trace_lw.py --out gpurun_out/r21_trace_lw_sync16.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_small_lw -c 1 -o gpurun_out/r21_lw14 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --simple-data > gpurun_out/r21_ncu.log 2>&1
