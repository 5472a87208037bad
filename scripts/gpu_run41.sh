set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_style.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r41_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:enkf_kernel -s 1 -c 1 -o gpurun_out/r41_enkf python scripts/profile_generic.py enkf > gpurun_out/r41_enkf.log 2>&1
tail -3 gpurun_out/r41_enkf.log
