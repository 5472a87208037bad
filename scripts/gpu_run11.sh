set -x
timeout 1200 python -m pytest tests -m gpu -x -q -k "enkf" 2>&1 | tail -25 | tee gpurun_out/r11_pytest_enkf.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r11_pytest.log
