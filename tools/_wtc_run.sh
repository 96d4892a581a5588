mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wtc.py -m gpu -q --timeout 600 > gpurun_out/pytest_wtc_a.log 2>&1
echo "rc=$?"; tail -30 gpurun_out/pytest_wtc_a.log
