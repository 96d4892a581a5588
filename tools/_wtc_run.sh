mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wtc.py -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/wtc29.txt
cat gpurun_out/wtc29.txt
