mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wtc.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
timeout 120 python tools/wtc_check.py 131072 500 50 3 2>&1 | sed -n 2,3p
} > gpurun_out/wtc22.txt 2>&1
cat gpurun_out/wtc22.txt
