timeout 300 python tools/run_once.py C5 4096 2 2>&1 | tail -2
bash tools/gpu_test.sh r2c
bash tools/gpu_sanitize2.sh
