#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py -m gpu -q --timeout 300 --tb=short > gpurun_out/sweep.log 2>&1
tail -3 gpurun_out/sweep.log
timeout 300 tests/cuda/build/test_device_functor | tail -3
timeout 120 python tools/run_once.py C4 131072 3 2>&1 | tail -1
