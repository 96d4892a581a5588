#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py -m gpu -q --timeout 300 --tb=short > gpurun_out/sweep.log 2>&1
tail -3 gpurun_out/sweep.log
