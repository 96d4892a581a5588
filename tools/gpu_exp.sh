#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 300 --tb=short -k jtj > gpurun_out/sweep.log 2>&1
tail -3 gpurun_out/sweep.log
