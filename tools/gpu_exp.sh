#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py tests/test_cpp_adaptor.py tests/test_device_functor.py -m gpu -q --timeout 300 --tb=short > gpurun_out/sweep.log 2>&1
tail -3 gpurun_out/sweep.log
timeout 120 python tools/run_once.py C2 100000 5 2>&1 | tail -1
timeout 120 python tools/run_once.py C3 100000 5 2>&1 | tail -1
for sb in 8192 16384 32768; do for st in 2 3; do echo "stage_bytes $sb stages $st"; TOB200_TPP_STAGE_BYTES=$sb TOB200_TPP_STAGES=$st timeout 120 python tools/run_once.py C2 100000 5 2>&1 | tail -1; TOB200_TPP_STAGE_BYTES=$sb TOB200_TPP_STAGES=$st timeout 120 python tools/run_once.py C3 100000 5 2>&1 | tail -1; done; done
