#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_cov.py -m gpu -q --timeout 300 2>&1 | tail -8
timeout 120 python tools/run_once.py C5 592 2 2>&1 | tail -2
