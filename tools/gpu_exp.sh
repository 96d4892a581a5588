#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | tail -15
