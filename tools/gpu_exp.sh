#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 300 2>&1 | tail -3
timeout 120 python tools/run_once.py C5 592 2 2>&1 | tail -2
TOB200_LG_NO_TMAP=1 timeout 120 python tools/run_once.py C5 592 2 2>&1 | tail -2
TOB200_LG_DEBUG=16 timeout 120 python tools/run_once.py C5 592 1 2>&1 | grep -v "^C5" | head -4
