#!/bin/bash
timeout 300 tests/cuda/build/test_device_functor
timeout 120 python tools/run_once.py C5 592 2 2>&1 | tail -2
