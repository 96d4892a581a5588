#!/bin/bash
# C5 iteration: the large-n parity tests, then C5 timings (FP16 split vs TF32 split)
tag=${1:-c5}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_large.py tests/test_gpu_parity.py tests/test_gpu_cov.py -m gpu -q --timeout 900 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -25 $out/pytest_gpu_$tag.log | cut -c1-250
{
timeout 300 python tools/run_once.py C5 592 2
TOB200_LG_FP16=0 timeout 300 python tools/run_once.py C5 592 2
timeout 300 python tools/run_once.py C5 4096 2
} 2>&1 | tee $out/timings_$tag.txt | cut -c1-400
