#!/bin/bash
out=gpurun_out; mkdir -p $out
{
for r in 2 3 4 5; do
  echo "== TOB200_LG_RAW_STAGES=$r"
  TOB200_LG_RAW_STAGES=$r timeout 300 python tools/run_once.py C5 592 2 2>&1 | tail -2
done
timeout 600 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 2>&1 | tail -3
} 2>&1 | tee $out/c5sweep_${1:-a}.txt | cut -c1-300
