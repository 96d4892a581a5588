#!/bin/bash
out=gpurun_out; mkdir -p $out
{
for d in 16 1 2 4 3 5 6 7; do
  echo "== TOB200_LG_DEBUG=$d"
  TOB200_LG_DEBUG=$d timeout 300 python tools/run_once.py C5 592 2 2>&1 | grep -v "^$" | tail -8
done
} 2>&1 | tee $out/c5dbg_${1:-a}.txt | cut -c1-300
