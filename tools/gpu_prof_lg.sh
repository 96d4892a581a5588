#!/bin/bash
tag=${1:-x}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_syrk -s 0 -c 1 \
  -f -o $out/prof_lg_syrk_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_syrk_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_solve -s 0 -c 1 \
  -f -o $out/prof_lg_solve_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_solve_$tag.log 2>&1
tail -2 $out/ncu_full_C5_syrk_$tag.log
