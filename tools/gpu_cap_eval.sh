#!/bin/bash
# one `ncu --set full` capture of lg_eval_kernel at the C5 shape (592 problems), summarised by tools/ncu_summary.py
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lg_eval -s 1 -c 1 \
  -f -o $out/prof_lg_eval_C5_r4 python tools/run_once.py C5 592 2 > $out/ncu_full_lg_eval_C5_r4.log 2>&1
python tools/ncu_summary.py $out/prof_lg_eval_C5_r4.ncu-rep 30 > $out/lg_eval_C5_ncu_full_summary_r4.txt 2>&1
head -40 $out/lg_eval_C5_ncu_full_summary_r4.txt | cut -c1-170
