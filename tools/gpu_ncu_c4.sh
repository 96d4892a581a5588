#!/bin/bash
# ncu --set full captures of wpp_lm_run_kernel at C4 shape for the main build and the listed variants
tag=${1:-c4}
out=gpurun_out
mkdir -p $out
for v in main $VARIANTS; do
  if [ "$v" != "main" ]; then export TOB200_LIB_OVERRIDE=$PWD/tinyopt_b200/libtinyopt_b200_$v.so; else unset TOB200_LIB_OVERRIDE; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wpp_lm_run -s 1 -c 1 \
    -f -o $out/prof_wpp_C4_${tag}_$v python tools/run_once.py C4 ${NPROB:-37888} 2 > $out/ncu_full_C4_${tag}_$v.log 2>&1
  python tools/ncu_summary.py $out/prof_wpp_C4_${tag}_$v.ncu-rep 45 > $out/ncu_summary_C4_${tag}_$v.txt 2>&1
  head -40 $out/ncu_summary_C4_${tag}_$v.txt
done
