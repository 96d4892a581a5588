#!/bin/bash
# C5 evidence: full GPU test suite, the C5 bench line, launch list and full ncu capture of the tensor-core kernel
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -5 $out/pytest_gpu_$tag.log
timeout 900 python bench.py --config C5 --steps 5 --warmup 3 > $out/bench_C5_$tag.json 2> $out/bench_C5_$tag.err
echo "bench C5 rc=$?"; cat $out/bench_C5_$tag.json; tail -5 $out/bench_C5_$tag.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $out/launches_C5_$tag.csv python tools/run_once.py C5 296 1 > $out/ncu_list_C5_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_syrk -s 0 -c 1 \
  -f -o $out/prof_lg_syrk_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_syrk_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_solve -s 0 -c 1 \
  -f -o $out/prof_lg_solve_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_solve_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_eval -s 0 -c 1 \
  -f -o $out/prof_lg_eval_C5_$tag python tools/run_once.py C5 296 1 > $out/ncu_full_C5_eval_$tag.log 2>&1
ls -la $out | tail -12
