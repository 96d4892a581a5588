#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, bench lines for the single-GPU configs, the
# ncu launch list of the bench command and one `--set full` capture of each family's top kernel.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r1'
# Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py turns the reports into the
# summaries committed under profiles/.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nproc > $out/nproc_$tag.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi_$tag.csv

timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -3 $out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log

for c in ${CONFIGS:-C2 C3 C4 C5}; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > $out/bench_${c}_$tag.json 2> $out/bench_${c}_$tag.err
  echo "bench $c rc=$?"; cat $out/bench_${c}_$tag.json
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_C2_$tag.json 2>&1
cat $out/bench_ref_C2_$tag.json

# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/launches_C2_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_bench_C2_$tag.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $out/launches_C5_$tag.csv python tools/run_once.py C5 296 1 > $out/ncu_list_C5_$tag.log 2>&1
# full captures of the dominant kernel of each family
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp_lm_run -s 1 -c 1 \
  -f -o $out/prof_tpp_C2_$tag python tools/run_once.py C2 100000 2 > $out/ncu_full_C2_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp_lm_run -s 1 -c 1 \
  -f -o $out/prof_tpp_C3_$tag python tools/run_once.py C3 100000 2 > $out/ncu_full_C3_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wpp_lm_run -s 1 -c 1 \
  -f -o $out/prof_wpp_C4_$tag python tools/run_once.py C4 16384 2 > $out/ncu_full_C4_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_syrk -s 0 -c 1 \
  -f -o $out/prof_lg_syrk_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_syrk_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_solve -s 0 -c 1 \
  -f -o $out/prof_lg_solve_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_solve_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_eval -s 0 -c 1 \
  -f -o $out/prof_lg_eval_C5_$tag python tools/run_once.py C5 296 1 > $out/ncu_full_C5_eval_$tag.log 2>&1
ls -la $out
