#!/bin/bash
# Round-2, second evidence set (after the mid-n tensor-core kernel, wtc.cuh): GPU parity suite, smoke, the default bench
# line and the reference arm as the driver runs them, the ncu launch list of the bench command, `--set full` captures of
# the C4 kernels at the bench size, compute-sanitizer passes over the new kernel.
#   gpurun --timeout 3000 -- 'bash tools/gpu_round3.sh r3'
tag=${1:-r3}
out=gpurun_out
mkdir -p $out
nproc > $out/nproc_$tag.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi_$tag.csv
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log; tail -12 $out/pytest_gpu_$tag.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -6 $out/smoke_$tag.log
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -5 $out/bench_all_$tag.err; head -c 600 $out/bench_all_$tag.json; echo
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"; head -c 400 $out/bench_ref_$tag.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-probes > $out/ncu_bench_$tag.log 2>&1
python tools/launch_shares.py $out/launches_bench_$tag.csv > $out/launches_bench_${tag}_shares.txt 2>&1; head -14 $out/launches_bench_${tag}_shares.txt
cap() {  # name kernel-regex config problems
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${5:-1} -c 1 \
    -f -o $out/prof_$1_$tag python tools/run_once.py $3 $4 2 > $out/ncu_full_$1_$tag.log 2>&1
  python tools/ncu_summary.py $out/prof_$1_$tag.ncu-rep 30 > $out/$1_ncu_full_summary_$tag.txt 2>&1
  head -24 $out/$1_ncu_full_summary_$tag.txt | cut -c1-160
}
cap wtc_C4 wtc_lm_run C4 1000000
python tools/ncu_traffic.py C4:$out/prof_wtc_C4_$tag.ncu-rep:1000000:1000000 > $out/ncu_traffic_$tag.log 2>&1
cp profiles/ncu_traffic.json $out/ncu_traffic_$tag.json
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "== memcheck wtc (C4 shape, 300 problems)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python tools/run_once.py C4 300 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|iters|Error|error" | head -8
echo "== memcheck wtc tests (odd n, partial chunks, degenerate systems)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_wtc.py -m gpu -q -x -k "parity and (30-100-51 or 25-501-44 or 9-500-50) or degenerate" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | head -8
echo "== racecheck wtc (C4 shape, 64 problems)"
timeout 900 $CS --tool racecheck --error-exitcode 9 python tools/run_once.py C4 64 1 2>&1 | grep -E "RACECHECK SUMMARY|hazard|iters|Error" | head -8
} 2>&1 | tee $out/sanitize_$tag.txt
ls -la $out | tail -5
