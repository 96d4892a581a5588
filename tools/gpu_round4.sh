#!/bin/bash
# Round-2, last evidence set (after the seam / covariance / functors above n = 55, numeric differentiation, sparse-H
# signature): GPU parity suite, smoke, the default bench line and the reference arm as the driver runs them, the ncu
# launch list of the bench command, compute-sanitizer passes over the new kernels.
#   gpurun --timeout 3000 -- 'bash tools/gpu_round4.sh r4'
tag=${1:-r4}
out=gpurun_out
mkdir -p $out
nproc > $out/nproc_$tag.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi_$tag.csv
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log; tail -12 $out/pytest_gpu_$tag.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -6 $out/smoke_$tag.log
bash tools/gpu_sanitize4.sh > /dev/null 2>&1; cat $out/sanitize_r4.txt
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -5 $out/bench_all_$tag.err; head -c 600 $out/bench_all_$tag.json; echo
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"; head -c 400 $out/bench_ref_$tag.json; echo
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-probes > $out/ncu_bench_$tag.log 2>&1
python tools/launch_shares.py $out/launches_bench_$tag.csv > $out/launches_bench_${tag}_shares.txt 2>&1; head -14 $out/launches_bench_${tag}_shares.txt
fi
ls -la $out | tail -5
