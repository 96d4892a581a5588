#!/bin/bash
# the default (all-config) bench line + the reference arm, as the driver runs them
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -5 $out/bench_all_$tag.err; head -c 9000 $out/bench_all_$tag.json
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"; head -c 3000 $out/bench_ref_$tag.json
