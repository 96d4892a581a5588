#!/bin/bash
# round-2 iteration: GPU parity suite + smoke + the default (all-config) bench line + the reference arm
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nproc > $out/nproc_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -15 $out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -5 $out/smoke_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$?"; tail -5 $out/bench_all_$tag.err; head -c 6000 $out/bench_all_$tag.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$?"; grep -i "elapsed" $out/bench_ref_$tag.err; head -c 3000 $out/bench_ref_$tag.json
