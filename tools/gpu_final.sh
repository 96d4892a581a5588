#!/bin/bash
# last call of the round: the whole GPU suite, smoke, and the default bench line on the final library
out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 > $out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_final.log; tail -4 $out/pytest_gpu_final.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_final.log
timeout 200 python bench.py --steps 20 --warmup 5 > $out/bench_all_final.json 2> $out/bench_all_final.err
echo "bench rc=$?"; head -c 300 $out/bench_all_final.json; echo
