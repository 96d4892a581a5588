#!/bin/bash
# compute-sanitizer passes over small instances of every kernel family (memcheck; racecheck for the
# shared-memory protocols).  gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
{
for cfg in "C2 2048" "C3 1024" "C4 600" "C5 6"; do
  set -- $cfg
  echo "== memcheck $1 $2"
  timeout 600 $CS --tool memcheck --error-exitcode 9 python tools/run_once.py $1 $2 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|iters|Error|error" | head -8
done
echo "== memcheck device functor"
timeout 600 $CS --tool memcheck --error-exitcode 9 tests/cuda/build/test_device_functor 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | head -8
echo "== memcheck cov tests"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cov.py -m gpu -q -x -k "small_bitexact and (6 or 50) or large_bitexact and 96" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | head -8
for cfg in "C4 300" "C2 1024"; do
  set -- $cfg
  echo "== racecheck $1 $2"
  timeout 900 $CS --tool racecheck --error-exitcode 9 python tools/run_once.py $1 $2 1 2>&1 | grep -E "RACECHECK SUMMARY|hazard|iters|Error" | head -8
done
} 2>&1 | tee $out/sanitize.txt
