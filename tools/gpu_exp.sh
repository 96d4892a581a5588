#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu_final.log; tail -4 $out/pytest_gpu_final.log
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "== memcheck wpp double lm_run / step / cov (pytest subset)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solver.py -m gpu -q -x -k "wpp_double or (solver_matches_fused_run and (20 or 40 or 9 or 55))" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | head -8
echo "== racecheck wpp double + step"
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "solver_matches_fused_run and (20 or 9)" 2>&1 | grep -E "RACECHECK SUMMARY|hazard|passed|failed" | head -8
echo "== memcheck C5 (tensor-map loader) 6 problems"
timeout 600 $CS --tool memcheck --error-exitcode 9 python tools/run_once.py C5 6 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|iters" | head -4
} 2>&1 | tee $out/sanitize2.txt
