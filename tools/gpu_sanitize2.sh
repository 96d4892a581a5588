#!/bin/bash
# compute-sanitizer passes over what round 2 added: the general family, the user-filled-accumulator step, the
# final-Hessian export, the robust functor runs, the FP16-split J^T J.   gpurun --timeout 2400 -- 'bash tools/gpu_sanitize2.sh'
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
F='ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|passed|failed|iters|Error|error'
{
echo "== memcheck C5 6 (FP16-split lg_syrk, lg_eval, lg_solve)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python tools/run_once.py C5 6 1 2>&1 | grep -E "$F" | head -6
echo "== memcheck C4 300 (conflict-free block columns)"
timeout 600 $CS --tool memcheck --error-exitcode 9 python tools/run_once.py C4 300 1 2>&1 | grep -E "$F" | head -6
echo "== memcheck general family / hg step / final Hessian (pytest subset)"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_general.py tests/test_gpu_solver.py tests/test_gpu_parity.py \
  -k "(double_above_55 and 56-6) or (inverse_path_above_55 and float32-56) or (build_solve_general and float64-3-20) or (option_variants_double_n60 and optkw0) or hg_singular or (hg_prior and float32-12) or (final_hessian and float32-40-64-20) or (final_hessian and float32-5-256-64)" 2>&1 | grep -E "$F" | head -8
echo "== memcheck device functor program (Jets, manual rows, robust M-estimators)"
timeout 900 $CS --tool memcheck --error-exitcode 9 tests/cuda/build/test_device_functor 2>&1 | grep -E "ERROR SUMMARY|Invalid|all device|failed" | head -6
echo "== racecheck general family (double n = 56) and the hg step"
timeout 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_general.py tests/test_gpu_solver.py \
  -k "(double_above_55 and 56-6) or (inverse_path_above_55 and float32-56) or hg_singular" 2>&1 | grep -E "$F" | head -8
echo "== racecheck C5 4 (FP16-split producers / MMA / epilogue protocol)"
timeout 1500 $CS --tool racecheck --error-exitcode 9 python tools/run_once.py C5 4 1 2>&1 | grep -E "$F" | head -6
} 2>&1 | tee $out/sanitize_r2.txt
