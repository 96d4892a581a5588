#!/bin/bash
# compute-sanitizer passes over what the last session of round 2 added: the SolverType seam / user-filled accumulators /
# InvCov on the general family (gn_solve_kernel modes 0 and 3, gn_import_hg_kernel), the sparse-H scatter, and the run-time-n
# functor drivers (functor_eval_kernel: Jets, own rows, numeric differentiation).
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize4.sh'
out=gpurun_out; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
F='ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|passed|failed|Error|error'
{
echo "== memcheck: seam above n = 55, hg above 55, sparse H, option variants (pytest subset)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_solver.py \
  -k "(seam_above_55_matches and (float64-9-150-56 or float32-7-200-72)) or (hg_prior_above_55 and 64) or hg_sparse or seam_above_55_option" 2>&1 | grep -E "$F" | head -8
echo "== memcheck: InvCov on the general family (double n = 65, 96)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_cov.py -k "general_family and (65 or 96)" 2>&1 | grep -E "$F" | head -8
echo "== memcheck: device functor program (incl. the *Large drivers: Jets, own rows, numeric differentiation)"
timeout 1200 $CS --tool memcheck --error-exitcode 9 tests/cuda/build/test_device_functor 2>&1 | grep -E "ERROR SUMMARY|Invalid|all device|failed" | head -6
echo "== racecheck: seam above n = 55 (double n = 56), InvCov general family (n = 65)"
timeout 1200 $CS --tool racecheck --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_solver.py tests/test_gpu_cov.py \
  -k "(seam_above_55_matches and float64-9-150-56) or (general_family and 65)" 2>&1 | grep -E "$F" | head -8
} 2>&1 | tee $out/sanitize_r4.txt
