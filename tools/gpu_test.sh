#!/bin/bash
# GPU parity suite only (optionally a subset: TESTS="tests/test_gpu_general.py")
tag=${1:-t}
out=gpurun_out
mkdir -p $out
timeout 2400 python -m pytest ${TESTS:-tests} -m gpu -q --timeout 1200 --durations=8 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -${TAIL:-40} $out/pytest_gpu_$tag.log | cut -c1-400
