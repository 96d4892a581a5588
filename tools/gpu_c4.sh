#!/bin/bash
# iteration: parity suite (bit-exactness is the regression test), then C4 timings (main build + variants)
tag=${1:-c4}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -30 $out/pytest_gpu_$tag.log | cut -c1-300
{
timeout 300 python tools/run_once.py C4 ${NPROB:-262144} 3
for v in $VARIANTS; do
  echo "variant $v"; TOB200_LIB_OVERRIDE=$PWD/tinyopt_b200/libtinyopt_b200_$v.so timeout 300 python tools/run_once.py C4 ${NPROB:-262144} 3
done
} 2>&1 | tee $out/timings_$tag.txt | cut -c1-1500
