#!/bin/bash
# Development iteration on the GPU box: parity tests, then quick timings of C4 / C5 at reduced batch.
tag=${1:-it}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log
tail -15 $out/pytest_gpu_$tag.log
{
timeout 300 python tools/run_once.py C4 131072 3
timeout 300 python tools/run_once.py C5 592 2
TOB200_LG_TF32_TERMS=1 timeout 300 python tools/run_once.py C5 592 2
timeout 300 python tools/run_once.py C2 100000 5
timeout 300 python tools/run_once.py C3 100000 5
timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline
TOB200_HOST_CHUNKS=1 timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline
TOB200_HOST_CHUNKS=8 timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline
} 2>&1 | tee $out/timings_$tag.txt
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lg_syrk -s 0 -c 1 \
  -f -o $out/prof_lg_syrk_C5_$tag python tools/run_once.py C5 148 1 > $out/ncu_full_C5_syrk_$tag.log 2>&1
ls -la $out | tail
fi
