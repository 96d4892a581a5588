out=gpurun_out; tag=r3d
mkdir -p $out
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$?"; tail -5 $out/bench_all_$tag.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$?"
timeout 600 python -m pytest tests/test_gpu_wtc.py -m gpu -q --timeout 300 2>&1 | tail -2
