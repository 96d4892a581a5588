out=gpurun_out; tag=r3c
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log; tail -4 $out/pytest_gpu_$tag.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -6 $out/smoke_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$?"; tail -5 $out/bench_all_$tag.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$?"
