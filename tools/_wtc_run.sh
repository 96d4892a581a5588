mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wtc.py tests/test_gpu_parity.py -m gpu -q --timeout 600 > gpurun_out/pytest_wtc_b.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/pytest_wtc_b.log
timeout 300 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-probes 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(r['value'], r['roofline']['kernel'], r['roofline']['frac'], r['roofline']['traffic'])"
