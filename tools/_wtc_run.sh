mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_wtc.py -m gpu -q --timeout 300 -x 2>&1 | tail -15
timeout 120 python tools/wtc_check.py 131072 500 50 3 2>&1 | sed -n 2,3p
timeout 300 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(r['value'], json.dumps(r.get('probes'))[:900])"
} > gpurun_out/wtc26.txt 2>&1
cat gpurun_out/wtc26.txt
