mkdir -p gpurun_out
{
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 900 python bench.py --config C4 --steps 5 --warmup 3 > gpurun_out/bench_C4_r2e.json 2> gpurun_out/bench_C4_r2e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_C4_r2e.err
} > gpurun_out/wtc10.txt 2>&1
cat gpurun_out/wtc10.txt
python - <<'P'
import json
r=json.loads(open('gpurun_out/bench_C4_r2e.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], json.dumps(r['roofline'])[:1500])
print('e2e', r.get('e2e'))
P
