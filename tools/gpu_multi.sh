#!/bin/bash
# multi-GPU evidence (gpurun --gpus N -- 'bash tools/gpu_multi.sh N tag'): the DEFAULT bench line under torchrun, as the
# driver launches it, plus the reference arm
N=${1:-2}
tag=${2:-r2}
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_all_N${N}_$tag.json 2> $out/bench_all_N${N}_$tag.err
echo "bench N=$N rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -4 $out/bench_all_N${N}_$tag.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/bench_all_N${N}_$tag.json") if l.startswith("{")][-1])
    print("headline", d["config"]["workload"][:20], "value %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], d["scaling"], "e2e %.4g"%d["e2e"]["value"], d["run"]["parallelism"][:60], d["clocks"])
    for c in d.get("configs", []): print(c["name"], "value %.4g"%c["value"], "ms/step %.3f"%c["ms_per_step"], c["scaling"], "e2e %.4g"%c["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 5 --warmup 1 > $out/bench_ref_N${N}_$tag.json 2> $out/bench_ref_N${N}_$tag.err
echo "ref N=$N rc=$?"; head -c 400 $out/bench_ref_N${N}_$tag.json; echo
