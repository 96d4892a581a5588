#!/bin/bash
# multi-GPU evidence (gpurun --gpus N -- 'bash tools/gpu_multi.sh N tag'): bench under torchrun
N=${1:-2}
tag=${2:-r1}
out=gpurun_out
mkdir -p $out
for c in ${CONFIGS:-C2 C3 C4 C5}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --config $c --steps 5 --warmup 3 > $out/bench_${c}_N${N}_$tag.json 2> $out/bench_${c}_N${N}_$tag.err
  echo "bench $c N=$N rc=$?"; tail -1 $out/bench_${c}_N${N}_$tag.json | cut -c1-900
done
