#!/bin/bash
# Round-2 evidence set (one gpurun call): GPU parity suite, smoke, the default bench line (C2-C5 records, headline C4)
# and the reference arm as the driver runs them, the ncu launch list of the bench command, `--set full` captures of the
# dominant kernels at the bench sizes, timings of the general family.   gpurun --timeout 3000 -- 'bash tools/gpu_round2.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
nproc > $out/nproc_$tag.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/smi_$tag.csv
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=5 > $out/pytest_gpu_$tag.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_gpu_$tag.log; tail -12 $out/pytest_gpu_$tag.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -5 $out/smoke_$tag.log
t0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_all_$tag.json 2> $out/bench_all_$tag.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -3 $out/bench_all_$tag.err; head -c 1500 $out/bench_all_$tag.json; echo
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"; head -c 600 $out/bench_ref_$tag.json; echo
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $out/launches_bench_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-probes > $out/ncu_bench_$tag.log 2>&1
python tools/launch_shares.py $out/launches_bench_$tag.csv > $out/launches_bench_${tag}_shares.txt 2>&1; head -14 $out/launches_bench_${tag}_shares.txt
# full captures of the dominant kernels AT THE BENCH SIZES
cap() {  # name kernel-regex config problems
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${5:-1} -c 1 \
    -f -o $out/prof_$1_$tag python tools/run_once.py $3 $4 2 > $out/ncu_full_$1_$tag.log 2>&1
  python tools/ncu_summary.py $out/prof_$1_$tag.ncu-rep 30 > $out/$1_ncu_full_summary_$tag.txt 2>&1
  head -24 $out/$1_ncu_full_summary_$tag.txt | cut -c1-160
}
cap tpp_C2 tpp_lm_run C2 100000
cap tpp_C3 tpp_lm_run C3 100000
cap wpp_C4 wpp_lm_run C4 1000000
cap lg_syrk_C5 lg_syrk C5 4096 4
cap lg_solve_C5 lg_solve C5 592 4
cap lg_eval_C5 lg_eval C5 592 4
python tools/ncu_traffic.py C2:$out/prof_tpp_C2_$tag.ncu-rep:100000:100000 C3:$out/prof_tpp_C3_$tag.ncu-rep:100000:100000 \
  C4:$out/prof_wpp_C4_$tag.ncu-rep:1000000:1000000 C5:$out/prof_lg_syrk_C5_$tag.ncu-rep:4096:4096 > $out/ncu_traffic_$tag.log 2>&1
cp profiles/ncu_traffic.json $out/ncu_traffic_$tag.json
# the general family (coverage path): timings
python - <<'PY' 2>&1 | tee $out/gn_timings_$tag.txt
import sys, time, torch
sys.path.insert(0, '.')
import tinyopt_b200 as tb
ctx = tb.Context(0)
for dt, B, m, n in ((torch.float64, 296, 1024, 256), (torch.float64, 148, 4096, 512), (torch.float32, 148, 2048, 1024), (torch.float32, 148, 4096, 2048)):
    A, y, xs, x0 = ctx.synth_generate(B, m, n, dt, layout=tb.PROBLEM_MAJOR)
    kw = {} if dt == torch.float64 else dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = ctx.optimize_batch(A, y, x0, tb.options(**kw), layout=tb.PROBLEM_MAJOR)
        torch.cuda.synchronize(); dt_s = time.perf_counter() - t0
    it = int(out.results["num_iters"].sum())
    print(f"gn family {str(dt).split('.')[-1]} B={B} m={m} n={n}: iters={it} {dt_s*1e3:.1f} ms -> {it/dt_s:.0f} it/s; phases ms eval+accum {ctx.last_phase_ms(0)} solve {ctx.last_phase_ms(2)}")
PY
ls -la $out | tail -5
