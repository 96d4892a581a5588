#!/usr/bin/env python
"""bench.py — LM iterations/sec of the batched dense-NLLS hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (ours; N > 1 under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm: the oracle port,
                                                             all host threads, bounded sample)

One "step" = one complete batched LM solve (one tinyopt::Optimize() per problem) of the workload
config over synthetic inputs.  Metric = sum over problems of Output::num_iters / time.  Prints ONE
JSON line on rank 0.  See DESIGN.md §6 for what every field means.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs: name -> (B per GPU, m, n, dtype, option overrides)
FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
CONFIGS = {
    "C2": dict(B=100_000, m=30, n=6, dtype="f64", opts={}, desc="batch 100k problems, n=6 params, 30 residuals each, double"),
    "C3": dict(B=100_000, m=200, n=12, dtype="f32", opts=FLOAT_OPTS, desc="batch 100k problems, n=12 params, 200 residuals each, float"),
    # C4 is the sharded config: 1M problems in total, split over the ranks (strong scaling)
    "C4": dict(B=1_000_000, m=500, n=50, dtype="f32", opts=FLOAT_OPTS, strong=True,
               desc="batch 1M problems, n=50 params, 500 residuals, float, sharded over the GPUs"),
    # C5: the tensor-core config (tcgen05 3xTF32 JᵀJ + blocked LDLT), 4k problems in total
    "C5": dict(B=4096, m=4096, n=512, dtype="f32", opts=FLOAT_OPTS, strong=True,
               desc="batch 4k problems, n=512 params, 4096 residuals, float, tensor-core JᵀJ tile, sharded over the GPUs"),
}
SEED, ALPHA, SIGMA = 20261017, 0.1, 1e-2


def algorithmic_bytes(cfg, results):
    """s*(m*n + m) read + s*n written + 8 (cost) per problem-iteration that rebuilds
    H and g; a cost-only iteration needs only y (s*m) + 8."""
    s = 8 if cfg["dtype"] == "f64" else 4
    m, n = cfg["m"], cfg["n"]
    builds = int(results["num_builds"].astype(np.int64).sum())
    iters = int(results["num_iters"].astype(np.int64).sum())
    return builds * (s * (m * n + m) + s * n + 8) + (iters - builds) * (s * m + 8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bounded_sample(cfg, budget_bytes):
    """Problems of the workload whose inputs fit `budget_bytes` (a multiple of 32, at least 32,
    never more than the config's batch): the bounded sample of the CPU and e2e legs."""
    s = 8 if cfg["dtype"] == "f64" else 4
    per = s * (cfg["m"] * cfg["n"] + cfg["m"] + 2 * cfg["n"])
    return int(min(cfg["B"], max(32, (budget_bytes // per) // 32 * 32)))


def measured_tensor_peak():
    """TF32 dense peak: MEASURED_PEAKS.json has bf16 only; TF32 runs at half the bf16 rate on B200
    (tcgen05 K per instruction is 32 bytes for every dtype), so bf16 / 2 — said so in peak_source."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]) / 2, "measured bf16 burst / 2 (MEASURED_PEAKS.json bf16_tflops; no TF32 figure there)"
    except Exception:
        return 1590.0 / 2, "fallback bf16 1.59 PFLOP/s / 2 (B200_PROFILING.md)"


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel at the bench size, from
# the committed `ncu --set full` captures (profiles/r1_*_ncu_full_summary.txt); None where no capture at
# the bench size exists (C4's capture ran 16384 of the 1M problems)
NCU_TRAFFIC = {"C2": 245.520640e6 + 11.821312e6, "C3": 3.178103e9 + 13.871360e6, "C4": None}


def host_threads():
    """Every host core this process may run on — NOT OMP_NUM_THREADS, which torchrun pins to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference(cfg, sample_B, reps, nthreads=0):
    """The reference's CPU implementation of the path = the oracle port (Eigen is not in the image,
    so the reference itself cannot be compiled: DESIGN.md §3), OpenMP over problems, all host cores."""
    from oracle import oracle as O
    nthreads = nthreads or host_threads()
    dt = np.float64 if cfg["dtype"] == "f64" else np.float32
    A, y, xs, x0 = O.synth_generate(sample_B, cfg["m"], cfg["n"], dt, seed=SEED, alpha=ALPHA, sigma=SIGMA)
    opt = O.default_options(**cfg["opts"])
    w = min(2048, max(16, sample_B // 4))
    O.synth_lm_run(A[:w], y[:w], x0[:w], opt, alpha=ALPHA, nthreads=nthreads)  # warm-up
    times, iters, used = [], 0, 1
    for _ in range(reps):
        t0 = time.perf_counter()
        _, res, used = O.synth_lm_run(A, y, x0, opt, alpha=ALPHA, nthreads=nthreads)
        times.append(time.perf_counter() - t0)
        iters = int(res["num_iters"].astype(np.int64).sum())
    return iters, times, used


def run_reference(args, cfg, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = bounded_sample(cfg, 1 << 30)  # <= 1 GiB of inputs per step
    cpu_reference(cfg, sample_B, max(1, min(args.warmup, 2)))
    iters, times, used = cpu_reference(cfg, sample_B, args.steps)
    total = sum(times)
    value = iters * args.steps / total
    line = {
        "impl": "reference", "metric": "LM iterations/sec (batched dense NLLS)", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": f"{name}: {cfg['desc']}", "B": sample_B, "m": cfg["m"], "n": cfg["n"]},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": used, "kind": "port",
                         "sample": f"{sample_B} of the {cfg['B']} problems of {name} per step, {args.steps} steps, OpenMP over problems"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg, args.config)

    import torch
    import torch.distributed as dist

    import tinyopt_b200 as tb
    from tinyopt_b200 import api as tba

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    tdt = torch.float64 if cfg["dtype"] == "f64" else torch.float32
    m, n = cfg["m"], cfg["n"]
    strong = bool(cfg.get("strong"))
    from tinyopt_b200.shard import shard_range
    if strong:   # a fixed total batch split over the ranks
        lo, hi = shard_range(cfg["B"], rank, world)
    else:        # the config's batch per GPU: weak scaling over independent problems
        lo, hi = rank * cfg["B"], (rank + 1) * cfg["B"]
    B = hi - lo
    ctx = tb.Context(local_rank)
    opt = tb.options(**cfg["opts"])
    layout = tb.TILE32 if ctx.kernel_family(tdt, n) == 1 else tb.PROBLEM_MAJOR  # the family's native layout
    # rank g owns problems [lo, hi): generated in place from (seed, index), no scatter needed
    A, y, xs, x0 = ctx.synth_generate(B, m, n, tdt, p0=lo, seed=SEED, alpha=ALPHA, sigma=SIGMA, layout=layout)
    res_buf = torch.empty((B, tba.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    x = torch.empty_like(x0)

    lm_run = getattr(ctx._lib, f"tob200_lm_run_{cfg['dtype']}")
    c_alpha = (ctypes.c_double if cfg["dtype"] == "f64" else ctypes.c_float)(ALPHA)

    def step():
        x.copy_(x0)
        ctx._ck(lm_run(ctx._h, ctypes.byref(opt), tba._p(A), tba._p(y), c_alpha, layout, B, m, n, tba._p(x),
                       tba._p(res_buf)), "tob200_lm_run")

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    results = tba.decode_results(res_buf)
    iters_rank = int(results["num_iters"].astype(np.int64).sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    # per-launch duration of the dominant kernel, CUDA events on the launching stream, live
    family = ctx.kernel_family(tdt, n)
    kms, tensor_ms, tensor_launches = [], [], 0
    for _ in range(min(args.steps, 10 if family != 3 else 2)):
        x.copy_(x0)
        step()
        if family == 3:   # multi-kernel pipeline: the tensor-core JᵀJ launches of this solve, summed
            t_ms, tensor_launches = ctx.last_phase_ms(1)
            tensor_ms.append(t_ms)
            kms.append(sum(ctx.last_phase_ms(k)[0] for k in range(3)))
        else:
            kms.append(ctx.last_elapsed_ms())
    kernel_ms_avg = float(np.mean(kms))

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    it = torch.tensor([iters_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        # the one data-path collective: gather of the per-problem results (solutions stay sharded)
        from tinyopt_b200.shard import gather_rows
        Btot = cfg["B"] if strong else cfg["B"] * world
        gathered = gather_rows(res_buf, Btot, rank, world)
        assert gathered.shape[0] == Btot
    ms_total = float(t.item())
    iters_all = float(it.item())
    value = iters_all * args.steps / (ms_total * 1e-3)

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        abytes = algorithmic_bytes(cfg, results)
        achieved = abytes / (kernel_ms_avg * 1e-3) / 1e9
        if family == 3:
            # dominant kernel = lg_syrk_kernel (tcgen05): algorithmic flops m*n*(n+1) per rebuilt problem
            builds = int(results["num_builds"].astype(np.int64).sum())
            aflops = builds * m * n * (n + 1)
            t_ms = float(np.mean(tensor_ms))
            tpeak, tsrc = measured_tensor_peak()
            roof = {"bound": "tensor", "achieved": aflops / (t_ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                    "traffic": None, "kernel": "lg_syrk_kernel", "kernel_ms": t_ms / max(1, tensor_launches),
                    "launches_per_step": tensor_launches, "algorithmic_flops_per_step": aflops, "peak_source": tsrc,
                    "mode": "3xTF32 (3 MMAs per product term, FP32-level accuracy): the hardware executes 3x the algorithmic flops, "
                            "and 10 of 16 128x128 blocks for the n(n+1)/2 algorithmic triangle",
                    "pipeline_ms": {k: ctx.last_phase_ms(i)[0] for i, k in enumerate(("eval", "jtj", "solve"))},
                    "hbm_GBps_whole_pipeline": achieved}
            roof["frac"] = roof["achieved"] / roof["peak"]
        line = {
            "metric": "LM iterations/sec (batched dense NLLS)", "value": value, "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['desc']}", "B_per_gpu": B, "m": m, "n": n,
                       "options": "tinyopt defaults" + (" + float thresholds min_rerr_dec=1e-5 min_step_norm2=1e-9" if cfg["opts"] else ""),
                       "iters_per_problem": iters_rank / B, "parallelism": f"{world} x independent problem shards",
                       "l2": f"inputs {(A.numel() + y.numel()) * A.element_size() / 1e6:.0f} MB per step > 126 MB L2, no flush"},
            "roofline": roof if family == 3 else
                        {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC.get(args.config), "kernel": "tpp_lm_run_kernel" if family == 1 else "wpp_lm_run_kernel",
                         "kernel_ms": kernel_ms_avg, "algorithmic_bytes_per_launch": abytes, "peak_source": peak_src},
            "gpu_launches": int(launches), "clocks": clocks,
        }

    # e2e: the same solve through the C-ABI with HOST buffers (pinned), H2D + D2H inside
    if not args.no_e2e:
        # bounded to <= 4 GiB of pinned host memory (the whole batch for C2 / C3)
        Be = min(B, bounded_sample(cfg, 4 << 30))
        tiles = (Be + 31) // 32
        Asrc, ysrc = (A[:tiles], y[:tiles]) if layout == tb.TILE32 else (A[:Be], y[:Be])
        Ah = torch.empty(Asrc.shape, dtype=tdt, pin_memory=True); Ah.copy_(Asrc)
        yh = torch.empty(ysrc.shape, dtype=tdt, pin_memory=True); yh.copy_(ysrc)
        x0h = x0[:Be].cpu()
        xh = torch.empty(x0h.shape, dtype=tdt, pin_memory=True)
        rh = torch.empty((Be, tba.RESULT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
        rh_np = rh.numpy().view(tba.RESULT_DTYPE).reshape(-1)

        def e2e_step():
            xh.copy_(x0h)
            ctx.optimize_batch_host(Ah.numpy(), yh.numpy(), xh.numpy(), opt, alpha=ALPHA, layout=layout, B=Be, results=rh_np)

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(3, min(args.steps, 10))
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt_e2e = time.perf_counter() - t0
        te = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        if rank == 0:
            iters_e2e = int(rh_np["num_iters"].astype(np.int64).sum()) * world
            line["e2e"] = {"value": iters_e2e * ksteps / float(te.item()), "unit": "iterations/s",
                           "h2d_bytes_per_step": int((Ah.numel() + yh.numel() + xh.numel()) * Ah.element_size()),
                           "d2h_bytes_per_step": int(xh.numel() * xh.element_size() + rh.numel()),
                           "steps": ksteps, "problems_per_step": Be, "api": "tob200_lm_run_host (pinned host buffers)"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = min(B, bounded_sample(cfg, 1 << 30))
        iters, times, used = cpu_reference(cfg, Bc, 3)
        line["cpu_baseline"] = {"value": iters * len(times) / sum(times), "unit": "iterations/s", "cores": used, "kind": "port",
                                "sample": f"{Bc} of the {B} problems of {args.config} x {len(times)} repetitions, oracle port, OpenMP over problems"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
