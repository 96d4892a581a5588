#!/usr/bin/env python
"""bench.py — LM iterations/sec of the batched dense-NLLS hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (ours; N > 1 under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm: the oracle port,
                                                             all host threads, bounded samples)

One "step" = one complete batched LM solve (one tinyopt::Optimize() per problem) of a workload over
synthetic inputs; metric = sum over problems of Output::num_iters / time.  Prints ONE JSON line on
rank 0.  The headline (`value`, `roofline`, `e2e`, `cpu_baseline`) is C4 — the configuration
BASELINE.json quotes at 1/2/4/8 GPUs (1M problems, n = 50, m = 500, float; a FIXED total batch split
over the ranks: strong scaling) — and `configs` carries the same record for C2, C3 and C5
(`--config C2|C3|C4|C5` runs one of them alone as the headline).  Each record also holds two probes of
the path a tinyopt user with a host lambda hits: `build_solve` (materialised J, r; ONE launch; J read
once — the HBM probe of SURVEY.md §8d) and `solver_step` (the host-driven Step loop).  See DESIGN.md §6.
"""
import argparse
import ctypes
import gc
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

METRIC = "LM iterations/sec (batched dense NLLS)"
# BASELINE.json configs.  C2 / C3: the config's batch per GPU (weak); C4 / C5: a fixed
# total split over the ranks (strong), as BASELINE.json words them ("sharded 1/2/4/8 GPUs", "8 GPUs").
FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
CONFIGS = {
    "C2": dict(B=100_000, m=30, n=6, dtype="f64", opts={}, desc="batch 100k problems, n=6 params, 30 residuals each, double"),
    "C3": dict(B=100_000, m=200, n=12, dtype="f32", opts=FLOAT_OPTS, desc="batch 100k problems, n=12 params, 200 residuals each, float"),
    "C4": dict(B=1_000_000, m=500, n=50, dtype="f32", opts=FLOAT_OPTS, strong=True,
               desc="batch 1M problems, n=50 params, 500 residuals, float, sharded over the GPUs"),
    "C5": dict(B=4096, m=4096, n=512, dtype="f32", opts=FLOAT_OPTS, strong=True,
               desc="batch 4k problems, n=512 params, 4096 residuals, float, tensor-core JᵀJ tile, sharded over the GPUs"),
}
HEADLINE = "C4"
SEED, ALPHA, SIGMA = 20261017, 0.1, 1e-2


def config_of(name):
    """The `config` object of the JSON line: identical for our arm and the reference arm."""
    cfg = CONFIGS[name]
    return {"workload": f"{name}: {cfg['desc']}", "B": cfg["B"], "m": cfg["m"], "n": cfg["n"],
            "options": "tinyopt defaults" + (" + float thresholds min_rerr_dec=1e-5 min_step_norm2=1e-9" if cfg["opts"] else ""),
            "l2": "inputs larger than the 126 MB L2 (168 MB .. 102 GB per step), no flush"}


def elt(cfg):
    return 8 if cfg["dtype"] == "f64" else 4


def algorithmic_bytes(cfg, results):
    """s*(m*n + m) read + s*n written + 8 (cost) per problem-iteration that rebuilds
    H and g; a cost-only iteration needs only y (s*m) + 8."""
    s, m, n = elt(cfg), cfg["m"], cfg["n"]
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
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Median SM clock / reasons of the samples taken in [t0, t1] (all samples if none fall inside)."""
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def measured_peaks():
    """(HBM GB/s, source, bf16 burst TFLOP/s, source) from the driver-written MEASURED_PEAKS.json, else
    the fallbacks B200_PROFILING.md states."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return (float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d["bf16_tflops"]),
                "measured bf16 burst / 2 (MEASURED_PEAKS.json bf16_tflops; no TF32 figure there: tcgen05 K per instruction is "
                "32 bytes for every dtype, so TF32 runs at half the bf16 rate)")
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1590.0, "fallback bf16 1.59 PFLOP/s / 2 (B200_PROFILING.md)"


def ncu_traffic(name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, read from the
    committed `ncu --set full` capture digest (profiles/ncu_traffic.json, written by tools/ncu_summary.py
    from the .ncu-rep of the same command at the same size); None if no capture at the bench size exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get(name, {})
        if e.get("kernel") == kernel and e.get("problems") == e.get("bench_problems"):
            return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), e.get("source")
    except Exception:
        pass
    return None, None


def bounded_sample(cfg, budget_bytes, cap=None):
    """Problems of the workload whose inputs fit `budget_bytes` (a multiple of 32, at least 32, never
    more than `cap` or the config's batch): the bounded sample of the CPU / e2e / probe legs."""
    per = elt(cfg) * (cfg["m"] * cfg["n"] + cfg["m"] + 2 * cfg["n"])
    b = max(32, (budget_bytes // per) // 32 * 32)
    return int(min(cfg["B"], b, cap if cap else b))


def host_threads():
    """Every host core this process may run on — NOT OMP_NUM_THREADS, which torchrun pins to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---- the CPU arm: the oracle port (Eigen is not in the image, the reference cannot be compiled: DESIGN.md §3) ----
def cpu_run(cfg, sample_B, reps, fast, nthreads=0, warm=True):
    """iterations/s of the oracle over `sample_B` problems of the workload, OpenMP over problems.
    fast=True: the -DTOO_FAST build (same bits, rows in blocks of 8, vectorised; AVX-512 where the host has
    it); fast=False: the canonical one-serial-chain-at-a-time restatement."""
    from oracle import oracle as O
    nthreads = nthreads or host_threads()
    dt = np.float64 if cfg["dtype"] == "f64" else np.float32
    A, y, xs, x0 = O.synth_generate(sample_B, cfg["m"], cfg["n"], dt, seed=SEED, alpha=ALPHA, sigma=SIGMA)
    opt = O.default_options(**cfg["opts"])
    if warm:
        w = min(sample_B, max(nthreads, sample_B // 4))
        O.synth_lm_run(A[:w], y[:w], x0[:w], opt, alpha=ALPHA, nthreads=nthreads, fast=fast)
    times, iters, used = [], 0, 1
    for _ in range(reps):
        t0 = time.perf_counter()
        _, res, used = O.synth_lm_run(A, y, x0, opt, alpha=ALPHA, nthreads=nthreads, fast=fast)
        times.append(time.perf_counter() - t0)
        iters = int(res["num_iters"].astype(np.int64).sum())
    return iters, times, used


# problems per step of the CPU legs: about a second of work per step on a 16-core host
CPU_SAMPLE = {"C2": 100_000, "C3": 100_000, "C4": 10_464, "C5": 64}
CPU_SAMPLE_CANON = {"C2": 100_000, "C3": 50_000, "C4": 4_096, "C5": 16}


def cpu_baseline(name, reps=3, with_single_core=True):
    """cpu_baseline object: the fast build on all cores is `value`; the canonical (serial-chain) build and a
    1-core figure are reported beside it (SURVEY.md §8d asked for 1-core and all-core)."""
    from oracle import oracle as O
    cfg = CONFIGS[name]
    Bc = min(cfg["B"], CPU_SAMPLE[name])
    iters, times, used = cpu_run(cfg, Bc, reps, fast=True)
    value = iters * len(times) / sum(times)
    Bk = min(cfg["B"], CPU_SAMPLE_CANON[name])
    itk, tk, _ = cpu_run(cfg, Bk, 1, fast=False, warm=False)
    out = {"value": value, "unit": "iterations/s", "cores": used, "kind": "port",
           "sample": f"{Bc} of the {cfg['B']} problems of {name} x {len(times)} repetitions, OpenMP over problems, all host cores",
           "build": f"oracle -DTOO_FAST ({O.fast_variant()}: {'AVX-512' if O.fast_variant() == 'v4' else 'AVX2'}; rows in blocks of 8, interleaved "
                    "chains, vectorised accumulator loops; bit-identical to the canonical build)",
           "canonical_value": itk / sum(tk),
           "canonical_sample": f"{Bk} problems x 1, the canonical one-chain-at-a-time build (x86-64-v3, -ffp-contract=off)",
           "note": "a port of the reference's algorithm, not Eigen: Eigen's blocked GEMM / LDLT kernels would likely be faster "
                   "still on large n, so GPU/CPU ratios are upper bounds on the ratio against a real tinyopt build"}
    if with_single_core:
        B1 = max(32, min(Bc, Bc // max(1, used) * 2))
        it1, t1, _ = cpu_run(cfg, B1, 1, fast=True, nthreads=1, warm=False)
        out["single_core_value"] = it1 / sum(t1)
        out["single_core_sample"] = f"{B1} problems x 1, fast build, 1 thread"
    return out


def run_reference(args, names):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port, fast build, every
    host core) on the same configs / metric; each step is a bounded sample of the workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as O

    def one(name, steps, warmup):
        cfg = CONFIGS[name]
        Bc = min(cfg["B"], CPU_SAMPLE[name])
        if warmup > 0:
            cpu_run(cfg, Bc, min(warmup, 2), fast=True)
        iters, times, used = cpu_run(cfg, Bc, steps, fast=True, warm=False)
        total = sum(times)
        value = iters * steps / total
        return {"name": name, "config": config_of(name), "value": value, "unit": "iterations/s", "steps": steps,
                "ms_per_step": 1e3 * total / steps, "dtype": cfg["dtype"], "scaling": "strong" if cfg.get("strong") else "weak",
                "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": used, "kind": "port",
                                 "sample": f"{Bc} of the {cfg['B']} problems of {name} per step, {steps} steps, OpenMP over problems, "
                                           f"oracle -DTOO_FAST build ({O.fast_variant()})"},
                "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    head = names[-1]
    subs = [one(nm, max(1, min(args.steps, 5)), min(args.warmup, 1)) for nm in names[:-1]]
    h = one(head, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": h["value"], "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": h["ms_per_step"], "higher_is_better": True,
            "scaling": h["scaling"], "vs_baseline": None, "dtype": h["dtype"], "data": "synthetic", "config": h["config"],
            "cpu_baseline": h["cpu_baseline"], "e2e": h["e2e"], "gpu_launches": 0}
    if subs:
        line["configs"] = subs
    print(json.dumps(line), flush=True)


# ---- our arm --------------------------------------------------------------------------------------------------
class Pinned:
    """One pinned host arena shared by the e2e legs of every config (allocating pinned memory is slow)."""

    def __init__(self, nbytes):
        import torch
        self.buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        self.off = 0

    def reset(self):
        self.off = 0

    def take(self, shape, dtype):
        import torch
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.off = (self.off + 255) // 256 * 256
        if self.off + n > self.buf.numel():
            raise MemoryError("pinned arena too small")
        t = self.buf[self.off:self.off + n].view(dtype).view(*shape)
        self.off += n
        return t


def run_config(name, args, env, steps, warmup, sampler, pinned, headline):
    """One workload through tob200_lm_run_* on this rank's shard; returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist

    import tinyopt_b200 as tb
    from tinyopt_b200 import api as tba
    from tinyopt_b200.shard import shard_range

    cfg = CONFIGS[name]
    world, rank, local_rank, dev = env["world"], env["rank"], env["local_rank"], env["dev"]
    tdt = torch.float64 if cfg["dtype"] == "f64" else torch.float32
    m, n = cfg["m"], cfg["n"]
    strong = bool(cfg.get("strong"))
    if strong:   # a fixed total batch split over the ranks
        lo, hi = shard_range(cfg["B"], rank, world)
        per = -(-cfg["B"] // world)
        Btot = cfg["B"]
    else:        # the config's batch per GPU: weak scaling over independent problems
        lo, hi = rank * cfg["B"], (rank + 1) * cfg["B"]
        per = cfg["B"]
        Btot = cfg["B"] * world
    B = hi - lo
    ctx = tb.Context(local_rank)
    opt = tb.options(**cfg["opts"])
    family = ctx.kernel_family(tdt, n)
    layout = tb.TILE32 if family == 1 else tb.PROBLEM_MAJOR  # the family's native layout
    # rank g owns problems [lo, hi): generated in place from (seed, index), no scatter needed
    A, y, xs, x0 = ctx.synth_generate(B, m, n, tdt, p0=lo, seed=SEED, alpha=ALPHA, sigma=SIGMA, layout=layout)
    RB = tba.RESULT_DTYPE.itemsize
    res_buf = torch.zeros((per, RB), dtype=torch.uint8, device=dev)     # padded to the shard size: gathered as is
    gathered = torch.empty((per * world, RB), dtype=torch.uint8, device=dev) if world > 1 else None
    x = torch.empty_like(x0)
    lm_run = getattr(ctx._lib, f"tob200_lm_run_{cfg['dtype']}")
    c_alpha = (ctypes.c_double if cfg["dtype"] == "f64" else ctypes.c_float)(ALPHA)

    def solve():
        x.copy_(x0)
        ctx._ck(lm_run(ctx._h, ctypes.byref(opt), tba._p(A), tba._p(y), c_alpha, layout, B, m, n, tba._p(x),
                       tba._p(res_buf)), "tob200_lm_run")

    def step():
        solve()
        if world > 1:  # the one collective of the path: the gather of the per-problem results (NCCL)
            dist.all_gather_into_tensor(gathered, res_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    results = tba.decode_results(res_buf[:B])
    iters_rank = int(results["num_iters"].astype(np.int64).sum())

    launches0 = ctx.launch_count
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0

    # per-launch duration of the dominant kernel, CUDA events on the launching stream, live
    kms, tensor_ms, tensor_launches, phases = [], [], 0, None
    for _ in range(min(steps, 10 if family != 3 else 2)):
        solve()
        if family == 3:   # multi-kernel pipeline: the tensor-core JᵀJ launches of this solve, summed
            t_ms, tensor_launches = ctx.last_phase_ms(1)
            tensor_ms.append(t_ms)
            phases = {k: ctx.last_phase_ms(i)[0] for i, k in enumerate(("eval", "jtj", "solve"))}
            kms.append(sum(phases.values()))
        else:
            kms.append(ctx.last_elapsed_ms())
    kernel_ms_avg = float(np.mean(kms))

    # mid-n float runs take the tensor-core kernel (wtc.cuh, tolerance-held) by default: time the bit-exact
    # warp-per-problem kernel on the same inputs beside it (tob200_set_exact)
    wtc = (family == 2 and cfg["dtype"] == "f32" and os.environ.get("TOB200_WPP_TC", "1") != "0" and n >= 28 and m >= 192
           and (m * n) % 4 == 0)   # api.cu: lm_run_impl (kWtcMinN, kWtcMinM)
    exact_rec = None
    if wtc:
        ctx.set_exact(True)
        solve()
        ex_ms = []
        for _ in range(min(steps, 5)):
            solve()
            ex_ms.append(ctx.last_elapsed_ms())
        ctx.sync()
        ex_res = tba.decode_results(res_buf[:B])
        ctx.set_exact(False)
        ex_iters = int(ex_res["num_iters"].astype(np.int64).sum())
        exact_rec = {"kernel": "wpp_lm_run_kernel", "kernel_ms": float(np.mean(ex_ms)),
                     "value_this_gpu": ex_iters / (float(np.mean(ex_ms)) * 1e-3),
                     "same_num_iters_frac": float((ex_res["num_iters"] == results["num_iters"]).mean()),
                     "same_stop_reason_frac": float((ex_res["stop_reason"] == results["stop_reason"]).mean()),
                     "max_rel_final_cost_diff": float((np.abs(ex_res["final_cost"] - results["final_cost"])
                                                       / np.maximum(np.abs(ex_res["final_cost"]), 1e-300)).max()),
                     "note": "the bit-exact FFMA kernel (every sum the oracle's canonical fma chain) on the same shard, "
                             "device-timed per launch; the default kernel is held to the float tolerance of 1e-4 instead"}
        solve()   # leave the default kernel's results in res_buf
        ctx.sync()

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    it = torch.tensor([iters_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        got = tba.decode_results(gathered.view(world, per, RB)[rank, :B])
        assert np.array_equal(got["num_iters"], results["num_iters"])   # the gather carried this rank's rows
    ms_total = float(t.item())
    iters_all = float(it.item())
    value = iters_all * steps / (ms_total * 1e-3)

    rec = None
    if rank == 0:
        hbm_peak, hbm_src, bf16_peak, tf32_src = measured_peaks()
        abytes = algorithmic_bytes(cfg, results)
        achieved = abytes / (kernel_ms_avg * 1e-3) / 1e9
        s = elt(cfg)
        unique = B * (s * (m * n + m) + 2 * s * n + 64)
        if family == 3:
            # dominant kernel = lg_syrk_kernel (tcgen05): algorithmic flops m*n*(n+1) per rebuilt problem
            builds = int(results["num_builds"].astype(np.int64).sum())
            aflops = builds * m * n * (n + 1)
            t_ms = float(np.mean(tensor_ms))
            traffic, tsrc = ncu_traffic(name, "lg_syrk_kernel")
            fp16 = os.environ.get("TOB200_LG_FP16", "1") != "0"
            ach = aflops / (t_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": bf16_peak if fp16 else bf16_peak / 2, "unit": "TFLOP/s",
                    "traffic": traffic, "traffic_source": tsrc, "kernel": "lg_syrk_kernel", "kernel_ms": t_ms / max(1, tensor_launches),
                    "launches_per_step": tensor_launches, "algorithmic_flops_per_step": aflops,
                    "peak_source": ("measured bf16 burst (MEASURED_PEAKS.json bf16_tflops): the kernel issues tcgen05.mma.kind::f16"
                                    if fp16 else tf32_src),
                    "mode": ("3 x FP16 split (hi*hi + hi*lo + lo*hi on power-of-two scaled FP16 hi/lo parts, FP32 accumulate: FP32-level "
                             "accuracy)" if fp16 else "3xTF32") +
                            ": the hardware executes 3 (terms) x 1.25 (10 of 16 128x128 blocks for the n(n+1)/2 triangle) = 3.75x the "
                            "algorithmic flops",
                    "executed_frac": 3.75 * ach / (bf16_peak if fp16 else bf16_peak / 2),
                    "frac_vs_tf32_peak": ach / (bf16_peak / 2),
                    "pipeline_ms": phases, "hbm_GBps_whole_pipeline": achieved,
                    "whole_pipeline_frac_of_hbm": achieved / hbm_peak}
            roof["frac"] = roof["achieved"] / roof["peak"]
        else:
            kern = "tpp_lm_run_kernel" if family == 1 else ("wtc_lm_run_kernel" if wtc else "wpp_lm_run_kernel")
            traffic, tsrc = ncu_traffic(name, kern)
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "traffic_source": tsrc, "kernel": kern, "kernel_ms": kernel_ms_avg,
                    "algorithmic_bytes_per_launch": abytes, "peak_source": hbm_src,
                    # the whole loop is device resident: the inputs need to cross HBM only ONCE per launch if the
                    # later iterations hit on-chip copies; this is the bound against those unique bytes
                    "unique_bytes_per_launch": unique, "unique_GBps": unique / (kernel_ms_avg * 1e-3) / 1e9,
                    "unique_frac": unique / (kernel_ms_avg * 1e-3) / 1e9 / hbm_peak}
            if family == 2:  # C4 sits at the FP32 ridge: also report against the FFMA pipe
                builds = int(results["num_builds"].astype(np.int64).sum())
                flops = builds * (m * n * (n + 1) + 4 * m * n + n * n * n / 3 + 2 * n * n)
                fp32_peak = 148 * 128 * 2 * (sampler.window(tw0, tw1)["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
                roof["fp32"] = {"achieved_TFLOPs": flops / (kernel_ms_avg * 1e-3) / 1e12, "peak_TFLOPs": fp32_peak,
                                "frac": flops / (kernel_ms_avg * 1e-3) / 1e12 / fp32_peak,
                                "peak_source": "148 SMs x 128 FFMA lanes x 2 flop x max SM clock"}
                if wtc:
                    # J^T J runs on tcgen05.mma.kind::f16: per pair of problems and 16 rows three 128 x 128 x 16 MMAs
                    # (hi hi' + hi lo' + lo hi'), half of whose outputs (the two diagonal 64 x 64 blocks) are used
                    executed = builds / 2 * (-(-m // 32) * 2) * 3 * 2 * 128 * 128 * 16
                    roof["fp32"]["note"] = "algorithmic FP32 flops of the path (what an FFMA kernel would execute), for comparison only"
                    roof["tensor"] = {"executed_TFLOPs": executed / (kernel_ms_avg * 1e-3) / 1e12, "peak_TFLOPs": bf16_peak,
                                      "frac": executed / (kernel_ms_avg * 1e-3) / 1e12 / bf16_peak,
                                      "note": "executed tcgen05 flops (3-term FP16 split, 128 x 128 tiles per pair of problems) against the "
                                              "measured bf16 peak: the kernel is bound by shared-memory bandwidth, not by the tensor pipe"}
                    roof["parity"] = "tolerance-held: x / cost within 1e-4, iteration counts identical where decisions clear FP32 noise"
                    roof["exact_kernel"] = exact_rec
        rec = {"name": name, "config": config_of(name), "value": value, "unit": "iterations/s", "steps": steps, "warmup": warmup,
               "ms_per_step": ms_total / steps, "dtype": cfg["dtype"], "scaling": "strong" if strong else "weak",
               "roofline": roof, "gpu_launches": int(launches), "clocks": sampler.window(tw0, tw1),
               "run": {"B_per_gpu": B, "B_total": Btot, "iters_per_problem": iters_rank / max(1, B),
                       "parallelism": f"{world} x independent problem shards; results all-gathered inside every timed step"
                       if world > 1 else "1 GPU",
                       "input_MB_per_gpu": (A.numel() + y.numel()) * A.element_size() / 1e6}}

    # ---- e2e: the same solve through the C-ABI with HOST buffers (pinned), H2D + D2H inside -------------------
    if not args.no_e2e:
        pinned.reset()
        Be = min(B, bounded_sample(cfg, 4 << 30))   # <= 4 GiB of pinned host memory (the whole batch for C2 / C3)
        tiles = (Be + 31) // 32
        Asrc, ysrc = (A[:tiles], y[:tiles]) if layout == tb.TILE32 else (A[:Be], y[:Be])
        Ah = pinned.take(Asrc.shape, tdt); Ah.copy_(Asrc)
        yh = pinned.take(ysrc.shape, tdt); yh.copy_(ysrc)
        x0h = x0[:Be].cpu()
        xh = pinned.take(x0h.shape, tdt)
        rh = pinned.take((Be, RB), torch.uint8)
        rh_np = rh.numpy().view(tba.RESULT_DTYPE).reshape(-1)

        def e2e_step():
            xh.copy_(x0h)
            ctx.optimize_batch_host(Ah.numpy(), yh.numpy(), xh.numpy(), opt, alpha=ALPHA, layout=layout, B=Be, results=rh_np)

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(3, min(steps, 10))
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt_e2e = time.perf_counter() - t0
        te = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        ie = torch.tensor([float(rh_np["num_iters"].astype(np.int64).sum())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ie, op=dist.ReduceOp.SUM)
        if rank == 0:
            rec["e2e"] = {"value": float(ie.item()) * ksteps / float(te.item()), "unit": "iterations/s",
                          "h2d_bytes_per_step": int((Ah.numel() + yh.numel() + xh.numel()) * Ah.element_size()),
                          "d2h_bytes_per_step": int(xh.numel() * xh.element_size() + rh.numel()),
                          "steps": ksteps, "problems_per_step": Be, "api": "tob200_lm_run_host (pinned host buffers)"}
        del Ah, yh, xh, rh

    # ---- probes: the materialised-J path (what a tinyopt user with a host lambda hits) -------------------------
    if rank == 0 and not args.no_probes:
        rec["probes"] = probes(ctx, name, cfg, A, y, x0, layout, family, tdt, opt)

    del A, y, xs, x0, x, res_buf, gathered
    ctx.close()
    gc.collect()
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rec["cpu_baseline"] = cpu_baseline(name, reps=3 if headline else 2)
    return rec


def probes(ctx, name, cfg, A, y, x0, layout, family, tdt, opt):
    """(1) tob200_build_solve_*: ONE Build + Solve from materialised J, r in HBM: J is read exactly once, so
    algorithmic bytes == the bytes the kernel must move: the HBM roofline probe of SURVEY.md §8(d).
    (2) tob200_solver_step_*: the SolverType seam driven from the host (n <= 55): per LM iteration the caller's
    residual evaluation (here tob200_synth_eval on the device) writes J, r and the library does Build + damping +
    Solve + the Step bookkeeping; iterations/s of the whole loop."""
    import torch

    import tinyopt_b200 as tb
    from tinyopt_b200 import api as tba
    m, n, s = cfg["m"], cfg["n"], elt(cfg)
    hbm_peak, hbm_src, _, _ = measured_peaks()
    Bp = min(int(x0.shape[0]), bounded_sample(cfg, 16 << 30))   # this rank's shard; J is a second copy of A's size: A + J <= 32 GiB
    tiles = (Bp + 31) // 32
    Ap, yp = (A[:tiles], y[:tiles]) if layout == tb.TILE32 else (A[:Bp], y[:Bp])
    xp = x0[:Bp].contiguous()
    out = {}
    r, J = ctx.synth_eval(Ap, yp, xp, alpha=ALPHA, layout=layout)
    lam = torch.full((Bp,), 1e-4, dtype=tdt, device=xp.device)
    for _ in range(2):
        ctx.build_solve(J, r, lam, B=Bp, layout=layout)
    ms = []
    for _ in range(5):
        ctx.build_solve(J, r, lam, B=Bp, layout=layout)
        ms.append(ctx.last_elapsed_ms())
    t = float(np.mean(ms)) * 1e-3
    ab = Bp * (s * (m * n + m) + s * n + 8)
    out["build_solve"] = {"api": "tob200_build_solve (materialised J, r; one launch" + ("; eval + JᵀJ + solve kernels" if family == 3 else "") + ")",
                          "problems": Bp, "ms": t * 1e3, "problems_per_s": Bp / t, "algorithmic_bytes": ab,
                          "achieved_GBps": ab / t / 1e9, "peak_GBps": hbm_peak, "frac": ab / t / 1e9 / hbm_peak}
    if family != 3:
        solver = tba.BatchSolver(ctx, Bp, n, tdt, opt)
        t_best, iters = None, 0
        for rep in range(3):
            solver.reset(xp)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            passes = 0
            while True:
                r, J = ctx.synth_eval(Ap, yp, solver.x, alpha=ALPHA, layout=layout)   # the user's lambda
                solver.step(J, r, layout=layout)
                passes += 1
                if solver.num_active() == 0 or passes > 200:
                    break
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep == 0:
                iters = int(solver.results()["num_iters"].astype(np.int64).sum())
            t_best = dt if t_best is None else min(t_best, dt)
        solver.close()
        out["solver_step"] = {"api": "tob200_solver_step (host-driven Step loop; J, r from tob200_synth_eval each pass)",
                              "problems": Bp, "passes": passes, "ms": t_best * 1e3, "value": iters / t_best, "unit": "iterations/s"}
    del r, J
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="all", choices=sorted(CONFIGS) + ["all"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-probes", action="store_true")
    args = ap.parse_args()
    # sub-records first, the headline last (it is the one timed with exactly --steps / --warmup)
    names = [c for c in ("C2", "C3", "C5") if args.config == "all"] + [HEADLINE if args.config == "all" else args.config]
    if args.impl == "reference":
        return run_reference(args, names)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    env = dict(world=world, rank=rank, local_rank=local_rank, dev=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    pinned = None if args.no_e2e else Pinned((4 << 30) + (64 << 20))

    recs = []
    for nm in names:
        head = nm == names[-1]
        steps = args.steps if head else max(3, min(args.steps, 5 if nm == "C5" else 10))
        recs.append(run_config(nm, args, env, steps, args.warmup if head else 3, sampler, pinned, head))
    sampler.stop()
    if rank == 0:
        h = recs[-1]
        line = {"metric": METRIC, "value": h["value"], "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": h["scaling"],
                "vs_baseline": None, "dtype": h["dtype"], "data": "synthetic", "config": h["config"], "run": h["run"],
                "roofline": h["roofline"], "gpu_launches": h["gpu_launches"], "clocks": h["clocks"]}
        for k in ("e2e", "cpu_baseline", "probes"):
            if k in h:
                line[k] = h[k]
        if len(recs) > 1:
            line["configs"] = recs[:-1]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
