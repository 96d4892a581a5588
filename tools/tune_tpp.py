"""Kernel-time sweep of the thread-per-problem LM kernel over its launch knobs (GPU box only).
usage: python tools/tune_tpp.py C2|C3 [B]"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyopt_b200 as tb  # noqa: E402
from bench import CONFIGS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = CONFIGS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["B"]
tdt = torch.float64 if cfg["dtype"] == "f64" else torch.float32
family2 = cfg["n"] > 12
grid_stages = [int(v) for v in os.environ.get("SWEEP_STAGES", "2,3,4").split(",")]
grid_bytes = [int(v) for v in os.environ.get("SWEEP_BYTES", "2048,4096,8192,16384").split(",")]
grid_ctas = [int(v) for v in os.environ.get("SWEEP_CTAS", "0").split(",")]
data = None
for st, sb, ct in itertools.product(grid_stages, grid_bytes, grid_ctas):
    os.environ["TOB200_TPP_STAGES"] = str(st)
    os.environ["TOB200_WPP_STAGES"] = str(st)
    os.environ["TOB200_TPP_STAGE_BYTES"] = str(sb)
    os.environ["TOB200_TPP_CTAS_PER_SM"] = str(ct)
    ctx = tb.Context(0)
    if data is None:
        data = ctx.synth_generate(B, cfg["m"], cfg["n"], tdt, layout=tb.PROBLEM_MAJOR if family2 else tb.TILE32)
    A, y, xs, x0 = data
    opt = tb.options(**cfg["opts"])
    ms = []
    for i in range(8):
        out = ctx.optimize_batch(A, y, x0, opt)
        ms.append(ctx.last_elapsed_ms())
    it = int(out.results["num_iters"].astype(np.int64).sum())
    best = min(ms[2:])
    print(f"{name} stages={st} stage_bytes<={sb} ctas/sm={ct}: kernel {best:.4f} ms  {it / best / 1e6:.1f} G it/s... "
          f"({it / (best * 1e-3) / 1e9:.3f} G it/s)", flush=True)
    ctx.close()
