"""One batched LM solve of a bench config (for ncu captures): python tools/run_once.py C4 4096 [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyopt_b200 as tb  # noqa: E402
from bench import CONFIGS  # noqa: E402

name = sys.argv[1]
cfg = CONFIGS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["B"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tdt = torch.float64 if cfg["dtype"] == "f64" else torch.float32
ctx = tb.Context(0)
A, y, xs, x0 = ctx.synth_generate(B, cfg["m"], cfg["n"], tdt, layout=tb.PROBLEM_MAJOR if cfg["n"] > 12 else tb.TILE32)
for _ in range(reps):
    out = ctx.optimize_batch(A, y, x0, tb.options(**cfg["opts"]))
iters = int(out.results["num_iters"].sum())
ms = ctx.last_elapsed_ms()
print(name, B, "iters", iters, "kernel ms", ms, "Mit/s", iters / ms / 1e3)
if ctx.kernel_family(tdt, cfg["n"]) == 3:
    print("  phases (ms, launches): eval", ctx.last_phase_ms(0), "jtj", ctx.last_phase_ms(1), "solve", ctx.last_phase_ms(2))
