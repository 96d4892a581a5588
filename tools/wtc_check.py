"""Mid-n tensor-core kernel (wtc.cuh) against the bit-exact warp-per-problem kernel on the same device inputs:
python tools/wtc_check.py [B] [m] [n] [reps]      (prints agreement statistics and both timings)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyopt_b200 as tb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else 500
n = int(sys.argv[3]) if len(sys.argv) > 3 else 50
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ctx = tb.Context(0)
A, y, xs, x0 = ctx.synth_generate(B, m, n, torch.float32, layout=tb.PROBLEM_MAJOR)
opt = tb.options(min_rerr_dec=1e-5, min_step_norm2=1e-9) if os.environ.get('FLOAT_OPTS', '1') == '1' else tb.options()


def run(exact):
    ctx.set_exact(exact)
    for _ in range(reps):
        out = ctx.optimize_batch(A, y, x0, opt)
    ms = ctx.last_elapsed_ms()
    return out, ms


ref, ms_ref = run(True)
print(f"exact : {B} x m={m} n={n}: {ms_ref:.3f} ms, iters {int(ref.results['num_iters'].sum())}, "
      f"{ref.results['num_iters'].sum() / ms_ref / 1e3:.2f} M it/s", flush=True)
out, ms = run(False)
it = int(out.results["num_iters"].sum())
print(f"tc    : {ms:.3f} ms, iters {it}, {it / ms / 1e3:.2f} M it/s  (x{ms_ref / ms:.2f})", flush=True)
xr, xt = ref.x.double().cpu().numpy(), out.x.double().cpu().numpy()
rel = np.abs(xr - xt).max(axis=1) / np.maximum(np.abs(xr).max(axis=1), 1e-30)
same_it = ref.results["num_iters"] == out.results["num_iters"]
same_stop = ref.results["stop_reason"] == out.results["stop_reason"]
cr, ct = ref.results["final_cost"], out.results["final_cost"]
crel = np.abs(cr - ct) / np.maximum(np.abs(cr), 1e-30)
print(f"x rel err: max {rel.max():.3e} median {np.median(rel):.3e}; same iters {same_it.mean():.4f}; "
      f"same stop {same_stop.mean():.4f}; cost rel err max {crel.max():.3e}")
print("stop reasons exact:", np.unique(ref.results["stop_reason"], return_counts=True))
print("stop reasons tc   :", np.unique(out.results["stop_reason"], return_counts=True))
bad = np.where(~np.isfinite(xt).all(axis=1))[0]
if len(bad):
    print("non-finite x in", len(bad), "problems, first", bad[:8])
worst = np.argsort(-rel)[:5]
for w in worst:
    print("  problem", w, "rel", rel[w], "iters", ref.results["num_iters"][w], out.results["num_iters"][w],
          "stop", ref.results["stop_reason"][w], out.results["stop_reason"][w], "cost", cr[w], ct[w])
