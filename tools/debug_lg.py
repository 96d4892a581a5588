import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyopt_b200 as tb
from oracle import oracle as O
ctx = tb.Context(0)
np.set_printoptions(linewidth=250, precision=6)
FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
for (B, m, n, kw) in [(3, 1500, 320, {}), (3, 2048, 512, {}), (8, 300, 64, dict(damping_init=10.0, max_consec_failures=2)), (150, 300, 60, {})]:
    kw = {**FLOAT_OPTS, **kw}
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float32)
    x32, r32, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
    x64, r64, _ = O.synth_lm_run(A.astype(np.float64), y.astype(np.float64), x0.astype(np.float64), O.default_options(**kw))
    dA, dy, _, dx0 = ctx.synth_generate(B, m, n, torch.float32, layout=tb.PROBLEM_MAJOR)
    out = ctx.optimize_batch(dA, dy, dx0, tb.options(**kw), layout=tb.PROBLEM_MAJOR)
    rg = out.results
    print(f"=== B={B} m={m} n={n} {kw}")
    idx = np.nonzero((rg["num_iters"] != r32["num_iters"]) | (rg["stop_reason"] != r32["stop_reason"]))[0]
    print("mismatching problems:", idx[:10], "of", B)
    for p in list(idx[:4]) + [0]:
        for name, r in (("o32", r32), ("o64", r64), ("gpu", rg)):
            print(f"  p={p} {name}: iters={r['num_iters'][p]} stop={r['stop_reason'][p]} fails={r['num_failures'][p]} cost={r['final_cost'][p]:.9g} rerr={r['final_rerr_dec'][p]:.3e} lam={r['last_lambda'][p]:.4g}" + (f" margin={r['min_margin'][p]:.2e}" if 'min_margin' in r.dtype.names else "") + (f" builds={r['num_builds'][p]}" if 'num_builds' in r.dtype.names else ""))
    print("  x rel err gpu vs o32:", np.abs(out.x.cpu().numpy() - x32).max() / np.abs(x32).max(), " o64 vs o32:", np.abs(x64 - x32).max() / np.abs(x32).max())
