#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 120 tests/cuda/build/test_device_functor
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import tinyopt_b200 as tb
from oracle import oracle as O
ctx = tb.Context(0)
kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
A, y, xs, x0 = O.synth_generate(4096, 200, 12, np.float32)
xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw))
for layout in (tb.TILE32, tb.PROBLEM_MAJOR):
    dA, dy, _, dx0 = ctx.synth_generate(4096, 200, 12, torch.float32, layout=layout)
    out = ctx.optimize_batch(dA, dy, dx0, tb.options(**kw), layout=layout)
    print("layout", layout, "iters equal", np.array_equal(out.results["num_iters"], ro["num_iters"]), "x equal", np.array_equal(out.x.cpu().numpy(), xo),
          out.results["num_iters"][:4], ro["num_iters"][:4], out.results["final_cost"][:2], ro["final_cost"][:2])
PY
