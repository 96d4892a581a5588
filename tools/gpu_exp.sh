#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | tail -15
timeout 120 python tools/run_once.py C4 131072 3 2>&1 | tail -1
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import tinyopt_b200 as tb
ctx = tb.Context(0)
for (B,m,n) in ((100000,100,12),(100000,500,50)):
    A,y,xs,x0 = ctx.synth_generate(B,m,n,torch.float64,layout=tb.PROBLEM_MAJOR)
    for _ in range(3):
        out = ctx.optimize_batch(A,y,x0,tb.options())
    it = int(out.results["num_iters"].sum()); ms = ctx.last_elapsed_ms()
    print("double", B, m, n, "iters", it, "ms", ms, "M it/s", it/ms/1e3)
PY
