"""Regenerates tests/golden/synth_family.npz from the CPU oracle (python tests/golden/make_golden.py).
The fixture pins the synthetic problem family and the oracle's LM results on the
first 8 problems of the C2 and C3 shapes, so that neither can drift unnoticed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

out = {}
for tag, dtype, (B, m, n), kw in (("c2", np.float64, (8, 30, 6), {}),
                                  ("c3", np.float32, (8, 200, 12), dict(min_rerr_dec=1e-5, min_step_norm2=1e-9))):
    A, y, xs, x0 = O.synth_generate(B, m, n, dtype)
    x, res, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), nthreads=1)
    out.update({f"{tag}_A": A, f"{tag}_y": y, f"{tag}_xstar": xs, f"{tag}_x0": x0, f"{tag}_x": x,
                f"{tag}_num_iters": res["num_iters"], f"{tag}_stop_reason": res["stop_reason"],
                f"{tag}_final_cost": res["final_cost"]})
np.savez_compressed(os.path.join(HERE, "synth_family.npz"), **out)
print("wrote", os.path.join(HERE, "synth_family.npz"))
