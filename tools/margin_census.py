#!/usr/bin/env python
"""Decision-margin census of the benchmark workloads (SURVEY.md §7 hard part 2ii).

Eigen is not in the image, so "iteration counts identical to the reference" cannot be checked against
real Eigen output: its GEMM sums the residual rows in an implementation-defined order.  What CAN be
bounded is how many problems could possibly notice.  For every problem the oracle records

  sign_margin = min over Steps of |cost_k - cost_{k-1}| / |cost_k|      (accept / reject, optimizer.h:429)
  thr_margin  = min over Steps, stop tests of |v - threshold| / threshold   (optimizer.h:518-528)

A different summation order moves a cost by ~m*eps relative, so a problem with sign_margin above
1e-12 (double) / 2e-5 (float) and thr_margin above 0.1 takes the same branches under ANY backward-
stable order.  The census also does the experiment directly: it re-runs every problem with the rows
summed in REVERSE order (`reverse_rows`, a stand-in for "Eigen's order") and counts how many change
their iteration count or stop reason, and how far x moves.

Float workloads are reported twice: decisions of the float run itself (what a float tinyopt does),
and margins from a double run on the same float inputs (is the decision well defined at all).

  python tools/margin_census.py [--out profiles/r2_margin_census.json] [--c4 10240] [--c5 4096]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

FLOAT_OPTS = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9)
CONFIGS = {
    "C2": dict(B=100_000, m=30, n=6, dtype=np.float64, opts={}),
    "C3": dict(B=100_000, m=200, n=12, dtype=np.float32, opts=FLOAT_OPTS),
    "C4": dict(B=1_000_000, m=500, n=50, dtype=np.float32, opts=FLOAT_OPTS),
    "C5": dict(B=4096, m=4096, n=512, dtype=np.float32, opts=FLOAT_OPTS),
}


def census(name, cfg, sample, chunk_bytes=1 << 30):
    dt, m, n = cfg["dtype"], cfg["m"], cfg["n"]
    is_f32 = dt == np.float32
    sign_bar = 2e-5 if is_f32 else 1e-12
    opt = O.default_options(**cfg["opts"])
    per = (m * n + m + 2 * n) * np.dtype(dt).itemsize
    chunk = max(32, min(sample, chunk_bytes // per))
    acc = dict(n=0, robust=0, same_rev=0, same_rev_robust=0, same_f64=0, same_it_rev=0, same_it_f64=0, max_dx_rev=0.0, iters=0,
               hist={}, stops={}, sign_q=[], thr_q=[])
    t0 = time.perf_counter()
    for p0 in range(0, sample, chunk):
        b = min(chunk, sample - p0)
        A, y, xs, x0 = O.synth_generate(b, m, n, dt, p0=p0)
        x1, r1, _ = O.synth_lm_run(A, y, x0, opt, fast=True)                      # the canonical run (same bits as lib())
        x2, r2, _ = O.synth_lm_run(A, y, x0, opt, fast=True, reverse_rows=True)   # another summation order
        if is_f32:  # margins: the double run on the same float inputs
            _, r64, _ = O.synth_lm_run(A.astype(np.float64), y.astype(np.float64), x0.astype(np.float64), opt, fast=True)
            acc["same_f64"] += int(((r64["num_iters"] == r1["num_iters"]) & (r64["stop_reason"] == r1["stop_reason"])).sum())
            acc["same_it_f64"] += int((r64["num_iters"] == r1["num_iters"]).sum())
        else:
            r64 = r1
        robust = (r64["sign_margin"] > sign_bar) & (r64["thr_margin"] > 0.1)
        same = (r1["num_iters"] == r2["num_iters"]) & (r1["stop_reason"] == r2["stop_reason"])
        acc["n"] += b
        acc["robust"] += int(robust.sum())
        acc["same_rev"] += int(same.sum())
        acc["same_it_rev"] += int((r1["num_iters"] == r2["num_iters"]).sum())
        acc["same_rev_robust"] += int((same & robust).sum())
        acc["iters"] += int(r1["num_iters"].sum())
        den = np.maximum(np.abs(x1).max(axis=1), 1e-300)
        d = (np.abs(x1.astype(np.float64) - x2).max(axis=1) / den)[same]
        if d.size:
            acc["max_dx_rev"] = max(acc["max_dx_rev"], float(d.max()))
        for v, c in zip(*np.unique(r1["num_iters"], return_counts=True)):
            acc["hist"][int(v)] = acc["hist"].get(int(v), 0) + int(c)
        for v, c in zip(*np.unique(r1["stop_reason"], return_counts=True)):
            acc["stops"][O.STOP[int(v)]] = acc["stops"].get(O.STOP[int(v)], 0) + int(c)
        acc["sign_q"].append(r64["sign_margin"]); acc["thr_q"].append(r64["thr_margin"])
        print(f"  {name}: {acc['n']}/{sample} problems, {time.perf_counter() - t0:.0f} s", file=sys.stderr, flush=True)
    sq, tq = np.concatenate(acc["sign_q"]), np.concatenate(acc["thr_q"])
    N = acc["n"]
    out = {
        "config": name, "problems": N, "of": cfg["B"], "m": m, "n": n, "dtype": np.dtype(dt).name,
        "options": "tinyopt defaults" + (" + min_rerr_dec=1e-5 min_step_norm2=1e-9" if cfg["opts"] else ""),
        "iters_per_problem": acc["iters"] / N, "num_iters_histogram": acc["hist"], "stop_reasons": acc["stops"],
        "sign_margin_bar": sign_bar, "thr_margin_bar": 0.1,
        "margins_from": "double run on the same float inputs" if is_f32 else "the run itself",
        "robust_fraction": acc["robust"] / N,
        "same_num_iters_under_reversed_row_order": acc["same_it_rev"] / N,
        "same_decisions_under_reversed_row_order": acc["same_rev"] / N,
        "robust_and_same_under_reversed_order": (acc["same_rev_robust"] / acc["robust"]) if acc["robust"] else 1.0,
        "max_rel_dx_under_reversed_order": acc["max_dx_rev"],
        "sign_margin_quantiles": {q: float(np.quantile(sq, float(q))) for q in ("0.001", "0.01", "0.1", "0.5")},
        "thr_margin_quantiles": {q: float(np.quantile(tq, float(q))) for q in ("0.001", "0.01", "0.1", "0.5")},
        "seconds": time.perf_counter() - t0,
    }
    if is_f32:
        out["float_run_equals_double_run"] = acc["same_f64"] / N
        out["float_run_num_iters_equal_double_run"] = acc["same_it_f64"] / N
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_margin_census.json"))
    ap.add_argument("--c2", type=int, default=100_000)
    ap.add_argument("--c3", type=int, default=100_000)
    ap.add_argument("--c4", type=int, default=10_240)
    ap.add_argument("--c5", type=int, default=4096)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    sizes = {"C2": a.c2, "C3": a.c3, "C4": a.c4, "C5": a.c5}
    rows = []
    for name, cfg in CONFIGS.items():
        if a.only and name not in a.only.split(","):
            continue
        if sizes[name] <= 0:
            continue
        rows.append(census(name, cfg, min(sizes[name], cfg["B"])))
        print(json.dumps(rows[-1]), flush=True)
    with open(a.out, "w") as f:
        json.dump({"tool": "tools/margin_census.py", "oracle_threads": O.max_threads(), "rows": rows}, f, indent=1)
    print("| config | problems | robust | same num_iters, reversed row order | same num_iters + stop reason, reversed | robust ∧ same | "
          "float run == double run (iters / iters + stop) | max rel Δx (reversed) |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['config']} | {r['problems']} of {r['of']} | {r['robust_fraction']:.4f} | {r['same_num_iters_under_reversed_row_order']:.4f} | "
              f"{r['same_decisions_under_reversed_row_order']:.4f} | {r['robust_and_same_under_reversed_order']:.4f} | "
              f"{r.get('float_run_num_iters_equal_double_run', float('nan')):.4f} / {r.get('float_run_equals_double_run', float('nan')):.4f} | "
              f"{r['max_rel_dx_under_reversed_order']:.1e} |")


if __name__ == "__main__":
    main()
