"""Decision-margin census (tools/margin_census.py): how many problems of the benchmark workloads could possibly
notice that Eigen sums the residual rows in another order than the oracle.  Eigen is not in the image, so this is
the only defensible bound on "iteration counts identical to the reference" (optimizer.h:429,518-528).  The test
(a) re-runs a slice of every config and asserts the bounds, (b) checks the committed full-batch artefact
profiles/r2_margin_census.json (C2 and C3 whole, 10240 problems of C4, all 4096 of C5)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import margin_census as MC  # noqa: E402


def test_census_slice():
    rows = {nm: MC.census(nm, MC.CONFIGS[nm], sz) for nm, sz in (("C2", 4000), ("C3", 2000), ("C4", 256), ("C5", 8))}
    for nm, r in rows.items():
        # another summation order never changes the iteration COUNT on these slices ...
        assert r["same_num_iters_under_reversed_row_order"] == 1.0, (nm, r)
        # ... and moves the solution by rounding only
        assert r["max_rel_dx_under_reversed_order"] < (1e-14 if nm == "C2" else 2e-6), (nm, r)
        # every problem whose margins clear the bars takes exactly the same decisions
        assert r["robust_and_same_under_reversed_order"] == 1.0
    assert rows["C2"]["robust_fraction"] > 0.95 and rows["C3"]["robust_fraction"] > 0.98 and rows["C4"]["robust_fraction"] > 0.95
    # C5: the 4th Step's cost change is below FP32 resolution for (nearly) every problem (see
    # tests/test_gpu_large.py::test_lg_lm_run_parity_real_c5_shape): the margin-robust set is (nearly) empty,
    # the stop REASON is noise, the iteration count is not
    assert rows["C5"]["robust_fraction"] < 0.5


def test_committed_census_artefact():
    with open(os.path.join(ROOT, "profiles", "r2_margin_census.json")) as f:
        d = json.load(f)
    rows = {r["config"]: r for r in d["rows"]}
    assert set(rows) == {"C2", "C3", "C4", "C5"}
    assert rows["C2"]["problems"] == 100_000 and rows["C3"]["problems"] == 100_000
    assert rows["C4"]["problems"] >= 10_000 and rows["C5"]["problems"] == 4096
    for nm in ("C2", "C3", "C4"):
        assert rows[nm]["robust_fraction"] > 0.98
        assert rows[nm]["same_num_iters_under_reversed_row_order"] > 0.9999
        assert rows[nm]["robust_and_same_under_reversed_order"] == 1.0
    assert rows["C5"]["same_num_iters_under_reversed_row_order"] > 0.999
    assert rows["C5"]["float_run_num_iters_equal_double_run"] > 0.998
