"""SURVEY.md §8(f) rank 1 — the user's residual functor on the device
(include/tinyopt_b200_device.cuh): tests/cuda/test_device_functor.cu is built the way a tinyopt user
would build it (nvcc, sm_100a) and checks the sqrt2 golden vector through a Jet functor, that the
polynomial family written as a user functor equals tob200_lm_run_* (oracle-pinned by
test_gpu_parity.py) bit for bit, and that the Jet (autodiff) variant takes the same decisions."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CU = os.path.join(ROOT, "tests", "cuda")
EXE = os.path.join(CU, "build", "test_device_functor")


def build_exe():
    subprocess.run(["make", "-C", CU], check=True, capture_output=True)
    assert os.path.exists(EXE)


def test_device_functor_builds_and_fails_loudly_without_gpu():
    import torch
    build_exe()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by test_device_functor_on_gpu")
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 3, p.stdout + p.stderr
    assert "no CPU fallback" in p.stdout
    # the M-estimators of the header are host arithmetic too: pinned by tests/robust_norms.cpp's closed forms
    assert "robust norms: 7 M-estimators x 3 points: closed forms and scale == Jet derivative: ok" in p.stdout


@pytest.mark.gpu
def test_device_functor_on_gpu(tmp_path):
    build_exe()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=300,
                       env={**os.environ, "TOB200_FUNCTOR_DUMP": str(tmp_path)})
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all device functor checks passed" in p.stdout
    assert "closed forms and scale == Jet derivative: ok" in p.stdout
    assert "sqrt2 (Jet functor): x[0]=1.414213562373095" in p.stdout and "iters=5 stop=1" in p.stdout

    # Jets (autodiff) round differently from the closed-form Jacobian, so a decision that sits within FP32
    # rounding of its threshold may fall either way; every problem whose decisions CLEAR that noise (the
    # oracle's margins, double run on the same float inputs: tests/test_gpu_large.py::robust_decisions)
    # must take exactly the decisions of tob200_lm_run, in both precisions.
    import re

    import numpy as np

    from oracle import oracle as O
    dumps = sorted(tmp_path.glob("jets_*.bin"))
    assert len(dumps) >= 9, dumps
    for f in dumps:
        dt, n, m, B, _, solve = re.match(r"jets_(f\d+)_n(\d+)_m(\d+)_B(\d+)_(\w+)_(\w+)\.bin", f.name).groups()
        n, m, B = int(n), int(m), int(B)
        rec = np.fromfile(f, np.int32).reshape(B, 4)
        kw = dict(use_ldlt=1 if solve == "ldlt" else 0)
        if dt == "f32":
            kw.update(min_rerr_dec=1e-5, min_step_norm2=1e-9)
        A, y, xs, x0 = O.synth_generate(B, m, n, np.float32 if dt == "f32" else np.float64)
        _, r64, _ = O.synth_lm_run(A.astype(np.float64), y.astype(np.float64), x0.astype(np.float64), O.default_options(**kw))
        robust = (r64["sign_margin"] > (2e-5 if dt == "f32" else 1e-12)) & (r64["thr_margin"] > 0.1)
        assert robust.mean() >= 0.85, (f.name, robust.mean())
        same = (rec[:, 0] == rec[:, 2]) & (rec[:, 1] == rec[:, 3])
        assert same[robust].all(), (f.name, int((~same & robust).sum()))

    # robust re-weighting inside the accumulation, ON THE DEVICE, against the oracle's robust
    # variant of the family.  M-estimators built from IEEE +, -, *, /, sqrt only (Truncated, Huber, Tukey,
    # Geman-McClure) must agree bit for bit - solution, iteration count, stop reason, failures, final cost; the
    # ones through atan / log / exp (Arctan, Cauchy, Blake-Zisserman: libm and CUDA round differently) to 1e-9
    # relative on the problems that take the same number of Steps.
    rdumps = sorted(tmp_path.glob("robust_*.bin"))
    assert len(rdumps) == 11, rdumps
    for f in rdumps:
        kind, dt, n, m, B, fam = re.match(r"robust_k(\d)_(f\d+)_n(\d+)_m(\d+)_B(\d+)_(\w+)\.bin", f.name).groups()
        kind, n, m, B = int(kind), int(n), int(m), int(B)
        rec = np.fromfile(f, np.float64).reshape(B, 4 + n)
        npdt = np.float32 if dt == "f32" else np.float64
        kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dt == "f32" else {}
        th2 = 0.0625 if (fam == "warp" and n == 6) else 0.01
        A, y, xs, x0 = O.synth_generate(B, m, n, npdt)
        xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), robust=(kind, th2))
        iters, stop, cost, fails, xg = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4:]
        if kind in (1, 2, 3, 6):
            assert np.array_equal(iters, ro["num_iters"]) and np.array_equal(stop, ro["stop_reason"]), f.name
            assert np.array_equal(fails, ro["num_failures"]) and np.array_equal(cost, ro["final_cost"]), f.name
            assert np.array_equal(xg, xo.astype(np.float64)), f.name
        else:
            same = iters == ro["num_iters"]
            assert same.mean() > 0.9, (f.name, same.mean())
            err = np.abs(xg - xo)[same].max(axis=1) / np.maximum(np.abs(xo)[same].max(axis=1), 1e-300)
            assert err.max() < 1e-9, (f.name, err.max())

    # run-time n (the *Large drivers): the manual functor through the SolverType seam == the oracle bit for bit in both
    # precisions and for every n; numeric differentiation (diff/num_diff.h, central differences, the NORM as cost) ==
    # the oracle's numdiff variant bit for bit
    ldumps = sorted(tmp_path.glob("large_*.bin"))
    assert len(ldumps) == 14, ldumps
    for f in ldumps:
        tag, dt, n, m, B = re.match(r"large_(\w+)_(f\d+)_n(\d+)_m(\d+)_B(\d+)\.bin", f.name).groups()
        n, m, B = int(n), int(m), int(B)
        rec = np.fromfile(f, np.float64).reshape(B, 4 + n)
        npdt = np.float32 if dt == "f32" else np.float64
        kw = dict(min_rerr_dec=1e-5, min_step_norm2=1e-9) if dt == "f32" else {}
        A, y, xs, x0 = O.synth_generate(B, m, n, npdt)
        # numdiff: kCentral with h = FloatEpsilon; numdiffFwd: kForward; numdiffFast: kFastCentral with an explicit h
        nd = {"manual": None, "numdiff": (2, 0.0), "numdiffFwd": (1, 0.0),
              "numdiffFast": (3, float(npdt(1e-6 if dt == "f64" else 5e-4)))}[tag]
        xo, ro, _ = O.synth_lm_run(A, y, x0, O.default_options(**kw), numdiff=nd)
        iters, stop, cost, fails, xg = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4:]
        assert np.array_equal(iters, ro["num_iters"]) and np.array_equal(stop, ro["stop_reason"]), f.name
        assert np.array_equal(fails, ro["num_failures"]) and np.array_equal(cost, ro["final_cost"]), f.name
        assert np.array_equal(xg, xo.astype(np.float64)), f.name
