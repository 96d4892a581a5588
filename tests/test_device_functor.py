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
def test_device_functor_on_gpu():
    build_exe()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all device functor checks passed" in p.stdout
    assert "closed forms and scale == Jet derivative: ok" in p.stdout
    assert "sqrt2 (Jet functor): x[0]=1.414213562373095" in p.stdout and "iters=5 stop=1" in p.stdout
