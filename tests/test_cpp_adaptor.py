"""The C++ host side (include/tinyopt_b200.hpp over the C-ABI): tests/cpp/test_adaptor.cpp restates
the reference's tests/sqrt2.cpp, tests/basic.cpp:41-54, tests/solvers.cpp:25-45 and tests/circle.cpp
for a batch, and checks that the host-driven SolverType seam equals the device-resident loop bit for
bit.  Compiled with plain g++ (no CUDA headers) against the in-tree library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
EXE = os.path.join(CPP, "build", "test_adaptor")


def build_exe():
    subprocess.run(["make", "-C", CPP], check=True, capture_output=True)
    assert os.path.exists(EXE)


def test_cpp_adaptor_builds_and_fails_loudly_without_gpu():
    import torch
    build_exe()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by test_cpp_adaptor_on_gpu")
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 3, p.stdout + p.stderr
    assert "no CUDA device" in p.stdout and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_cpp_adaptor_on_gpu():
    build_exe()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all C++ adaptor checks passed" in p.stdout
    assert "sqrt2: x[0]=1.414213562373095" in p.stdout and "iters=5 stop=1" in p.stdout
