"""bench.py contract on a machine without a GPU: the reference arm (the CPU oracle port, every host core even
when torchrun pins OMP_NUM_THREADS=1) prints ONE JSON line with the keys the driver reads, and our arm
fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_json_contract():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun does to every rank
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("LM iterations/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))   # not OMP_NUM_THREADS
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # headline = C4 (the config BASELINE.json quotes at 1/2/4/8 GPUs), C2 / C3 / C5 as sub-records, and the
    # `config` object is the one our arm prints for the same workload (the driver compares them)
    import bench
    assert d["config"] == bench.config_of("C4") and d["gpu_launches"] == 0
    assert [c["name"] for c in d["configs"]] == ["C2", "C3", "C5"]
    for c in d["configs"]:
        assert c["config"] == bench.config_of(c["name"]) and c["value"] > 0 and c["cpu_baseline"]["kind"] == "port"


def test_reference_arm_single_config():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C2", "--steps", "2",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][0])
    assert d["config"]["workload"].startswith("C2") and "configs" not in d and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode != 0
    assert "{" not in p.stdout   # no bench line without the CUDA path
