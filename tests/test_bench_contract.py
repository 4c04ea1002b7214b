"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, on
the same `config` as our arm, timed on the host with the unmodified reference when it is staged."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_line():
    lines = _run("--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-batch", "16")
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "mentions/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the same `config` dict as our arm prints for the default workload
    sys.path.insert(0, ROOT)
    import bench
    ns = bench.parse_args.__globals__["argparse"].Namespace(dataset="wikidiverse", precision="fp32", edge_feature="scaler")
    assert d["config"] == bench.workload_config(ns, 1, 4096)
    from oracle import ref_import
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_import.available() else "port")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env=env) == []


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-legs",
                          "--no-e2e", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0          # no CPU fallback: the product arm fails loudly
