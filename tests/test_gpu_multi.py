"""NCCL data-parallel path on real GPUs (skipped on a single-GPU box): N ranks, each a shard of one global
batch, reproduce the single-GPU full-batch step.  The CPU/gloo version of the same protocol is
tests/test_ddp_gloo.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("edge_feature", ["scaler", "vector"])
def test_two_gpu_step_equals_single_gpu_full_batch_step(edge_feature):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533" if edge_feature == "scaler" else "29534",
           os.path.join(ROOT, "scripts", "gpu_dp_check.py"), edge_feature]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("DPCHECK ")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(line[-1][8:])
    assert res["loss_err"] < 1e-6 and res["grad_err"] < 1e-5 and res["rank_scores_err"] < 1e-5


def test_two_ranks_on_one_gpu_equal_the_full_batch_step():
    """The same check on a single-GPU box: two processes share cuda:0 and talk over gloo (NCCL refuses two ranks on one
    device).  Everything but the transport is the product path: sharded forward, packed score / label all-gather, global
    TripletLoss for the local rows, backward, ONE all-reduce over the live-gradient bucket with the loss share in its
    tail, Adam, gathered ranking scores."""
    env = dict(os.environ, DRIN_DP_SAME_DEVICE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "scripts", "gpu_dp_check.py"), "scaler"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in out.stdout.splitlines() if l.startswith("DPCHECK ")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(line[-1][8:])
    assert res["world"] == 2
    assert res["loss_err"] < 1e-6 and res["grad_err"] < 1e-5 and res["rank_scores_err"] < 1e-5
    assert res["param_err_after_adam"] < 5e-5          # Adam turns 1e-7 gradient noise into up to lr * noise / |g| on tiny entries


def test_fit_data_parallel_equals_single_process_fit():
    """drin_b200.fit with two ranks (one shared GPU, gloo) and a GLOBAL batch of 16 = the single-process fit of the same
    schedule: same permutations, the loss of the global batch (scores all-gathered), gradients summed."""
    env = dict(os.environ, DRIN_DP_SAME_DEVICE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29536", os.path.join(ROOT, "scripts", "gpu_fit_dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in out.stdout.splitlines() if l.startswith("FITCHECK ")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(line[-1][9:])
    assert res["records"] == ["training", "validating", "testing"] * 2
    assert res["loss_err"] < 1e-5 and res["param_err"] < 5e-4        # 10 Adam steps amplify 1e-7 gradient noise
    assert res["topk_err"] < 1e-9                                    # training accuracies are summed over the ranks' shards
