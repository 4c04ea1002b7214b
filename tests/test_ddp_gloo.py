"""World-size-2 check of the data-parallel protocol on CPU (gloo): shard the mentions, all-gather the
scores BEFORE the loss (it couples the whole batch), take the gradient of the GLOBAL loss for the local
rows, sum parameter gradients with one flat all-reduce -- the result must equal the full-batch reference
step.  Host logic under test: drin_b200.trainer's collective helpers (one packed all-gather, one bucket all-reduce) and
the no-averaging convention; the per-rank
arithmetic is done by the CPU oracle (the CUDA kernels are covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drin_b200.synthetic import make_batch, spread_weights
from drin_b200.trainer import gather_rows, gather_scores_and_labels, reduce_grads_and_loss
from oracle import drin_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = O.DrinConfig(num_candidates_model=11, triplet_margin=0.05)
    B = 8
    batch = make_batch("wikidiverse", B, 9, 10)
    sd = spread_weights(O.init_state(cfg, 0))
    bl = B // world
    shard = [t[rank * bl:(rank + 1) * bl] for t in batch]
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    scores = O.forward(leaves, shard[:-1], cfg)
    scores_all, labels_all = gather_scores_and_labels(scores.detach(), shard[-1], None)     # ONE all-gather
    assert labels_all.dtype == torch.uint8 and torch.equal(labels_all, batch[-1])
    share, dscores = O.triplet_sharded(scores_all, labels_all, cfg.triplet_margin, rank * bl, bl)
    scores.backward(dscores)
    live = [k for k in sd if leaves[k].grad is not None]
    flat = torch.cat([leaves[k].grad.flatten() for k in live])
    bucket = torch.cat([flat, torch.zeros(4)])            # gradients + loss share: ONE all-reduce, summed NOT averaged
    loss = reduce_grads_and_loss(bucket, flat.numel(), share, None)
    flat = bucket[:flat.numel()]
    if rank == 0:
        s_ref, l_ref, g_ref = O.train_step_grads(sd, batch[:-1], batch[-1], cfg)
        ref_flat = torch.cat([g_ref[k].flatten() for k in live])
        out["scores_equal"] = bool(torch.allclose(scores_all, s_ref, rtol=1e-5, atol=1e-6))
        out["loss_err"] = abs(float(loss) - float(l_ref)) / abs(float(l_ref))
        out["grad_err"] = float((flat - ref_flat).abs().max() / ref_flat.abs().max())
        out["dead"] = sorted(k for k in sd if g_ref[k] is None) == sorted(k for k in sd if k not in live)
    dist.destroy_process_group()


def test_two_rank_data_parallel_step_equals_full_batch_step():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    assert res["scores_equal"] and res["dead"]
    assert res["loss_err"] < 1e-5
    assert res["grad_err"] < 1e-4


def test_sharded_loss_closed_form_matches_autograd():
    g = torch.Generator().manual_seed(1)
    s = (torch.rand(12, 11, generator=g) * 2 - 1).requires_grad_(True)
    y = torch.eye(10, dtype=torch.uint8)[torch.randint(0, 10, (12,), generator=g)]
    y[3] = 0                                              # gold not among the candidates
    O.triplet_loss(y, s, 0.25).backward()
    shares, parts = zip(*[O.triplet_sharded(s.detach(), y, 0.25, r * 4, 4) for r in range(3)])
    assert torch.allclose(torch.cat(parts), s.grad, atol=1e-7)
    assert abs(float(sum(shares)) - float(O.triplet_loss(y, s.detach(), 0.25))) < 1e-6


def test_gather_rows_without_process_group_is_identity():
    t = torch.arange(6).view(3, 2)
    assert gather_rows(t) is t
