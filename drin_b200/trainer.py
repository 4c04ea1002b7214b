"""Fused train / ranking steps around ``drin_b200.Model`` (the body of the reference's
``MELModel._forward_step`` + ``configure_optimizers``, upstream train.py:30-44,55-56) and their
data-parallel form over ``torch.distributed``.

Data parallelism: mentions are sharded across ranks, parameters replicated.
  * The reference loss couples every mention with every score of the batch (common/utils.py:41-42), so
    the local ``[B_loc, C]`` scores and labels are all-gathered BEFORE the loss; each rank then gets the
    gradient of the GLOBAL loss for its own rows.  The result equals the reference run at the global
    batch size, not at the local one.
  * Parameter gradients are summed with ONE all-reduce over the flat gradient buffer (26.8 MB live,
    31.5 MB total) -- no averaging: the 1/B_glob^2 normalisation is already in dL/dscores.
  * Ranking needs no communication beyond a final gather of the scores.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

from .loss import triplet_loss_sharded
from .model import Model
from .optim import FusedAdam


def _dist_on(group) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def gather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equal-sized row shards into the global [B_glob, ...] tensor (rank order)."""
    if not _dist_on(group):
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class Trainer:
    def __init__(self, model: Model, lr: float = 1e-3, margin: float = 0.25, group=None):
        self.model, self.margin, self.group = model, float(margin), group
        self.opt = FusedAdam(model, lr=lr)
        self.last_scores: Optional[torch.Tensor] = None

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if _dist_on(self.group) else 0

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if _dist_on(self.group) else 1

    def forward_backward(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        """batch: the loader's 15-tuple (14 inputs + uint8 labels), local shard.  Leaves summed gradients in
        model.flat_grads and returns the global loss (device scalar)."""
        m = self.model
        inputs, y = tuple(batch[:-1]), batch[-1]
        params = m._param_views()
        scores, ctx = m._engine.forward(inputs, params, training=True, num_candidates_model=m.num_candidates_model)
        self.last_scores = scores
        b_loc = scores.shape[0]
        scores_all = gather_rows(scores, self.group)
        labels_all = gather_rows(y.to(torch.uint8), self.group)
        loss, dscores = triplet_loss_sharded(scores_all, labels_all, self.margin, self.rank * b_loc, b_loc)
        m._engine.backward(ctx, inputs, params, dscores, m._grad_views())
        if _dist_on(self.group):
            dist.all_reduce(m.flat_grads, op=dist.ReduceOp.SUM, group=self.group)   # one 31.5 MB bucket
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss.reshape(())

    def step(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        loss = self.forward_backward(batch)
        self.opt.step()
        return loss

    @torch.no_grad()
    def rank_scores(self, batch: Sequence[torch.Tensor], gather: bool = False) -> torch.Tensor:
        """Ranking inference (upstream test_step under no_grad): scores [B_loc, C]; optionally gathered."""
        m = self.model
        inputs = tuple(batch[:14])
        scores, _ = m._engine.forward(inputs, m._param_views(), training=False,
                                      num_candidates_model=m.num_candidates_model)
        return gather_rows(scores, self.group) if gather else scores
