"""Fused train / ranking steps around ``drin_b200.Model`` (the body of the reference's
``MELModel._forward_step`` + ``configure_optimizers``, upstream train.py:30-44,55-56) and their
data-parallel form over ``torch.distributed``.

Data parallelism: mentions are sharded across ranks, parameters replicated.
  * The reference loss couples every mention with every score of the batch (common/utils.py:41-42), so
    the local ``[B_loc, C]`` scores and labels are all-gathered BEFORE the loss; each rank then gets the
    gradient of the GLOBAL loss for its own rows.  The result equals the reference run at the global
    batch size, not at the local one.
  * Parameter gradients are summed with ONE all-reduce over the live part of the flat gradient buffer (26.8 MB;
    the dead parameters sit behind it) -- no averaging: the 1/B_glob^2 normalisation is already in dL/dscores.  The per-rank loss
    shares travel in the same bucket, and the labels travel with the scores: two collectives per step in total.
  * Ranking needs no communication beyond a final gather of the scores.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

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
    if dist.get_backend(group) == "gloo":         # CPU / single-device test setups: gloo gathers host tensors only
        host = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather(list(host.chunk(world)), local.detach().cpu().contiguous(), group=group)
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def gather_scores_and_labels(scores: torch.Tensor, labels: torch.Tensor, group=None):
    """Global ``[B_glob, C]`` scores and ``[B_glob, C-1]`` uint8 labels from the local shards with ONE all-gather: the
    one-hot labels ride along with the scores as 0/1 floats."""
    if not _dist_on(group):
        return scores, labels.to(torch.uint8)
    Cn = scores.shape[1]
    packed = gather_rows(torch.cat([scores, labels.to(scores.dtype)], dim=1), group)
    return packed[:, :Cn].contiguous(), packed[:, Cn:].to(torch.uint8)


def reduce_grads_and_loss(bucket: torch.Tensor, n: int, loss_share: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the flat gradients (``bucket[:n]``) and the per-rank loss shares over the ranks with ONE all-reduce: the
    loss share travels in ``bucket[n]``.  No averaging: 1/B_glob^2 is already in dL/dscores."""
    if not _dist_on(group):
        return loss_share
    bucket[n:n + 1].copy_(loss_share.reshape(1))
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket[n:n + 1].clone()


class Trainer:
    def __init__(self, model: Model, lr: float = 1e-3, margin: float = 0.25, group=None, overlap_allreduce: bool = False):
        self.model, self.margin, self.group = model, float(margin), group
        self.opt = FusedAdam(model, lr=lr)
        self.last_scores: Optional[torch.Tensor] = None
        # data parallel, optional: all-reduce the GCN-layer gradients on a side stream while the four input-projection
        # weight-gradient GEMMs (the tail of backward) are still running.  Measured on 8 B200: no gain (4.27-4.34 ms
        # vs 4.23-4.26 ms per step) -- the NCCL kernels take SMs from the persistent GEMMs -- so it is off by default.
        self.overlap_allreduce = overlap_allreduce or os.environ.get("DRIN_OVERLAP_ALLREDUCE", "0") == "1"
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self._layers_done: Optional[torch.cuda.Event] = None

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if _dist_on(self.group) else 0

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if _dist_on(self.group) else 1

    def forward_backward(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        """batch: the loader's 15-tuple (14 inputs + uint8 labels) or a ``store.IndexedBatch``, local shard.
        Leaves summed gradients in model.flat_grads and returns the global loss (device scalar)."""
        m = self.model
        if hasattr(batch, "mention_index"):          # resident feature store: rows are gathered by the front end
            inputs, y = batch, batch.labels
        else:
            inputs, y = tuple(batch[:-1]), batch[-1]
        params = m._param_views()
        scores, ctx = m._engine.forward(inputs, params, training=True, num_candidates_model=m.num_candidates_model)
        self.last_scores = scores
        b_loc = scores.shape[0]
        scores_all, labels_all = gather_scores_and_labels(scores, y, self.group)
        loss, dscores = triplet_loss_sharded(scores_all, labels_all, self.margin, self.rank * b_loc, b_loc)
        if _dist_on(self.group) and self.overlap_allreduce:
            return self._backward_overlapped(ctx, inputs, params, dscores, loss)
        m._engine.backward(ctx, inputs, params, dscores, m._grad_views())
        ctx.release()
        # 26.8 MB of live gradients + this rank's share of the loss in the bucket's tail
        loss = reduce_grads_and_loss(m.flat_grads_bucket, m.n_live, loss, self.group)
        m._flat_grads_valid = True          # FusedAdam reads the flat buffer directly
        return loss.reshape(())

    def _backward_overlapped(self, ctx, inputs, params, dscores, loss_share) -> torch.Tensor:
        """Backward with the all-reduce split in two: [GCN-layer gradients + loss share] goes out on a side stream as
        soon as the library signals that they are final, the vertex-encoder gradients follow when backward is done."""
        m = self.model
        dev = m.flat_params.device
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=dev)
            self._layers_done = torch.cuda.Event()
            self._layers_done.record()                    # creates the underlying cudaEvent_t
        bucket, n, off = m.flat_grads_bucket, m.n_live, m.layer_grad_offset()
        bucket[n:n + 1].copy_(loss_share.reshape(1))      # before backward: covered by the layers-done event
        m._engine.backward(ctx, inputs, params, dscores, m._grad_views(), layers_done=self._layers_done)
        ctx.release()
        self._comm_stream.wait_event(self._layers_done)
        with torch.cuda.stream(self._comm_stream):
            work = dist.all_reduce(bucket[off:], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        dist.all_reduce(bucket[:off], op=dist.ReduceOp.SUM, group=self.group)
        work.wait()                                       # the current stream waits for the side-stream collective
        m._flat_grads_valid = True
        return bucket[n:n + 1].clone().reshape(())

    def step(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        loss = self.forward_backward(batch)
        self.opt.step()
        return loss

    def rank_scores(self, batch: Sequence[torch.Tensor], gather: bool = False) -> torch.Tensor:
        """Ranking inference (upstream test_step under no_grad): scores [B_loc, C]; optionally gathered."""
        return rank_scores(self.model, batch, self.group, gather)


@torch.no_grad()
def rank_scores(model: Model, batch, group=None, gather: bool = False) -> torch.Tensor:
    """Forward without backward state over the loader's batch (first 14 tensors) or a ``store.IndexedBatch``."""
    inputs = batch if hasattr(batch, "mention_index") else tuple(batch[:14])
    scores, _ = model._engine.forward(inputs, model._param_views(), training=False,
                                      num_candidates_model=model.num_candidates_model)
    return gather_rows(scores, group) if gather else scores


class Evaluator:
    """Validation / test loop body of the reference (``MELModel._forward_step`` with type 1 / 2, upstream train.py:30-43)
    without its per-step host synchronisations: the reference formats ``float(loss)`` and every ``metric.compute()`` into
    a log line each step and, for ``output_test_result``, copies the scores to the host and writes them out per step
    (train.py:35-43).  Here the loss sum and the top-k hit counters stay on the device (``drin_topk_hits``), the scores
    of every step are kept in HBM, and ``compute()`` / ``write_results()`` do ONE device->host read at the end.

    Per-step loss semantics are the reference's: TripletLoss over the batch of that step; ``compute()["loss"]`` is the
    mean over steps, i.e. what Lightning's epoch aggregation of the returned losses reports."""

    def __init__(self, model: Model, margin: float = 0.25, top_k: Sequence[int] = (1, 3, 5), keep_scores: bool = False):
        from .loss import TopkAccuracy
        self.model, self.margin, self.keep_scores = model, float(margin), bool(keep_scores)
        self.metric = TopkAccuracy(list(top_k), device=model.flat_params.device)
        self.reset()

    def reset(self) -> None:
        self.metric.reset()
        self._loss_sum = torch.zeros(1, dtype=torch.float32, device=self.model.flat_params.device)
        self._steps = 0
        self._scores: List[torch.Tensor] = []
        self._labels: List[torch.Tensor] = []

    @torch.no_grad()
    def step(self, batch) -> torch.Tensor:
        """batch: the loader's 15-tuple or a ``store.IndexedBatch``.  Returns the step's loss (device scalar, no sync)."""
        y = batch.labels if hasattr(batch, "mention_index") else batch[-1]
        scores = rank_scores(self.model, batch)
        loss, _ = triplet_loss_sharded(scores, y, self.margin)
        self._loss_sum += loss
        self.metric.update(scores, y)
        self._steps += 1
        if self.keep_scores:
            self._scores.append(scores)          # forward() returns a fresh tensor every call
            self._labels.append(y.to(torch.uint8))
        return loss.reshape(())

    def compute(self, acc_correction: float = 0.0) -> dict:
        """One host read.  ``acc_correction``: the reference divides top-k accuracy by ``1 - acc_correction[type]``
        (train.py:38, args.py:115,123) to discount mentions whose gold entity is not among the candidates."""
        acc = (self.metric.correct.double() / max(self.metric.total, 1) / (1.0 - acc_correction)).tolist()
        return {"loss": float(self._loss_sum) / max(self._steps, 1),
                "topk": {k: a for k, a in zip(self.metric.top_k, acc)}, "mentions": self.metric.total}

    def write_results(self, file, batch_size: Optional[int] = None) -> int:
        """The reference's ``test-result.txt`` (train.py:40-43): per mention ``"{index}:\t{scores as a list}\n{labels}\n"``
        with ``index = i + batch_idx * batch_size``.  ``file``: path or text file object.  Returns the mention count."""
        if not self.keep_scores:
            raise RuntimeError("Evaluator(keep_scores=True) is needed to write the result file")
        own = isinstance(file, (str, os.PathLike))
        fh = open(file, "w") if own else file
        n = 0
        try:
            scores = [s.cpu() for s in self._scores]         # device -> host once, after the loop
            labels = [y.cpu() for y in self._labels]
            for batch_idx, (s, y) in enumerate(zip(scores, labels)):
                bs = batch_size if batch_size is not None else (scores[0].shape[0])
                for i, sample in enumerate(s.tolist()):
                    fh.write(f"{i + batch_idx * bs}:\t{sample}\n{y[i]}\n")
                    n += 1
            fh.flush()
        finally:
            if own:
                fh.close()
        return n


class GraphedStoreStep:
    """One CUDA graph for the whole train step (front end .. Adam) over a resident ``FeatureStore``.

    The reference's default batch is 64 mentions (common/args.py:118,126): at that size the ~60 kernels of a step are
    a few microseconds each and launch / Python overhead dominates.  The step is captured once for a fixed batch size;
    ``step(idx)`` copies the ``[B]`` mention indices into the graph's static index buffer and replays it.  Single
    process only (the data-parallel step keeps the eager path: its collectives run between the captured pieces).
    """

    def __init__(self, trainer: "Trainer", store, batch_size: int, warmup: int = 2):
        if trainer.world != 1:
            raise RuntimeError("GraphedStoreStep supports a single process; use Trainer.step for data parallelism")
        self.trainer, self.store, self.B = trainer, store, int(batch_size)
        dev = store.device
        self.idx = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self._host_idx = torch.zeros(self.B, dtype=torch.int64).pin_memory()
        self._idx_copied: Optional[torch.cuda.Event] = None
        # optimizer / parameter state is advanced by warm-up and capture runs: snapshot and restore it
        opt, m = trainer.opt, trainer.model
        saved = (m.flat_params.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt._step.clone())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # allocates the workspace / scratch outside the capture
                trainer.step(store.select(self.idx))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="relaxed"):
            self.loss = trainer.step(store.select(self.idx))
            self.scores = trainer.last_scores
        # the graph has raw pointers into the engine's workspace(s) and the loss scratch baked in: keep them alive for
        # as long as the graph exists, whatever other batch sizes the engine or the loss see in between
        from .loss import scratch_tensors
        self._keep = list(m._engine.pool.tensors()) + scratch_tensors() + [m.flat_grads, store]
        with torch.no_grad():
            m.flat_params.copy_(saved[0])
            opt.exp_avg.copy_(saved[1])
            opt.exp_avg_sq.copy_(saved[2])
            opt._step.copy_(saved[3])

    def step(self, idx) -> torch.Tensor:
        """idx: [B] mention indices (host tensor / list, or device tensor).  Returns the loss (static device scalar)."""
        idx = torch.as_tensor(idx, dtype=torch.int64)
        if idx.numel() != self.B:
            raise ValueError(f"graph was captured for {self.B} mentions, got {idx.numel()}")
        if idx.is_cuda:
            self.idx.copy_(idx, non_blocking=True)
        else:
            if int(idx.min()) < 0 or int(idx.max()) >= len(self.store):
                raise IndexError("mention index out of range")
            if self._idx_copied is not None:
                self._idx_copied.synchronize()          # the previous step's async copy has read the pinned buffer
            self._host_idx.copy_(idx)
            self.idx.copy_(self._host_idx, non_blocking=True)
            if self._idx_copied is None:
                self._idx_copied = torch.cuda.Event()
            self._idx_copied.record()
        self.graph.replay()
        return self.loss


class HostFeeder:
    """Host -> device path for the loader's CPU batches (what Lightning's ``batch.to(device)`` does before
    reference train.py:46), double buffered on a side stream and copying only the bytes the path reads:

    * ``mention_text_feature`` is ``[B, 128, 768]`` (393 KB/mention) but the model only ever reads the rows
      ``start:end`` of each mention (ghmfc.py:55-60): the span rows are gathered on the host into a pinned
      ``[B, Ls, 768]`` buffer (``Ls`` = longest span of the batch) and start/end are rebased to ``0:len``;
    * ``mention_text_mask`` is never read by the DRIN configuration (ghmfc.py:25-26): a tiny placeholder is sent.

    Everything else is copied as is.  ``last_bytes`` is the number of bytes actually moved for the last batch.
    """

    def __init__(self, device, slots: int = 2, compact_spans: bool = True):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict() for _ in range(slots)]
        self.next_slot = 0
        self.compact_spans = compact_spans
        self.last_bytes = 0

    def _pinned(self, slot, name, shape, dtype):
        t = slot.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype).pin_memory()
            slot[name] = t
        return t

    def _dev(self, slot, name, like):
        t = slot.get(name)
        if t is None or tuple(t.shape) != tuple(like.shape) or t.dtype != like.dtype:
            t = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            slot[name] = t
        return t

    def submit(self, host_batch: Sequence[torch.Tensor]) -> int:
        """Start moving one CPU batch (14 inputs [+ labels]); returns the slot to pass to ``get``."""
        sid = self.next_slot
        self.next_slot = (self.next_slot + 1) % len(self.slots)
        slot = self.slots[sid]
        if slot.get("ready") is not None:
            slot["ready"].synchronize()        # the slot's previous async copy has read its pinned staging buffers
        srcs = list(host_batch)
        mtf, start, end = srcs[0], srcs[2], srcs[3]
        Lm = mtf.shape[1]
        ok = self.compact_spans and bool(((start >= 0) & (end >= start) & (end <= Lm)).all())
        if ok:
            lens = end - start
            Ls = max(int(lens.max()), 1)
            idx = (start.unsqueeze(1) + torch.arange(Ls)).clamp_(max=Lm - 1)               # [B, Ls]
            compact = self._pinned(slot, "h_mtf", (mtf.shape[0], Ls, mtf.shape[2]), mtf.dtype)
            torch.gather(mtf, 1, idx.unsqueeze(-1).expand(-1, -1, mtf.shape[2]), out=compact)
            h_start = self._pinned(slot, "h_start", start.shape, start.dtype).zero_()
            h_end = self._pinned(slot, "h_end", end.shape, end.dtype).copy_(lens)
            h_mask = self._pinned(slot, "h_mask", (mtf.shape[0], 1), torch.int64).zero_()
            srcs[0], srcs[1], srcs[2], srcs[3] = compact, h_mask, h_start, h_end
        else:
            srcs = [t if t.is_pinned() else t.pin_memory() for t in srcs]
        done = slot.get("used")
        with torch.cuda.stream(self.stream):
            if done is not None:
                self.stream.wait_event(done)           # the previous consumer of this slot has finished
            out, nbytes = [], 0
            for i, h in enumerate(srcs):
                d = self._dev(slot, f"d{i}", h)
                d.copy_(h, non_blocking=True)
                nbytes += h.numel() * h.element_size()
                out.append(d)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        slot["batch"], slot["ready"] = out, ev
        self.last_bytes = nbytes
        return sid

    def get(self, sid: int) -> List[torch.Tensor]:
        """Device batch of a submitted slot; the current stream waits for its copy."""
        slot = self.slots[sid]
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return slot["batch"]

    def release(self, sid: int) -> None:
        """Call after the work that reads the slot has been enqueued on the current stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.slots[sid]["used"] = ev
