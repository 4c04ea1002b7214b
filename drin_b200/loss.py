"""Device versions of the reference's loss and metric (upstream common/utils.py:26-73).

``TripletLoss`` keeps the reference call signature ``loss(y_true, y_pred)`` and its cross-batch
semantics; forward and backward are one C-ABI call (csrc/loss.cu).  ``triplet_loss_sharded`` is the
data-parallel form: every rank passes the gathered global scores and gets its share of the loss plus the
gradient of the GLOBAL loss for its own rows.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_scratch = {}          # (B, C, device) -> scratch tensor; a few shapes alternate in practice (train / eval / last batch)
_SCRATCH_KEEP = 8


def _loss_scratch(B: int, Cn: int, device) -> torch.Tensor:
    key = (B, Cn, str(device))
    t = _scratch.pop(key, None)
    if t is None:
        n = C.c_size_t(0)
        _lib.check(_lib.load().drin_loss_scratch_bytes(C.c_int32(B), C.c_int32(Cn), C.byref(n)), "drin_loss_scratch_bytes")
        t = torch.empty(n.value, dtype=torch.uint8, device=device)
        while len(_scratch) >= _SCRATCH_KEEP:         # least recently used first (dicts keep insertion order)
            _scratch.pop(next(iter(_scratch)))
    _scratch[key] = t                                 # most recently used last
    return t


def scratch_tensors():
    """The cached scratch buffers (a CUDA graph that captured the loss keeps references to them)."""
    return list(_scratch.values())


def triplet_loss_sharded(scores_all: torch.Tensor, labels_all: torch.Tensor, margin: float, row_offset: int = 0,
                         rows_local: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores_all [B_glob, C] fp32, labels_all [B_glob, C-1] uint8 -> (loss share [1], dscores [rows_local, C])."""
    if not scores_all.is_cuda:
        raise RuntimeError("drin_b200 loss needs CUDA tensors (no CPU fallback)")
    B, Cn = scores_all.shape
    if labels_all.shape != (B, Cn - 1):
        raise RuntimeError(f"labels must be [B, C-1] = {(B, Cn - 1)}, got {tuple(labels_all.shape)}")
    rows_local = B - row_offset if rows_local is None else rows_local
    scores_all = scores_all.contiguous().float()
    labels_all = labels_all.contiguous().to(torch.uint8)
    with torch.cuda.device(scores_all.device):
        loss = torch.empty(1, dtype=torch.float32, device=scores_all.device)
        dscores = torch.empty(rows_local, Cn, dtype=torch.float32, device=scores_all.device)
        scratch = _loss_scratch(B, Cn, scores_all.device)
        _lib.check(_lib.load().drin_triplet_loss(_ptr(scores_all), _ptr(labels_all), C.c_int32(B), C.c_int32(Cn),
                                                 C.c_int32(row_offset), C.c_int32(rows_local), C.c_float(margin),
                                                 _ptr(loss), _ptr(dscores), _ptr(scratch), _stream()),
                   "drin_triplet_loss")
    return loss, dscores


class _TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, margin):
        loss, dscores = triplet_loss_sharded(y_pred, y_true, margin)
        ctx.save_for_backward(dscores)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dscores,) = ctx.saved_tensors
        return dscores * g, None, None


class TripletLoss:
    """common/utils.py:26-43.  y_true: one-hot [B, C-1]; y_pred: scores [B, C] (gold slot last) or [B, C-1]."""

    def __init__(self, margin):
        self.margin = float(margin)

    def __call__(self, y_true, y_pred):
        if y_pred.shape[1] == y_true.shape[1]:          # already sliced: give the kernel a dummy gold slot
            y_pred = torch.cat([y_pred, torch.zeros_like(y_pred[:, :1])], dim=1)
        return _TripletFn.apply(y_pred, y_true, self.margin)


class TopkAccuracy:
    """common/utils.py:46-73 with device-side counters (no host sync in update())."""

    def __init__(self, top_k: Sequence[int] | int, device="cuda") -> None:
        self.top_k = [int(top_k)] if isinstance(top_k, int) else [int(k) for k in top_k]
        self.device = torch.device(device)
        self.reset()

    def reset(self):
        self.correct = torch.zeros(len(self.top_k), dtype=torch.int64, device=self.device)
        self.total = 0

    def update(self, y_pred: torch.Tensor, y_true: torch.Tensor):
        B, Cn = y_pred.shape
        if Cn == y_true.shape[1]:
            y_pred = torch.cat([y_pred, torch.zeros_like(y_pred[:, :1])], dim=1)
            Cn += 1
        y_pred = y_pred.detach().contiguous().float()
        y_true = y_true.contiguous().to(torch.uint8)
        ks = (C.c_int32 * len(self.top_k))(*self.top_k)
        with torch.cuda.device(y_pred.device):
            _lib.check(_lib.load().drin_topk_hits(_ptr(y_pred), _ptr(y_true), C.c_int32(B), C.c_int32(Cn), ks,
                                                  C.c_int32(len(self.top_k)), _ptr(self.correct), _stream()),
                       "drin_topk_hits")
        self.total += B

    def compute(self):
        acc = self.correct.double() / max(self.total, 1)
        return acc[0] if len(self.top_k) == 1 else acc
