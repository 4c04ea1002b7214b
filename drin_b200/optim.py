"""Adam over the model's flat parameter buffer as one kernel launch (reference train.py:55-56:
``torch.optim.Adam(self.parameters(), lr=learning_rate)`` with PyTorch defaults)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam:
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        flat = model.flat_params
        # the step count lives on the device so that a captured CUDA graph of the train step can be replayed
        self._step = torch.zeros(1, dtype=torch.int32, device=flat.device)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.skip = model.skip_mask()      # parameters whose grad is None upstream are never stepped

    def _collect_autograd_grads(self) -> None:
        """``Model.forward`` + ``loss.backward()`` (the reference's Lightning flow, train.py:46-56) leaves the gradients
        in ``p.grad``, not in the flat buffer the kernel reads (only ``Trainer`` writes that one directly): copy them in.
        Parameters whose grad is None must be exactly the reference's dead ones (they are masked out of the update)."""
        m = self.model
        views, dead, found = m._grad_views(), set(m._dead), False
        for name, p in m.named_parameters():
            if p.grad is None:
                if name not in dead:
                    raise RuntimeError(f"FusedAdam.step(): parameter {name} has no gradient -- run backward first "
                                       "(frozen parameters are not supported)")
                continue
            found = True
            if p.grad.data_ptr() != views[name].data_ptr():
                views[name].copy_(p.grad)
        if not found:
            raise RuntimeError("FusedAdam.step(): no gradients -- neither Trainer.forward_backward nor loss.backward() ran")

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.model._flat_grads_valid = False
        for p in self.model.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self) -> None:
        m = self.model
        if not getattr(m, "_flat_grads_valid", False):
            self._collect_autograd_grads()
        flat, grad = m.flat_params, m.flat_grads
        if self.exp_avg.device != flat.device:
            raise RuntimeError("model moved after the optimizer was built")
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(flat.device):
            self._step.add_(1)
            _lib.check(_lib.load().drin_adam_step_dev(
                p(flat), p(grad), p(self.exp_avg), p(self.exp_avg_sq), p(self.skip), C.c_int64(flat.numel()),
                p(self._step), C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]),
                C.c_float(self.eps), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "drin_adam_step_dev")
        m._flat_grads_valid = False        # consumed: the next step needs a new backward (Trainer or autograd)

    @property
    def step_count(self) -> int:
        return int(self._step)

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, lr=self.lr,
                    betas=self.betas, eps=self.eps)

    def load_state_dict(self, sd):
        self._step.fill_(int(sd["step"]))
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
