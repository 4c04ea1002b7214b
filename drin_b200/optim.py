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
        self.step_count = 0
        flat = model.flat_params
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.skip = model.dead_mask()      # parameters whose grad is None upstream are never stepped

    def step(self) -> None:
        m = self.model
        flat, grad = m.flat_params, m.flat_grads
        if self.exp_avg.device != flat.device:
            raise RuntimeError("model moved after the optimizer was built")
        self.step_count += 1
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(flat.device):
            _lib.check(_lib.load().drin_adam_step(
                p(flat), p(grad), p(self.exp_avg), p(self.exp_avg_sq), p(self.skip), C.c_int64(flat.numel()),
                C.c_int32(self.step_count), C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]),
                C.c_float(self.eps), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "drin_adam_step")

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, lr=self.lr,
                    betas=self.betas, eps=self.eps)

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
