"""A 100-line stand-in for the ``lightning`` package -- TEST INFRASTRUCTURE, never imported by the product.

The reference's entry point ``train.py`` drives everything through Lightning (``import lightning as pl``,
train.py:7,20,59,112-122,134-144), which is absent from this image.  This stub implements exactly the surface train.py
uses, with Lightning's automatic-optimisation semantics, so the UNMODIFIED ``train.py`` can be executed in tests:

  pl.seed_everything(seed)                              -> python / numpy / torch seeds
  pl.LightningModule                                    -> nn.Module with ``configure_optimizers`` / ``*_step`` hooks
  pl.Callback                                           -> no-op hooks
  pl.Trainer(max_epochs=, accelerator=, callbacks=, ..) -> ``fit(model, train_loader, val_loader)``: per batch
        move to the device, ``loss = training_step(batch, i)``, ``zero_grad``, ``backward``, ``optimizer.step()``;
        the validation loop runs under no_grad inside the epoch, before ``on_train_epoch_end``; ``test(model, loader)``
        runs ``test_step`` under no_grad.  ``current_epoch`` counts from 0 like Lightning's.
"""
from __future__ import annotations

import random
import sys
import types

import numpy as np
import torch


def seed_everything(seed: int = 0) -> int:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


class LightningModule(torch.nn.Module):
    trainer = None

    def log(self, *a, **k):
        pass


class Callback:
    def on_fit_start(self, trainer, pl_module): ...
    def on_train_epoch_start(self, trainer, pl_module): ...
    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx): ...
    def on_train_epoch_end(self, trainer, pl_module): ...
    def on_validation_epoch_start(self, trainer, pl_module): ...
    def on_validation_epoch_end(self, trainer, pl_module): ...
    def on_test_epoch_start(self, trainer, pl_module): ...
    def on_test_epoch_end(self, trainer, pl_module): ...


def _to(batch, device):
    return [t.to(device) if torch.is_tensor(t) else t for t in batch]


class Trainer:
    def __init__(self, max_epochs=1, accelerator=None, callbacks=None, devices=1, **_ignored):
        self.max_epochs = max_epochs
        self.device = torch.device("cuda" if accelerator == "gpu" else "cpu")
        cbs = callbacks if callbacks is not None else []
        self.callbacks = list(cbs) if isinstance(cbs, (list, tuple)) else [cbs]
        self.current_epoch = 0

    def _call(self, hook, *args):
        for cb in self.callbacks:
            getattr(cb, hook)(self, *args)

    def fit(self, model, train_dataloaders, val_dataloaders=None):
        model.trainer = self
        model.to(self.device)
        opt = model.configure_optimizers()
        self._call("on_fit_start", model)
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            model.train()
            self._call("on_train_epoch_start", model)
            for i, batch in enumerate(train_dataloaders):
                batch = _to(batch, self.device)
                loss = model.training_step(batch, i)
                opt.zero_grad()
                loss.backward()
                opt.step()
                self._call("on_train_batch_end", model, loss, batch, i)
            if val_dataloaders is not None:
                model.eval()
                self._call("on_validation_epoch_start", model)
                with torch.no_grad():
                    for i, batch in enumerate(val_dataloaders):
                        model.validation_step(_to(batch, self.device), i)
                self._call("on_validation_epoch_end", model)
                model.train()
            self._call("on_train_epoch_end", model)

    def test(self, model, dataloaders):
        model.trainer = self
        model.to(self.device)
        model.eval()
        self._call("on_test_epoch_start", model)
        with torch.no_grad():
            for i, batch in enumerate(dataloaders):
                model.test_step(_to(batch, self.device), i)
        self._call("on_test_epoch_end", model)


def install() -> None:
    """Register the stub as ``lightning`` (and ``lightning.pytorch``) unless the real package is importable."""
    if "lightning" in sys.modules:
        return
    try:
        import lightning  # noqa: F401
        return
    except ImportError:
        pass
    mod = types.ModuleType("lightning")
    for name in ("seed_everything", "LightningModule", "Callback", "Trainer"):
        setattr(mod, name, globals()[name])
    mod.__stub__ = True
    sys.modules["lightning"] = mod
    sys.modules["lightning.pytorch"] = mod
    mod.pytorch = mod
