"""Import the UNMODIFIED reference (starreeze/drin) -- TEST / BASELINE INFRASTRUCTURE.

``/root/reference`` exists only in the build container, never on the GPU box; ``oracle/_ref/`` is the staged copy of
the four files of the path (``oracle/make_ref.py``; git-ignored, travels with gpurun).  Used by
``oracle/make_golden.py`` (golden-vector generation), by tests that skip when neither tree is present and by
``bench.py``'s CPU reference legs.  Recipe from SURVEY.md section 8(c):
  * ``common.args`` is patched BEFORE any model module is imported (consumers star-import it, so the
    values are copied at import time);
  * ``torchmetrics`` (absent from this image) is stubbed with a 6-line ``Metric``;
  * to switch dataset shape the reference modules are purged from ``sys.modules`` and re-imported.
"""
from __future__ import annotations

import os
import sys
import types

import torch

def _find_root() -> str:
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for root in (os.environ.get("DRIN_REFERENCE_ROOT"), "/root/reference", staged):
        if root and os.path.isfile(os.path.join(root, "drin", "model.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "drin", "model.py"))


def _stub_torchmetrics() -> None:
    if "torchmetrics" in sys.modules:
        return
    tm = types.ModuleType("torchmetrics")

    class Metric(torch.nn.Module):
        def add_state(self, name, default, dist_reduce_fx=None):
            setattr(self, name, default)

    tm.Metric = Metric
    sys.modules["torchmetrics"] = tm


def load(dataset: str, num_candidates: int, entity_tokens: int = 64, **overrides):
    """Returns (model_module, utils_module, args_module) of the reference configured for `dataset`."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True            # the reference tree is read-only
    for name in [m for m in sys.modules if m.split(".")[0] in ("common", "baselines", "drin")]:
        del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _stub_torchmetrics()
    import common.args as a

    a.use_device = "cpu"
    a.dataset_name = dataset
    a.num_candidates_data = num_candidates
    a.max_entity_attr_token_len = entity_tokens
    a.num_candidates_model = num_candidates + 1
    for k, v in overrides.items():
        setattr(a, k, v)
    import common.utils as u
    import drin.model as m

    return m, u, a
