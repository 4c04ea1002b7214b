"""Shared helpers for the parity tests (oracle = checker only)."""
import glob
import os

import torch

from drin_b200.synthetic import make_batch, spread_weights
from oracle import drin_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_model_cases():
    """Scalar-edge cases first (tests index into them), then the gcn_edge_feature="vector" cases."""
    paths = [p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")) if "triplet_loss" not in p]
    return sorted(paths, key=lambda p: ("vector" in os.path.basename(p), p))


def load_case(path):
    """Rebuild (cfg, batch, state, fixture) of a golden case from its seed; checksums guard drift."""
    fx = torch.load(path, weights_only=False)
    case = fx["case"]
    ov = case.get("overrides", {})
    cfg = O.DrinConfig(num_candidates_model=case["cands"] + 1,
                       num_gcn_layers=ov.get("num_gcn_layers", 2),
                       gcn_edge_enabled=tuple(ov.get("gcn_edge_enabled", (1, 1, 1, 1))),
                       gcn_edge_type=ov.get("gcn_edge_type", "dynamic"),
                       gcn_edge_feature=ov.get("gcn_edge_feature", "scaler"),
                       triplet_margin=ov.get("triplet_margin", 0.25))
    batch = make_batch(case["dataset"], case["B"], case["seed"], case["cands"], **case.get("batch_kw", {}))
    sd = O.init_state(cfg, seed=0)
    if case["weights"] == "spread":
        sd = spread_weights(sd)
    assert _md5(batch) == fx["input_checksum"], "synthetic generator drifted from the golden fixture"
    assert _md5(sd.values()) == fx["weight_checksum"], "weight init drifted from the golden fixture"
    return cfg, batch, sd, fx


def _md5(tensors):
    import hashlib
    return [hashlib.md5(t.contiguous().numpy().tobytes()).hexdigest() for t in tensors]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (tensor-wise relative error used for the 1e-4 fp32 bar)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
