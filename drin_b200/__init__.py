"""drin_b200: B200-native (sm_100a) training / ranking hot path of starreeze/drin.

The arithmetic lives in ``libdrin_b200.so`` (C ABI: include/drin_b200.h); this package is the thin
Python host side that mirrors the reference's module interface.
"""
from .model import Model  # noqa: F401
from .loss import TripletLoss, TopkAccuracy  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .trainer import Evaluator, GraphedStoreStep, HostFeeder, Trainer  # noqa: F401
from .store import FeatureStore, IndexedBatch  # noqa: F401
from .fit import evaluate, fit  # noqa: F401

__all__ = ["Model", "TripletLoss", "TopkAccuracy", "FusedAdam", "Trainer", "Evaluator", "GraphedStoreStep", "HostFeeder", "FeatureStore", "IndexedBatch",
           "fit", "evaluate", "install_as_reference_module"]


def install_as_reference_module() -> None:
    """Make ``from drin import model as model_module`` (reference train.py:13-14) resolve to this package's
    Model without touching the reference tree: aliases ``drin.model`` in ``sys.modules``."""
    import sys
    import types

    from . import model as _m

    pkg = sys.modules.get("drin")
    if pkg is None:
        pkg = types.ModuleType("drin")
        pkg.__path__ = []  # type: ignore[attr-defined]
        sys.modules["drin"] = pkg
    sys.modules["drin.model"] = _m
    pkg.model = _m  # type: ignore[attr-defined]
