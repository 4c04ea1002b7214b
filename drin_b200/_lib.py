"""ctypes binding of libdrin_b200.so (the C ABI declared in include/drin_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

DRIN_MAX_LAYERS = 8
c_void_p, c_int, c_int32, c_int64, c_float, c_size_t = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float, C.c_size_t


class DrinConfig(C.Structure):
    _fields_ = [
        ("batch", c_int32), ("candidates", c_int32), ("mention_tokens", c_int32), ("entity_tokens", c_int32),
        ("regions", c_int32), ("mention_objects", c_int32), ("entity_objects", c_int32), ("embed_dim", c_int32),
        ("resnet_dim", c_int32), ("gcn_layers", c_int32), ("precision", c_int32), ("training", c_int32),
        ("edge_enabled", c_float * 4), ("static_edges", c_int32), ("indexed", c_int32),
        ("vector_edges", c_int32),
    ]


class DrinInputs(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "mention_text_feature", "mention_text_mask", "mention_start_pos", "mention_end_pos",
        "mention_image_feature", "mention_object_feature", "mention_object_score", "entity_text_feature",
        "entity_text_mask", "entity_image_feature", "entity_object_feature", "entity_object_score",
        "miet_similarity", "mtei_similarity", "mention_index", "entity_index")]


class DrinLayerParams(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("w_h", "b_h", "w_u", "b_u", "w_v", "b_v", "ln_w", "ln_b", "w_m", "b_m")]


class DrinParams(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("w_mt", "b_mt", "w_et", "b_et", "w_mi", "b_mi", "w_ei", "b_ei")] + [
        ("layer", DrinLayerParams * DRIN_MAX_LAYERS)]


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building in-tree first if nvcc is present and the .so is stale or absent)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        override = os.environ.get("DRIN_B200_LIB")        # A/B testing of two builds on the same box
        if override:
            path = override
        elif not _build.is_fresh():
            if _build.have_nvcc():
                _build.build()          # a compile / link error after a source edit must surface, never a stale library
            elif not os.path.exists(path):
                raise RuntimeError("libdrin_b200.so is missing and there is no nvcc to build it")
            else:
                import warnings
                warnings.warn("libdrin_b200.so does not match the sources and nvcc is absent: using the shipped library")
        lib = C.CDLL(path)
        lib.drin_last_error.restype = C.c_char_p
        # the ctypes mirrors above must match the structs the library was compiled with
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        lib.drin_struct_sizes(C.byref(a), C.byref(b), C.byref(c))
        want = (C.sizeof(DrinConfig), C.sizeof(DrinInputs), C.sizeof(DrinParams))
        if (a.value, b.value, c.value) != want:
            raise RuntimeError(f"libdrin_b200.so ABI mismatch: struct sizes {(a.value, b.value, c.value)} != {want} "
                               "(stale library? rebuild with python -m drin_b200.build --force)")
        _lib = lib
        return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().drin_last_error().decode(errors="replace")
        raise RuntimeError(f"drin_b200 {what} failed (status {status}): {msg}")
