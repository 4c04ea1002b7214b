"""Host-side driver of the CUDA hot path: turns the reference's 14-tensor batch and state_dict-named
parameters into the plain-pointer structs of the C ABI (include/drin_b200.h) and owns the workspace.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libdrin_b200.so.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

FP32, BF16 = 0, 1

INPUT_NAMES = (
    "mention_text_feature", "mention_text_mask", "mention_start_pos", "mention_end_pos",
    "mention_image_feature", "mention_object_feature", "mention_object_score", "entity_text_feature",
    "entity_text_mask", "entity_image_feature", "entity_object_feature", "entity_object_score",
    "miet_similarity", "mtei_similarity",
)

# state_dict keys (reference drin/model.py:21-24,111-119,159-162) -> fields of drin_params
VERTEX_KEYS = (
    ("vertex_encoder.mention_text_encoder.final_layer.linear.weight", "w_mt"),
    ("vertex_encoder.mention_text_encoder.final_layer.linear.bias", "b_mt"),
    ("vertex_encoder.entity_text_encoder.final_layer.weight", "w_et"),
    ("vertex_encoder.entity_text_encoder.final_layer.bias", "b_et"),
    ("vertex_encoder.mention_image_linear.weight", "w_mi"),
    ("vertex_encoder.mention_image_linear.bias", "b_mi"),
    ("vertex_encoder.entity_image_linear.weight", "w_ei"),
    ("vertex_encoder.entity_image_linear.bias", "b_ei"),
)
LAYER_FIELDS = (
    ("w_h.weight", "w_h"), ("w_h.bias", "b_h"), ("w_u.weight", "w_u"), ("w_u.bias", "b_u"),
    ("w_v.weight", "w_v"), ("w_v.bias", "b_v"), ("layer_norm.weight", "ln_w"), ("layer_norm.bias", "ln_b"),
)
# gcn_edge_feature == "vector": w_m is a Linear too, created right after w_h (drin/model.py:111-116)
LAYER_FIELDS_VECTOR = LAYER_FIELDS[:2] + (("w_m.weight", "w_m"), ("w_m.bias", "b_m")) + LAYER_FIELDS[2:]


def layer_fields(vector_edges: bool = False):
    return LAYER_FIELDS_VECTOR if vector_edges else LAYER_FIELDS


def param_keys(num_layers: int, vector_edges: bool = False) -> List[str]:
    keys = [k for k, _ in VERTEX_KEYS]
    for l in range(num_layers):
        keys += [f"gcn_layers.{l}.{s}" for s, _ in layer_fields(vector_edges)]
    return keys


def param_shapes(num_layers: int, D: int = 768, R: int = 2048, vector_edges: bool = False) -> Dict[str, Tuple[int, ...]]:
    shapes = {
        VERTEX_KEYS[0][0]: (D, D), VERTEX_KEYS[1][0]: (D,), VERTEX_KEYS[2][0]: (D, D), VERTEX_KEYS[3][0]: (D,),
        VERTEX_KEYS[4][0]: (D, R), VERTEX_KEYS[5][0]: (D,), VERTEX_KEYS[6][0]: (D, R), VERTEX_KEYS[7][0]: (D,),
    }
    H = D // 2 if vector_edges else D          # vector edges: w_u / w_v map D -> D/2 (model.py:113-116)
    for l in range(num_layers):
        for s, _ in layer_fields(vector_edges):
            if s in ("w_u.weight", "w_v.weight"):
                shape = (H, D)
            elif s in ("w_u.bias", "w_v.bias"):
                shape = (H,)
            elif s in ("w_h.weight", "w_m.weight"):
                shape = (D, D)
            else:
                shape = (D,)
            shapes[f"gcn_layers.{l}.{s}"] = shape
    return shapes


def dead_param_keys(num_layers: int, static_edges: bool = False, vector_edges: bool = False) -> List[str]:
    """Parameters that never receive a gradient in the reference (grad is None): the edge update of the
    last GCN layer is dead code (drin/model.py:131-134); with gcn_edge_type="static" no layer
    runs an edge update (model.py:135-136).  With vector edges the edge update also owns w_m."""
    layers = range(num_layers) if static_edges else (num_layers - 1,)
    names = ("w_u.weight", "w_u.bias", "w_v.weight", "w_v.bias")
    if vector_edges:
        names = ("w_m.weight", "w_m.bias") + names
    return [f"gcn_layers.{l}.{s}" for l in layers for s in names]


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class Problem:
    """Shapes of one call, derived from the batch exactly like the reference switches on tensor rank
    (drin/model.py:43-44,73-75,78-83)."""
    B: int
    C: int
    Lm: int
    Le: int          # 0 = WikiDiverse layout
    P: int
    Om: int
    Oe: int
    D: int
    R: int
    precision: int

    def key(self):
        return (self.B, self.C, self.Lm, self.Le, self.P, self.Om, self.Oe, self.D, self.R, self.precision)


def inspect_batch(batch: Sequence[torch.Tensor], num_candidates_model: Optional[int] = None) -> Problem:
    if len(batch) != 14:
        raise ValueError(f"DRIN forward expects a 14-tensor batch (drin/model.py:164-180), got {len(batch)}")
    (mtf, _mm, start, end, mif, mof, mos, etf, emask, eif, eof, eos, miet, mtei) = batch
    for name, t in zip(INPUT_NAMES, batch):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback in drin_b200)")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    feat_dtype = mtf.dtype
    if feat_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"feature dtype {feat_dtype} not supported (float32 or bfloat16)")
    for name, t in (("mention_image_feature", mif), ("mention_object_feature", mof), ("entity_text_feature", etf),
                    ("entity_image_feature", eif), ("entity_object_feature", eof)):
        if t.dtype != feat_dtype:
            raise RuntimeError(f"{name} has dtype {t.dtype}, expected {feat_dtype} like mention_text_feature")
    for name, t in (("mention_object_score", mos), ("entity_object_score", eos), ("miet_similarity", miet),
                    ("mtei_similarity", mtei)):
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    for name, t in (("mention_start_pos", start), ("mention_end_pos", end)):
        if t.dtype != torch.int64:
            raise RuntimeError(f"{name} must be int64, got {t.dtype}")
    if mtf.dim() != 3:
        raise RuntimeError(f"mention_text_feature must be [B, Lm, D], got {tuple(mtf.shape)}")
    B, Lm, D = mtf.shape
    if etf.dim() == 3:
        Le = 0
    elif etf.dim() == 4:
        Le = etf.shape[2]
        if emask.dtype != torch.int64 or tuple(emask.shape) != (B, etf.shape[1], Le):
            raise RuntimeError(f"entity_text_mask must be int64 [B, C, Le], got {emask.dtype} {tuple(emask.shape)}")
    else:
        raise RuntimeError(f"entity_text_feature must have rank 3 or 4, got {tuple(etf.shape)}")
    Cc = etf.shape[1]
    if num_candidates_model is not None and Cc != num_candidates_model:
        # the reference hard-codes num_candidates_model in its expand() calls (model.py:72,80-85,146,150,208)
        raise RuntimeError(f"batch has {Cc} candidate slots but num_candidates_model = {num_candidates_model}")
    if mif.dim() != 3 or mif.shape[0] != B:
        raise RuntimeError(f"mention_image_feature must be [B, P, R], got {tuple(mif.shape)}")
    P, R = mif.shape[1], mif.shape[2]
    if mof.dim() == 4:
        if mof.shape[2] != 1:
            raise RuntimeError("mention_object_feature: only a singleton crop dim is supported ([B, Om, 1, R])")
        Om = mof.shape[1]
    elif mof.dim() == 3:
        Om = mof.shape[1]
    else:
        raise RuntimeError(f"mention_object_feature has bad shape {tuple(mof.shape)}")
    if eof.dim() == 5:
        if eof.shape[3] != 1:
            raise RuntimeError("entity_object_feature: only a singleton crop dim is supported ([B, C, Oe, 1, R])")
        Oe = eof.shape[2]
    elif eof.dim() == 4:
        Oe = eof.shape[2]
    else:
        raise RuntimeError(f"entity_object_feature has bad shape {tuple(eof.shape)}")
    if eif.dim() == 4 and eif.shape[2] != 1:
        raise RuntimeError("entity_image_feature: only [B, C, R] or [B, C, 1, R] is supported")

    def want(name, t, shape):
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")

    want("mention_start_pos", start, (B,))
    want("mention_end_pos", end, (B,))
    want("mention_object_score", mos, (B, Om))
    want("entity_object_score", eos, (B, Cc, Oe))
    want("miet_similarity", miet, (B, Cc))
    want("mtei_similarity", mtei, (B, Cc))
    if eif.numel() != B * Cc * R or eof.numel() != B * Cc * Oe * R or mof.numel() != B * Om * R:
        raise RuntimeError("image / object feature shapes are inconsistent with [B, C, R]")
    if etf.shape[0] != B or etf.shape[-1] != D:
        raise RuntimeError(f"entity_text_feature has bad shape {tuple(etf.shape)}")
    return Problem(B, Cc, Lm, Le, P, Om, Oe, D, R, BF16 if feat_dtype == torch.bfloat16 else FP32)


class _Lease:
    """Ownership of one pooled workspace by one forward call.  ``live``: saved activations are still needed;
    ``done``: a backward has run (the workspace may be taken over by a later forward, which bumps its generation so
    that a second backward through this lease fails loudly instead of reading overwritten activations)."""
    __slots__ = ("entry", "gen", "state", "__weakref__")

    def __init__(self, entry, gen):
        self.entry, self.gen, self.state = entry, gen, "live"

    def valid(self) -> bool:
        return self.state != "released" and self.entry["gen"] == self.gen

    def release(self) -> None:
        if self.valid():
            self.entry["lease"] = None
        self.state = "released"


class WorkspacePool:
    """Workspaces of one Engine.  Every training forward owns its workspace until its backward has run (or its
    autograd graph is freed), so ``loss(model(b1)) + loss(model(b2))`` or a no_grad validation forward between
    forward and backward work like they do with the reference module; the steady state of a train loop is still
    ONE workspace (a finished lease is reclaimed by the next forward)."""

    def __init__(self):
        self.entries: List[dict] = []

    def _lease_of(self, e) -> Optional[_Lease]:
        ref = e["lease"]
        lease = ref() if ref is not None else None
        if lease is None or lease.state == "released" or lease.gen != e["gen"]:
            e["lease"] = None
            return None
        return lease

    def acquire(self, need: int, device) -> _Lease:
        mine = [e for e in self.entries if e["ws"].device == device]
        free = [e for e in mine if self._lease_of(e) is None]
        fit = sorted((e for e in free if e["ws"].numel() >= need), key=lambda e: e["ws"].numel())
        if fit:
            e = fit[0]
        else:
            free_ids = {id(e) for e in free}           # identity, never ==: the entries are dicts holding tensors
            done = sorted((e for e in mine if id(e) not in free_ids and self._lease_of(e).state == "done"
                           and e["ws"].numel() >= need), key=lambda e: e["ws"].numel())
            if done:
                e = done[0]
            else:                                      # too small and unowned: let the allocator have them back
                self.entries = [x for x in self.entries if id(x) not in free_ids]
                e = dict(ws=torch.empty(need, dtype=torch.uint8, device=device), gen=0, lease=None)
                self.entries.append(e)
        e["gen"] += 1
        lease = _Lease(e, e["gen"])
        e["lease"] = weakref.ref(lease)
        return lease

    def tensors(self) -> List[torch.Tensor]:
        return [e["ws"] for e in self.entries]


class ForwardCtx:
    """State of one forward call: problem, config struct, workspace and its lease.  Unpacks like the former
    ``(pb, cfg, ws)`` tuple."""
    __slots__ = ("pb", "cfg", "ws", "lease")

    def __init__(self, pb, cfg, ws, lease):
        self.pb, self.cfg, self.ws, self.lease = pb, cfg, ws, lease

    def __iter__(self):
        return iter((self.pb, self.cfg, self.ws))

    def release(self) -> None:
        """The saved activations are no longer needed (called by Trainer after backward)."""
        self.lease.release()


class Engine:
    """One instance per module: owns the workspace pool and marshals calls into the C ABI."""

    def __init__(self, num_layers: int, edge_enabled: Sequence[float] = (1, 1, 1, 1), static_edges: bool = False,
                 vector_edges: bool = False):
        self.lib = _lib.load()
        self.num_layers = int(num_layers)
        self.static_edges = bool(static_edges)
        self.vector_edges = bool(vector_edges)
        self.edge_enabled = tuple(float(x) for x in edge_enabled)
        if len(self.edge_enabled) != 4:
            raise ValueError("gcn_edge_enabled must have 4 entries")
        self.pool = WorkspacePool()

    # ---- struct marshalling -----------------------------------------------------------------
    def config(self, pb: Problem, training: bool, indexed: bool = False) -> _lib.DrinConfig:
        cfg = _lib.DrinConfig()
        cfg.batch, cfg.candidates, cfg.mention_tokens, cfg.entity_tokens = pb.B, pb.C, pb.Lm, pb.Le
        cfg.regions, cfg.mention_objects, cfg.entity_objects = pb.P, pb.Om, pb.Oe
        cfg.embed_dim, cfg.resnet_dim, cfg.gcn_layers = pb.D, pb.R, self.num_layers
        cfg.precision, cfg.training = pb.precision, int(training)
        for i in range(4):
            cfg.edge_enabled[i] = self.edge_enabled[i]
        cfg.static_edges = int(self.static_edges)
        cfg.indexed = int(indexed)
        cfg.vector_edges = int(self.vector_edges)
        return cfg

    @staticmethod
    def inputs(batch) -> _lib.DrinInputs:
        """batch: the 14 tensors of one batch, or a store.IndexedBatch (resident tables + row indices)."""
        s = _lib.DrinInputs()
        indexed = hasattr(batch, "mention_index")
        for name, t in zip(INPUT_NAMES, batch.tables if indexed else batch):
            setattr(s, name, t.data_ptr())
        if indexed:
            s.mention_index = batch.mention_index.data_ptr()
            s.entity_index = 0 if batch.entity_index is None else batch.entity_index.data_ptr()
        return s

    def params(self, tensors: Dict[str, torch.Tensor], D: int, R: int) -> _lib.DrinParams:
        """tensors: state_dict-keyed fp32 CUDA tensors (parameters, or gradient buffers of the same shapes)."""
        s = _lib.DrinParams()
        shapes = param_shapes(self.num_layers, D, R, self.vector_edges)
        for key, fld in VERTEX_KEYS:
            setattr(s, fld, self._checked(tensors, key, shapes[key]))
        for l in range(self.num_layers):
            for suffix, fld in layer_fields(self.vector_edges):
                key = f"gcn_layers.{l}.{suffix}"
                setattr(s.layer[l], fld, self._checked(tensors, key, shapes[key]))
        return s

    @staticmethod
    def _checked(tensors, key, shape) -> int:
        t = tensors.get(key)
        if t is None:
            return 0
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"parameter {key}: need contiguous CUDA float32 {shape}, got {t.dtype} {tuple(t.shape)}")
        if t.data_ptr() % 16:
            raise RuntimeError(f"parameter {key} is not 16-byte aligned")
        return t.data_ptr()

    # ---- workspace --------------------------------------------------------------------------
    def workspace_bytes(self, cfg) -> int:
        n = C.c_size_t(0)
        _lib.check(self.lib.drin_workspace_bytes(C.byref(cfg), C.byref(n)), "drin_workspace_bytes")
        return n.value

    def workspace(self, cfg, device) -> _Lease:
        return self.pool.acquire(self.workspace_bytes(cfg), device)

    # ---- calls ------------------------------------------------------------------------------
    def forward(self, batch, params: Dict[str, torch.Tensor], training: bool,
                num_candidates_model: Optional[int] = None):
        indexed = hasattr(batch, "mention_index")
        if indexed:
            pb = batch.problem()
            if num_candidates_model is not None and pb.C != num_candidates_model:
                raise RuntimeError(f"store has {pb.C} candidate slots but num_candidates_model = {num_candidates_model}")
            if not batch.mention_index.is_cuda:
                raise RuntimeError("the feature store must live on a CUDA device (no CPU fallback in drin_b200)")
            dev = batch.mention_index.device
        else:
            pb = inspect_batch(batch, num_candidates_model)
            dev = batch[0].device
        cfg = self.config(pb, training, indexed)
        with torch.cuda.device(dev):
            lease = self.workspace(cfg, dev)
            ws = lease.entry["ws"]
            scores = torch.empty(pb.B, pb.C, dtype=torch.float32, device=dev)
            ins = self.inputs(batch)
            ps = self.params(params, pb.D, pb.R)
            try:
                _lib.check(self.lib.drin_forward(C.byref(cfg), C.byref(ins), C.byref(ps), _ptr(ws),
                                                 C.c_size_t(ws.numel()), _ptr(scores), _stream()), "drin_forward")
            except Exception:
                lease.release()
                raise
        if not training:
            lease.release()        # nothing is saved: the next call on this stream may reuse the memory
        return scores, ForwardCtx(pb, cfg, ws, lease)

    def backward(self, ctx, batch, params, dscores: torch.Tensor, grads: Dict[str, torch.Tensor],
                 layers_done: Optional[torch.cuda.Event] = None) -> None:
        """layers_done: optional event recorded on the current stream once every GCN-layer and bias gradient is final
        (only the four input-projection weight gradients still follow)."""
        pb, cfg, ws = ctx
        if not cfg.training:
            raise RuntimeError("backward needs a forward with training=True (activations were not saved)")
        if not ctx.lease.valid():
            raise RuntimeError("the activations of this forward are gone: its workspace was released or taken over by a "
                               "later forward after a first backward (a second backward through the same graph is only "
                               "possible when no other forward ran in between)")
        dscores = dscores.contiguous()
        with torch.cuda.device(dscores.device):
            ins = self.inputs(batch)
            ps = self.params(params, pb.D, pb.R)
            gs = self.params(grads, pb.D, pb.R)
            ev = C.c_void_p(0 if layers_done is None else layers_done.cuda_event)
            _lib.check(self.lib.drin_backward_ex(C.byref(cfg), C.byref(ins), C.byref(ps), _ptr(ws),
                                                 C.c_size_t(ws.numel()), _ptr(dscores), C.byref(gs), ev, _stream()),
                       "drin_backward")
        ctx.lease.state = "done"

    def debug_buffer(self, ctx, name: str, layer: int = 0) -> torch.Tensor:
        """Copy of a named fp32 intermediate of the last forward (tests only)."""
        pb, cfg, ws = ctx
        ptr, rows, cols = C.c_void_p(0), C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.drin_debug_buffer(C.byref(cfg), _ptr(ws), name.encode(), C.c_int32(layer), C.byref(ptr),
                                              C.byref(rows), C.byref(cols)), "drin_debug_buffer")
        off = ptr.value - ws.data_ptr()
        n = rows.value * cols.value
        return ws[off:off + 4 * n].view(torch.float32).view(rows.value, cols.value).clone()
