"""Drop-in replacement for the reference's ``drin.model.Model`` (upstream drin/model.py:156-209).

Same no-argument constructor, same ``forward(batch)`` over the 14-tensor batch of drin/data.py, same
``state_dict`` keys and (given the same seed) bit-identical initial weights -- but forward and backward
run as hand-written sm_100a kernels behind the C ABI of ``include/drin_b200.h``.  There is no PyTorch or
CPU fallback: a missing library or a non-CUDA batch raises.

Hyper-parameters are read, like upstream, from the flat ``common.args`` module when it is importable
(args.py:25-36,101); keyword overrides exist for tests and benchmarks that have no ``common`` package.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import engine as E

_DEFAULTS = dict(
    gcn_embed_dim=768, num_gcn_layers=2, bert_embed_dim=768, resnet_embed_dim=2048,
    num_candidates_model=None,          # None: taken from the batch
    gcn_edge_type="dynamic", gcn_edge_feature="scaler", gcn_edge_enabled=(1, 1, 1, 1),
    gcn_vertex_activation="gelu", gcn_edge_activation="sigmoid",
    mention_final_layer_name="linear", mention_final_representation="avg extract",
    entity_final_layer_name="linear", entity_final_pooling="avg", online_bert=False,
    mention_final_output_dim=768, entity_final_output_dim=768,
)


def _read_args(overrides: dict) -> dict:
    cfg = dict(_DEFAULTS)
    try:  # the reference's flat config module (values are read at construction, like its star-imports)
        import common.args as a  # type: ignore

        for k in cfg:
            if hasattr(a, k):
                cfg[k] = getattr(a, k)
    except ImportError:
        pass
    cfg.update(overrides)
    return cfg


def _check_supported(cfg: dict) -> None:
    want = dict(gcn_vertex_activation="gelu",
                gcn_edge_activation="sigmoid", mention_final_layer_name="linear",
                mention_final_representation="avg extract", entity_final_layer_name="linear",
                entity_final_pooling="avg", online_bert=False)
    bad = {k: cfg[k] for k, v in want.items() if cfg[k] != v}
    if cfg["gcn_edge_type"] not in ("dynamic", "static"):
        bad["gcn_edge_type"] = cfg["gcn_edge_type"]
    if cfg["gcn_edge_feature"] not in ("scaler", "vector"):
        bad["gcn_edge_feature"] = cfg["gcn_edge_feature"]
    if bad:
        raise NotImplementedError(
            f"drin_b200 implements DRIN with scalar or vector edges (dynamic or static, any layer count / edge "
            f"mask); unsupported settings: {bad} (see SURVEY.md section 8f)")
    if cfg["gcn_embed_dim"] != 768 or cfg["bert_embed_dim"] != 768:
        raise NotImplementedError("kernels are built for gcn_embed_dim = bert_embed_dim = 768")
    if cfg["mention_final_output_dim"] != cfg["gcn_embed_dim"] or cfg["entity_final_output_dim"] != cfg["gcn_embed_dim"]:
        raise NotImplementedError("final output dims must equal gcn_embed_dim (args.py:38-39)")


# ---- module tree with the reference's attribute names (only the parameter holders) ----------------
class _AvgLinear(nn.Module):       # baselines/ghmfc.py:63-69
    def __init__(self, i, o):
        super().__init__()
        self.linear = nn.Linear(i, o)


class _MentionEncoder(nn.Module):  # baselines/ghmfc.py:152-165 (linear mode)
    def __init__(self, i, o):
        super().__init__()
        self.final_layer = _AvgLinear(i, o)


class _EntityEncoder(nn.Module):   # baselines/ghmfc.py:202-212 (linear mode)
    def __init__(self, i, o):
        super().__init__()
        self.final_layer = nn.Linear(i, o)


class _VertexEncoder(nn.Module):   # drin/model.py:19-24
    def __init__(self, D, Db, R):
        super().__init__()
        self.mention_text_encoder = _MentionEncoder(Db, D)
        self.entity_text_encoder = _EntityEncoder(Db, D)
        self.mention_image_linear = nn.Linear(R, D)
        self.entity_image_linear = nn.Linear(R, D)


class _GCNLayer(nn.Module):        # drin/model.py:109-119
    def __init__(self, D, vector_edges=False):
        super().__init__()
        self.w_h = nn.Linear(D, D)
        if vector_edges:               # model.py:112: w_m is nn.Identity() (no parameters) for scalar edges
            self.w_m = nn.Linear(D, D)
        H = D // 2 if vector_edges else D
        self.w_u = nn.Linear(D, H)
        self.w_v = nn.Linear(D, H)
        self.layer_norm = nn.LayerNorm(D)


class _DrinFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward and backward are single C-ABI calls."""

    @staticmethod
    def forward(ctx, model: "Model", batch, need_grad: bool, *params):
        names = model._param_names
        pdict = dict(zip(names, params))
        scores, ectx = model._engine.forward(batch, pdict, training=need_grad,
                                             num_candidates_model=model.num_candidates_model)
        ctx.model, ctx.batch, ctx.ectx, ctx.pdict = model, batch, ectx, pdict
        return scores

    @staticmethod
    def backward(ctx, dscores):
        model = ctx.model
        # a fresh buffer per backward: autograd may keep (steal) the returned tensors as .grad
        grads = model._grad_views(torch.empty_like(model.flat_params))
        model._engine.backward(ctx.ectx, ctx.batch, ctx.pdict, dscores, grads)
        dead = set(model._dead)
        out = []
        for name, p in zip(model._param_names, ctx.pdict.values()):
            # dead parameters get None exactly like the reference (last layer's w_u / w_v: grad is None)
            out.append(None if (name in dead or not p.requires_grad) else grads[name])
        return (None, None, None, *out)


class Model(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = _read_args(overrides)
        _check_supported(cfg)
        self.cfg = cfg
        D, Db, R = cfg["gcn_embed_dim"], cfg["bert_embed_dim"], cfg["resnet_embed_dim"]
        self.num_gcn_layers = int(cfg["num_gcn_layers"])
        self.num_candidates_model = cfg["num_candidates_model"]
        # same creation order as upstream -> same weights under the same seed
        self.vertex_encoder = _VertexEncoder(D, Db, R)
        self.vector_edges = cfg["gcn_edge_feature"] == "vector"
        self.gcn_layers = nn.ModuleList([_GCNLayer(D, self.vector_edges) for _ in range(self.num_gcn_layers)])
        self._param_names: List[str] = [n for n, _ in self.named_parameters()]
        assert self._param_names == E.param_keys(self.num_gcn_layers, self.vector_edges), \
            "state_dict keys drifted from the reference"
        self.static_edges = cfg["gcn_edge_type"] == "static"
        self._dead = E.dead_param_keys(self.num_gcn_layers, self.static_edges, self.vector_edges)
        self._engine_obj: Optional[E.Engine] = None
        self._flat: Optional[torch.Tensor] = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._flatten()

    # ---- flat parameter / gradient storage (one buffer: one Adam launch, one all-reduce bucket) ----
    TAIL = 4     # floats between the live and the dead parameters: slot 0 carries the rank's loss share in the all-reduce

    def _flatten(self) -> None:
        """Layout of the flat buffer: [live parameters in state_dict order | TAIL floats | dead parameters].  The
        parameters that never get a gradient upstream (the last layer's w_u / w_v, 4.7 MB) sit at the end so
        that the data-parallel all-reduce is ONE contiguous bucket without them: gradients[: n_live + TAIL]."""
        params = dict(self.named_parameters())
        dev = next(iter(params.values())).device
        dead = set(self._dead)
        order = [n for n in self._param_names if n not in dead] + [None] + [n for n in self._param_names if n in dead]
        total = sum(p.numel() for p in params.values()) + self.TAIL
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self._offsets: Dict[str, tuple] = {}
        with torch.no_grad():
            for name in order:
                if name is None:
                    self._n_live = off
                    off += self.TAIL
                    continue
                p = params[name]
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = flat[off:off + n].view_as(p)
                self._offsets[name] = (off, n, tuple(p.shape))
                off += n
        self._flat = flat
        self._flat_grad = None
        self._flat_grads_valid = False      # set by Trainer once it has written the flat gradient buffer

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._flatten()          # .to()/.cuda() re-create the parameter storages: restore the flat layout
        return out

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def n_live(self) -> int:
        """Number of leading elements of the flat buffers that belong to parameters with a gradient."""
        return self._n_live

    @property
    def flat_grads(self) -> torch.Tensor:
        """Gradient buffer, index-aligned with ``flat_params`` (dead and TAIL slots are never written)."""
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros(self._flat.numel(), dtype=torch.float32, device=self._flat.device)
        return self._flat_grad

    @property
    def flat_grads_bucket(self) -> torch.Tensor:
        """The all-reduce bucket: the live gradients plus the TAIL slots; the data-parallel trainer puts its share of
        the loss in ``bucket[n_live]`` so gradients and loss are summed by ONE collective (26.8 MB, no dead zeros)."""
        return self.flat_grads[:self._n_live + self.TAIL]

    def _grad_views(self, flat: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        fg = self.flat_grads if flat is None else flat
        return {k: fg[o:o + n].view(shape) for k, (o, n, shape) in self._offsets.items() if k not in self._dead}

    def _param_views(self) -> Dict[str, torch.Tensor]:
        return {k: self._flat[o:o + n].view(shape) for k, (o, n, shape) in self._offsets.items()}

    def layer_grad_offset(self) -> int:
        """Offset in the flat buffers where the GCN-layer parameters start (the vertex-encoder tensors come first)."""
        return self._offsets["gcn_layers.0.w_h.weight"][0]

    def dead_mask(self) -> torch.Tensor:
        """uint8 mask over the flat buffer: 1 where the reference never produces a gradient."""
        m = torch.zeros(self._flat.numel(), dtype=torch.uint8, device=self._flat.device)
        for k in self._dead:
            o, n, _ = self._offsets[k]
            m[o:o + n] = 1
        return m

    def skip_mask(self) -> torch.Tensor:
        """What the optimizer must not touch: the dead parameters and the TAIL slots."""
        m = self.dead_mask()
        m[self._n_live:self._n_live + self.TAIL] = 1
        return m

    @property
    def _engine(self) -> E.Engine:
        if self._engine_obj is None:
            self._engine_obj = E.Engine(self.num_gcn_layers, self.cfg["gcn_edge_enabled"], self.static_edges,
                                        self.vector_edges)
        return self._engine_obj

    # ---- the reference's forward signature (drin/model.py:164) ----------------------------------
    def forward(self, batch: Sequence[torch.Tensor]) -> torch.Tensor:
        batch = tuple(batch)
        params = [p for _, p in self.named_parameters()]
        if not params[0].is_cuda:
            raise RuntimeError("drin_b200.Model must live on a CUDA device (no CPU fallback)")
        # grad mode is off inside Function.forward, so decide here whether backward state must be kept
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _DrinFunction.apply(self, batch, need_grad, *params)
