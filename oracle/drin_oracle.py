"""CPU oracle for the DRIN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product path (``drin_b200``) never does: it fails loudly when
the CUDA library is missing.

This is a closed-form fp32 PyTorch restatement of the reference's algorithm (the reference itself is
pure PyTorch; its arithmetic lives in ATen).  Each function cites the reference file:line it follows
(paths relative to the upstream repo starreeze/drin).

Parity status: the reference ships no tests / golden vectors ("parity unpinned" by its own tests, see
SURVEY.md section 4).  This oracle is therefore pinned against outputs of the reference itself, imported
unmodified in the build container by ``oracle/make_golden.py``; those outputs are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks the oracle against them.

State is a plain ``dict[str, Tensor]`` using the reference's ``state_dict`` keys, so weights can be
copied between the reference, this oracle and the CUDA module without renaming.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# state_dict keys of drin/model.py:157-162 (24 tensors, 7 875 072 elements for the default config)
K_MT_W = "vertex_encoder.mention_text_encoder.final_layer.linear.weight"
K_MT_B = "vertex_encoder.mention_text_encoder.final_layer.linear.bias"
K_ET_W = "vertex_encoder.entity_text_encoder.final_layer.weight"
K_ET_B = "vertex_encoder.entity_text_encoder.final_layer.bias"
K_MI_W = "vertex_encoder.mention_image_linear.weight"
K_MI_B = "vertex_encoder.mention_image_linear.bias"
K_EI_W = "vertex_encoder.entity_image_linear.weight"
K_EI_B = "vertex_encoder.entity_image_linear.bias"


def gcn_keys(layer: int) -> Dict[str, str]:
    p = f"gcn_layers.{layer}."
    return {
        "w_h": p + "w_h.weight", "b_h": p + "w_h.bias",
        "w_u": p + "w_u.weight", "b_u": p + "w_u.bias",
        "w_v": p + "w_v.weight", "b_v": p + "w_v.bias",
        "ln_w": p + "layer_norm.weight", "ln_b": p + "layer_norm.bias",
        "w_m": p + "w_m.weight", "b_m": p + "w_m.bias",      # gcn_edge_feature == "vector" only (model.py:112)
    }


@dataclass
class DrinConfig:
    """The subset of common/args.py the hot path reads (args.py:25-36,45,52,101)."""
    num_candidates_model: int = 11           # args.py:101 (10+1 WikiDiverse, 100+1 WikiMEL)
    gcn_embed_dim: int = 768                 # args.py:25
    num_gcn_layers: int = 2                  # args.py:26
    bert_embed_dim: int = 768                # args.py:45
    resnet_embed_dim: int = 2048             # args.py:52
    gcn_edge_enabled: Tuple[float, ...] = (1, 1, 1, 1)   # args.py:34
    gcn_edge_type: str = "dynamic"           # args.py:32 ("dynamic" | "static")
    gcn_edge_feature: str = "scaler"         # args.py:33 ("scaler" | "vector")
    triplet_margin: float = 0.25             # args.py:117,125


def state_dict_keys(cfg: DrinConfig) -> List[str]:
    """Parameter names in the reference's creation order (model.py:21-24,111-119,159-162)."""
    keys = [K_MT_W, K_MT_B, K_ET_W, K_ET_B, K_MI_W, K_MI_B, K_EI_W, K_EI_B]
    for l in range(cfg.num_gcn_layers):
        k = gcn_keys(l)
        keys += [k["w_h"], k["b_h"]]
        if cfg.gcn_edge_feature == "vector":          # w_m is a Linear only for vector edges (model.py:112)
            keys += [k["w_m"], k["b_m"]]
        keys += [k["w_u"], k["b_u"], k["w_v"], k["b_v"], k["ln_w"], k["ln_b"]]
    return keys


def init_state(cfg: DrinConfig, seed: int = 0) -> Dict[str, Tensor]:
    """torch default nn.Linear / nn.LayerNorm init in the reference's creation order, so that
    ``torch.manual_seed(seed)`` reproduces the reference's weights (model.py:19-24,109-119)."""
    torch.manual_seed(seed)
    D, R, Db = cfg.gcn_embed_dim, cfg.resnet_embed_dim, cfg.bert_embed_dim
    sd: Dict[str, Tensor] = {}

    def lin(kw, kb, out_f, in_f):
        m = torch.nn.Linear(in_f, out_f)
        sd[kw], sd[kb] = m.weight.detach().clone(), m.bias.detach().clone()

    lin(K_MT_W, K_MT_B, D, Db)      # MentionEncoder -> AvgLinear (ghmfc.py:163-165)
    lin(K_ET_W, K_ET_B, D, Db)      # EntityEncoder.final_layer (ghmfc.py:210-211)
    lin(K_MI_W, K_MI_B, D, R)       # model.py:23
    lin(K_EI_W, K_EI_B, D, R)       # model.py:24
    for l in range(cfg.num_gcn_layers):
        k = gcn_keys(l)
        lin(k["w_h"], k["b_h"], D, D)   # model.py:111
        vec = cfg.gcn_edge_feature == "vector"
        if vec:
            lin(k["w_m"], k["b_m"], D, D)   # model.py:112 (Identity, no parameters, for scalar edges)
        lin(k["w_u"], k["b_u"], D // 2 if vec else D, D)   # model.py:113-116 (list comprehension: w_u then w_v)
        lin(k["w_v"], k["b_v"], D // 2 if vec else D, D)
        sd[k["ln_w"]], sd[k["ln_b"]] = torch.ones(D), torch.zeros(D)   # model.py:119
    return sd


# --------------------------------------------------------------------------------------------
# stage functions
# --------------------------------------------------------------------------------------------
def cosine(x: Tensor, y: Tensor, eps: float = 1e-8) -> Tensor:
    """nn.CosineSimilarity(dim=-1) as ATen >= 2.0 computes it: each norm is clamped to eps before
    the division (model.py:57,162; SURVEY 8a header)."""
    xn = torch.linalg.vector_norm(x, dim=-1, keepdim=True).clamp_min(eps)
    yn = torch.linalg.vector_norm(y, dim=-1, keepdim=True).clamp_min(eps)
    return ((x / xn) * (y / yn)).sum(-1)


def span_mean(mtf: Tensor, start: Tensor, end: Tensor, loops: bool = False) -> Tensor:
    """Avg.avg, baselines/ghmfc.py:55-60: mean of the token rows start[b]:end[b] of every mention."""
    if loops:  # same per-mention loop as the reference (used when timing the CPU baseline)
        out = torch.empty(mtf.shape[0], mtf.shape[-1], dtype=mtf.dtype)
        for b in range(mtf.shape[0]):
            out[b] = mtf[b, int(start[b]):int(end[b])].mean(0)
        return out
    L = mtf.shape[1]
    pos = torch.arange(L).unsqueeze(0)
    sel = ((pos >= start.unsqueeze(1)) & (pos < end.unsqueeze(1))).to(mtf.dtype)       # [B, L]
    cnt = sel.sum(1, keepdim=True)
    return torch.einsum("bl,bld->bd", sel, mtf) / cnt        # 0/0 -> NaN like mean of empty slice


def entity_text_pool(etf: Tensor, emask: Tensor, loops: bool = False) -> Tensor:
    """EntityEncoder offline branch, baselines/ghmfc.py:237-249.
    rank 3 (WikiDiverse): features are already pooled (ghmfc.py:238-239).
    rank 4 (WikiMEL): mean over tokens 1 .. n-2 where n = sum(mask) (drops CLS and SEP)."""
    if etf.dim() == 3:
        return etf
    if loops:
        B, C = etf.shape[:2]
        out = torch.empty(B, C, etf.shape[-1], dtype=etf.dtype)
        for b in range(B):
            n = emask[b].sum(-1)
            for c in range(C):
                out[b, c] = etf[b, c, 1:int(n[c]) - 1].mean(0)
        return out
    n = emask.sum(-1, keepdim=True)                                   # [B, C, 1]
    pos = torch.arange(etf.shape[2]).view(1, 1, -1)
    sel = ((pos >= 1) & (pos < n - 1)).to(etf.dtype)                  # [B, C, Le]
    return torch.einsum("bcl,bcld->bcd", sel, etf) / sel.sum(-1, keepdim=True)


def vertex_encode(sd, batch, loops=False) -> List[Tensor]:
    """VertexEncoder.forward, drin/model.py:26-46 -> [mt, mi, et, ei]."""
    (mtf, _mmask, start, end, mif, _mof, _mos, etf, emask, eif, _eof, _eos, _miet, _mtei) = batch
    span = span_mean(mtf, start, end, loops)
    mt = F.linear(span, sd[K_MT_W], sd[K_MT_B])                       # ghmfc.py:66,69
    et = F.linear(entity_text_pool(etf, emask, loops), sd[K_ET_W], sd[K_ET_B])   # ghmfc.py:250
    mi = F.linear(mif.mean(-2), sd[K_MI_W], sd[K_MI_B])               # model.py:41-42
    if eif.dim() == 4:                                                # model.py:43-44
        eif = eif.mean(-2)
    ei = F.linear(eif, sd[K_EI_W], sd[K_EI_B])                        # model.py:45
    return [mt, mi, et, ei]


def edge_encode(batch, loops=False) -> Tuple[Tensor, Tensor]:
    """EdgeEncoder.forward, drin/model.py:60-94 -> (tt, ii), both [B, C]; uses RAW features."""
    (mtf, _mmask, start, end, _mif, mof, mos, etf, _emask, _eif, eof, eos, _miet, _mtei) = batch
    span = span_mean(mtf, start, end, loops).unsqueeze(1)             # model.py:71-72
    ecls = etf[:, :, 0] if etf.dim() == 4 else etf                    # model.py:73-75
    tt = cosine(span, ecls)                                           # model.py:76
    if mof.dim() == 4:                                                # model.py:78-79
        mof = mof.mean(-2)
    if eof.dim() == 5:                                                # model.py:82-83
        eof = eof.mean(-2)
    sim = torch.zeros(tt.shape, dtype=tt.dtype)
    den = torch.zeros(tt.shape, dtype=tt.dtype)
    for i in range(mof.shape[1]):                                     # model.py:86-91
        for j in range(eof.shape[2]):
            w = mos[:, None, i] * eos[:, :, j]
            sim = sim + cosine(mof[:, None, i], eof[:, :, j]) * w
            den = den + w
    return tt, sim / (den + 1e-9)                                     # model.py:92


def gcn_layer_vector(sd, layer: int, cfg: DrinConfig, V: List[Tensor], E: List[Tensor]):
    """GCNLayer.forward for gcn_edge_feature == "vector", drin/model.py:121-153: every edge is a [B, C, D]
    tensor, messages are elementwise products (model.py:139-146 without the scalar expand), and the dynamic
    edge update is e' = sigmoid(W_m(cat[W_u u, W_v v] + e)) with W_u, W_v: D -> D/2 (model.py:112-116,148-152)."""
    k = gcn_keys(layer)
    mt, mi, et, ei = V
    E = [e * m for e, m in zip(E, cfg.gcn_edge_enabled)]              # model.py:122
    e0, e1, e2, e3 = E                                                # order tt, ti, it, ii; each [B, C, D]
    a_mt = (e0 * et).mean(1) + (e1 * ei).mean(1)
    a_mi = (e2 * et).mean(1) + (e3 * ei).mean(1)
    a_et = e0 * mt.unsqueeze(1) + e2 * mi.unsqueeze(1)
    a_ei = e1 * mt.unsqueeze(1) + e3 * mi.unsqueeze(1)

    def upd(a, x):                                                    # model.py:128
        h = F.linear(a + x, sd[k["w_h"]], sd[k["b_h"]])
        return F.gelu(F.layer_norm(h, (h.shape[-1],), sd[k["ln_w"]], sd[k["ln_b"]], 1e-5))

    newV = [upd(a_mt, mt), upd(a_mi, mi), upd(a_et, et), upd(a_ei, ei)]
    if cfg.gcn_edge_type != "dynamic":                                # model.py:135-136
        return newV, E
    C = et.shape[1]
    fu = {0: F.linear(mt, sd[k["w_u"]], sd[k["b_u"]]), 1: F.linear(mi, sd[k["w_u"]], sd[k["b_u"]])}     # [B, D/2]
    fv = {2: F.linear(et, sd[k["w_v"]], sd[k["b_v"]]), 3: F.linear(ei, sd[k["w_v"]], sd[k["b_v"]])}     # [B, C, D/2]
    newE = []
    for e, (ui, vi) in zip(E, ((0, 2), (0, 3), (1, 2), (1, 3))):
        m = torch.cat([fu[ui].unsqueeze(1).expand(-1, C, -1), fv[vi]], dim=-1)        # model.py:150-152
        newE.append(torch.sigmoid(F.linear(m + e, sd[k["w_m"]], sd[k["b_m"]])))       # model.py:133
    return newV, newE


def gcn_layer(sd, layer: int, cfg: DrinConfig, V: List[Tensor], E: List[Tensor]):
    """GCNLayer.forward for scalar edges (dynamic or static edge type), drin/model.py:121-153."""
    if cfg.gcn_edge_feature == "vector":
        return gcn_layer_vector(sd, layer, cfg, V, E)
    k = gcn_keys(layer)
    mt, mi, et, ei = V
    C = et.shape[1]
    E = [e * m for e, m in zip(E, cfg.gcn_edge_enabled)]              # model.py:122
    e0, e1, e2, e3 = [e.unsqueeze(-1) for e in E]                     # order tt, ti, it, ii
    # model.py:102 vertex_graph + :139-146 convolute_vertex (mean over ALL C slots)
    a_mt = (e0 * et).mean(1) + (e1 * ei).mean(1)
    a_mi = (e2 * et).mean(1) + (e3 * ei).mean(1)
    a_et = e0 * mt.unsqueeze(1) + e2 * mi.unsqueeze(1)
    a_ei = e1 * mt.unsqueeze(1) + e3 * mi.unsqueeze(1)

    def upd(a, x):                                                    # model.py:128
        h = F.linear(a + x, sd[k["w_h"]], sd[k["b_h"]])
        return F.gelu(F.layer_norm(h, (h.shape[-1],), sd[k["ln_w"]], sd[k["ln_b"]], 1e-5))

    newV = [upd(a_mt, mt), upd(a_mi, mi), upd(a_et, et), upd(a_ei, ei)]
    if cfg.gcn_edge_type != "dynamic":                                # model.py:135-136: the MASKED edges pass through
        return newV, E
    # model.py:104,131-134,148-153: e' = sigmoid(mean_D(W_u u * W_v v) + e)
    fu = {0: F.linear(mt, sd[k["w_u"]], sd[k["b_u"]]), 1: F.linear(mi, sd[k["w_u"]], sd[k["b_u"]])}
    fv = {2: F.linear(et, sd[k["w_v"]], sd[k["b_v"]]), 3: F.linear(ei, sd[k["w_v"]], sd[k["b_v"]])}
    newE = []
    for e, (ui, vi) in zip(E, ((0, 2), (0, 3), (1, 2), (1, 3))):
        s = (fu[ui].unsqueeze(1) * fv[vi]).mean(-1)
        newE.append(torch.sigmoid(s + e))
    return newV, newE


def forward(sd: Dict[str, Tensor], batch: Sequence[Tensor], cfg: DrinConfig, loops: bool = False) -> Tensor:
    """Model.forward, drin/model.py:164-209 -> scores [B, C]."""
    V = vertex_encode(sd, batch, loops)
    tt, ii = edge_encode(batch, loops)
    miet, mtei = batch[12], batch[13]
    E = [tt, mtei / 100, miet / 100, ii]                              # model.py:201-204
    if cfg.gcn_edge_feature == "vector":                              # model.py:202: scalars broadcast over D
        E = [e.unsqueeze(-1).expand(-1, -1, cfg.gcn_embed_dim) for e in E]
    for l in range(cfg.num_gcn_layers):                               # model.py:205-206
        V, E = gcn_layer(sd, l, cfg, V, E)
    return cosine(V[0].unsqueeze(1), V[2])                            # model.py:207-209


def triplet_loss(y_true: Tensor, y_pred: Tensor, margin: float) -> Tensor:
    """TripletLoss.__call__, common/utils.py:35-43 in closed form.  Line 42 subtracts the ENTIRE
    [B, C-1] score matrix from mention i's positive, so the loss couples the whole batch:
        loss = 1/(B*B*(C-1)) * sum_i sum_b sum_c max(s[b,c] - p[i] + margin, 0)."""
    s = y_pred[:, :-1] if y_pred.shape[1] != y_true.shape[1] else y_pred
    p = (s * y_true.to(s.dtype)).sum(-1)                              # utils.py:38-39 (sign folded)
    t = (s.unsqueeze(0) - p.view(-1, 1, 1) + margin).clamp_min(0)     # [i, b, c]
    return t.mean((1, 2)).sum() / y_true.shape[0]


def triplet_loss_loops(y_true: Tensor, y_pred: Tensor, margin: float) -> Tensor:
    """Same loss with the reference's per-mention Python loop (CPU-baseline timing only)."""
    if y_pred.shape[1] != y_true.shape[1]:
        y_pred = y_pred[:, :-1]
    y_pred = -y_pred
    pos = torch.sum(y_pred * y_true, dim=-1)
    loss = 0.0
    for i in range(y_true.shape[0]):
        loss = loss + torch.mean(torch.maximum(pos[i] - y_pred + margin, torch.tensor(0.0)))
    return loss / y_true.shape[0]


def topk_hits(y_pred: Tensor, y_true: Tensor, k: int) -> int:
    """TopkAccuracy.update, common/utils.py:60-66: gold is a hit when its score >= the k-th largest
    score of the row (ties count as hits); the appended gold slot is excluded."""
    s = y_pred[:, :-1] if y_pred.shape[1] != y_true.shape[1] else y_pred
    thr = torch.topk(s, k).values[:, -1:]
    return int((y_true.to(torch.int64) * (s >= thr)).sum())


def ranking(y_pred: Tensor) -> Tensor:
    """Candidate order per mention (descending score, stable) over the real candidates."""
    return torch.argsort(y_pred[:, :-1], dim=-1, descending=True, stable=True)


def train_step_grads(sd, batch, y_true, cfg: DrinConfig, loops: bool = False):
    """train.py:32-34 + loss.backward(): returns (scores, loss, {key: grad or None})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    scores = forward(leaves, batch, cfg, loops)
    loss = (triplet_loss_loops if loops else triplet_loss)(y_true, scores, cfg.triplet_margin)
    loss.backward()
    return scores.detach(), loss.detach(), {k: v.grad for k, v in leaves.items()}


def adam_step(params: Dict[str, Tensor], grads: Dict[str, Tensor], state: dict, lr=1e-3,
              betas=(0.9, 0.999), eps=1e-8) -> None:
    """torch.optim.Adam defaults (train.py:55-56); parameters whose grad is None are skipped."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(p))
        v = state.setdefault("v", {}).setdefault(k, torch.zeros_like(p))
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def triplet_sharded(scores_all: Tensor, labels_all: Tensor, margin: float, row0: int, rows: int):
    """Data-parallel form of the loss, closed form.  Returns (this shard's share of the global
    loss, d(global loss)/d(scores of rows row0..row0+rows)) with
        dL/ds[b,c] = k * (A[b,c] - y[b,c] * N_b),  k = 1 / (B * B * (C-1)),
        A[b,c] = #{i : s[b,c] - p_i + margin > 0},  N_b = #{(b',c') : s[b',c'] - p_b + margin > 0}."""
    s = scores_all[:, :-1]
    y = labels_all.to(s.dtype)
    B, Cm1 = s.shape
    k = 1.0 / (B * B * Cm1)
    p = (s * y).sum(-1)
    sl = s[row0:row0 + rows]
    active = (sl.unsqueeze(-1) - p.view(1, 1, -1) + margin) > 0                  # [rows, C-1, B]
    share = k * (sl.unsqueeze(-1) - p.view(1, 1, -1) + margin).clamp_min(0).sum()
    A = active.sum(-1).to(s.dtype)
    N = ((s.unsqueeze(0) - p[row0:row0 + rows].view(-1, 1, 1) + margin) > 0).sum((1, 2)).to(s.dtype)
    d = k * (A - y[row0:row0 + rows] * N.unsqueeze(1))
    return share, torch.cat([d, torch.zeros(rows, 1, dtype=s.dtype)], dim=1)
