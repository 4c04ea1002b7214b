"""Deterministic synthetic batches in the layout of the reference's data loader.

The 15-tuple mirrors ``MELData.__getitem__`` + default collate (reference drin/data.py:85-126):
14 model inputs followed by the uint8 one-hot ``answer`` rows.  Shapes follow the cached-feature
files written by the reference's preprocessing (preprocess/bert.py:94-108, resnet.py:97-99,162,
clip.py:143); value distributions follow SURVEY.md section 8(d).
"""
from __future__ import annotations

from typing import List, Optional

import torch

WIKIDIVERSE = "wikidiverse"
WIKIMEL = "wikimel"


def make_batch(
    dataset: str = WIKIDIVERSE,
    batch_size: int = 32,
    seed: int = 0,
    num_candidates: Optional[int] = None,     # real candidates per mention (C - 1)
    mention_tokens: int = 128,                # args.py:72
    entity_tokens: int = 64,                  # args.py:85 (WikiMEL only)
    bert_dim: int = 768,
    resnet_dim: int = 2048,
    regions: int = 49,                        # args.py:53
    mention_objects: int = 3,                 # args.py:57
    entity_objects: int = 1,
    signed_images: bool = False,              # N(0,1) variant instead of post-ReLU |N(0,1)|
    device: str = "cpu",
    pin: bool = False,
    generate_on_device: bool = False,
) -> List[torch.Tensor]:
    if num_candidates is None:
        num_candidates = 10 if dataset == WIKIDIVERSE else 100
    B, C = batch_size, num_candidates + 1       # +1: the appended gold slot (prepare.py:86,181)
    # device="cpu": the values every parity test and golden fixture is built on.  A CUDA device generates
    # on the GPU (same distributions, different values) so benchmarks need not spend minutes in CPU randn.
    gen_dev = "cpu" if (device == "cpu" or not generate_on_device) else device
    g = torch.Generator(device=gen_dev).manual_seed(seed)
    _randn, _rand, _randint, _arange, _eye, _zeros, _ones = (torch.randn, torch.rand, torch.randint, torch.arange,
                                                             torch.eye, torch.zeros, torch.ones)

    def randn(*s):
        return _randn(*s, generator=g, device=gen_dev)

    def img(*s):
        x = randn(*s)
        return x if signed_images else x.abs()

    def scores(*s):
        x = _rand(*s, generator=g, device=gen_dev)
        return x * (_rand(*s, generator=g, device=gen_dev) >= 0.1)     # ~10 % exact zeros (resnet.py:117-118)

    mtf = randn(B, mention_tokens, bert_dim)
    mmask = _ones(B, mention_tokens, dtype=torch.int64, device=gen_dev)
    start = _randint(1, 20, (B,), generator=g, device=gen_dev)
    end = torch.clamp(start + _randint(1, 6, (B,), generator=g, device=gen_dev), max=mention_tokens)
    mif = img(B, regions, resnet_dim)
    mof = img(B, mention_objects, 1, resnet_dim)
    mos = scores(B, mention_objects)
    if dataset == WIKIDIVERSE:
        etf = randn(B, C, bert_dim)
        emask = _zeros(B, dtype=torch.int64, device=gen_dev)            # collated int 0 (data.py:86)
        eif = img(B, C, resnet_dim)
        eof = img(B, C, entity_objects, resnet_dim)
    elif dataset == WIKIMEL:
        etf = randn(B, C, entity_tokens, bert_dim)
        n = _randint(4, entity_tokens + 1, (B, C, 1), generator=g, device=gen_dev)
        emask = (_arange(entity_tokens, device=gen_dev).view(1, 1, -1) < n).to(torch.int64)
        eif = img(B, C, 1, resnet_dim)
        eof = img(B, C, entity_objects, 1, resnet_dim)
    else:
        raise ValueError(f"unknown dataset {dataset!r}")
    eos = scores(B, C, entity_objects)
    miet = 20 + 5 * randn(B, C)
    mtei = 20 + 5 * randn(B, C)
    # answer in 0..C-1; value C-1 = "gold not among the candidates" -> all-zero label row (data.py:159-161)
    ans = _randint(0, C, (B,), generator=g, device=gen_dev)
    onehot = torch.cat([_eye(C - 1, dtype=torch.uint8, device=gen_dev), _zeros(1, C - 1, dtype=torch.uint8, device=gen_dev)])
    y = onehot[ans]
    out = [mtf, mmask, start, end, mif, mof, mos, etf, emask, eif, eof, eos, miet, mtei, y]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu" and gen_dev == "cpu":
        out = [t.to(device, non_blocking=True) for t in out]
    return out


def batch_bytes(batch) -> int:
    return sum(t.numel() * t.element_size() for t in batch)


def spread_weights(state: dict, scale: float = 3.0) -> dict:
    """Weight set (B) of SURVEY 8(d): default init with W_h scaled so scores spread past the triplet
    margin and both hinge branches / the cross-batch coupling are exercised."""
    out = {k: v.clone() for k, v in state.items()}
    for k in out:
        if k.endswith("w_h.weight"):
            out[k] *= scale
    return out
