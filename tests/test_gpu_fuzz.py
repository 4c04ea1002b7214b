"""Seeded shape / configuration fuzzing of the whole train step against the CPU oracle: odd batch sizes (1, primes),
2..24 candidate slots, 1-4 GCN layers, edge masks, static / dynamic edges, scalar / vector edges, both dataset layouts, ragged WikiMEL entity
lengths (down to the 3-token minimum) and both kernel families (CTA-per-mention and warp-per-mention)."""
import ctypes as C
import random

import pytest
import torch

import drin_b200
from drin_b200 import _lib
from drin_b200.synthetic import make_batch, spread_weights
from oracle import drin_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
OPTIONS = ("score_bwd_variant", "score_fwd_variant", "layer_fwd_variant", "layer_bwd_variant")


def _cases():
    rng = random.Random(20251018)
    out = []
    for i in range(24):
        wm = i % 3 == 2
        out.append(dict(
            dataset="wikimel" if wm else "wikidiverse",
            B=rng.choice([1, 2, 3, 5, 7, 13, 17, 33]),
            cands=rng.choice([1, 2, 3, 6, 10, 23]) if not wm else rng.choice([1, 4, 9]),
            layers=rng.choice([1, 2, 2, 3]),
            mask=rng.choice([(1, 1, 1, 1), (1, 1, 1, 1), (0, 1, 1, 1), (1, 0, 0, 1), (1, 1, 1, 0)]),
            static=rng.random() < 0.25,
            variant=rng.choice([-1, 1]),
            seed=100 + i,
            kw=dict(entity_tokens=rng.choice([4, 5, 16]), mention_tokens=rng.choice([24, 32])) if wm else {},
        ))
    for i, c in enumerate(out):
        c["vector"] = False
        if i % 4 == 1:
            c["variant"] = 2          # warp kernels + the column-wise first-layer backward (drawn values stay untouched)
    rng = random.Random(20251019)            # gcn_edge_feature="vector" cases (appended: the cases above keep their draws)
    for i in range(10):
        wm = i % 3 == 2
        out.append(dict(
            dataset="wikimel" if wm else "wikidiverse",
            B=rng.choice([1, 2, 3, 5, 7, 13, 33]),
            cands=rng.choice([1, 2, 3, 6, 10, 23]) if not wm else rng.choice([1, 4, 9]),
            layers=rng.choice([2, 2, 3, 4]),
            mask=rng.choice([(1, 1, 1, 1), (1, 1, 1, 1), (0, 1, 1, 1), (1, 0, 0, 1), (1, 1, 1, 0)]),
            static=False, vector=True, variant=-1, seed=200 + i,
            kw=dict(entity_tokens=rng.choice([4, 5, 16]), mention_tokens=rng.choice([24, 32])) if wm else {},
        ))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c['dataset'][:5]}-B{c['B']}-C{c['cands'] + 1}-L{c['layers']}-"
                                                         f"{'st' if c['static'] else 'dy'}{'-vec' if c['vector'] else ''}-v{c['variant']}")
def test_fuzzed_train_step_matches_oracle(case):
    lib = _lib.load()
    for name in OPTIONS:
        _lib.check(lib.drin_debug_option(name.encode(), C.c_int32(case["variant"])), "drin_debug_option")
    try:
        cfg = O.DrinConfig(num_candidates_model=case["cands"] + 1, num_gcn_layers=case["layers"],
                           gcn_edge_enabled=case["mask"], gcn_edge_type="static" if case["static"] else "dynamic",
                           gcn_edge_feature="vector" if case["vector"] else "scaler")
        batch = make_batch(case["dataset"], case["B"], case["seed"], case["cands"], **case["kw"])
        if case["dataset"] == "wikimel":          # ragged entity lengths incl. the 3-token minimum (CLS, one token, SEP)
            Le = batch[8].shape[-1]
            n = torch.randint(3, Le + 1, batch[8].shape[:2] + (1,), generator=torch.Generator().manual_seed(case["seed"]))
            batch[8] = (torch.arange(Le).view(1, 1, -1) < n).to(torch.int64)
        sd = spread_weights(O.init_state(cfg, 0))
        s_ref, l_ref, g_ref = O.train_step_grads(sd, batch[:-1], batch[-1], cfg)
        model = drin_b200.Model(num_gcn_layers=cfg.num_gcn_layers, gcn_edge_enabled=cfg.gcn_edge_enabled,
                                gcn_edge_type=cfg.gcn_edge_type, gcn_edge_feature=cfg.gcn_edge_feature,
                                num_candidates_model=cfg.num_candidates_model)
        model.load_state_dict(sd)
        model = model.cuda()
        db = [t.cuda() for t in batch]
        scores = model(db[:-1])
        loss = drin_b200.TripletLoss(cfg.triplet_margin)(db[-1], scores)
        loss.backward()
        assert rel_err(scores.detach().cpu(), s_ref) < 1e-4
        assert abs(float(loss) - float(l_ref)) <= 1e-4 * abs(float(l_ref)) + 1e-7
        for k, p in model.named_parameters():
            if g_ref[k] is None:
                assert p.grad is None, k
            elif float(g_ref[k].abs().max()) > 0:
                assert rel_err(p.grad.cpu(), g_ref[k]) < 1e-4, k
            else:
                assert float(p.grad.abs().max()) < 1e-12, k
    finally:
        for name in OPTIONS:
            _lib.check(lib.drin_debug_option(name.encode(), C.c_int32(-1)), "drin_debug_option")


@pytest.mark.parametrize("dataset,B,cands,kw", [
    ("wikidiverse", 5, 6, dict(mention_objects=1, regions=7, entity_objects=1)),
    ("wikidiverse", 9, 10, dict(mention_objects=4, regions=1, entity_objects=2)),
    ("wikimel", 4, 7, dict(mention_objects=2, regions=16, entity_objects=2, entity_tokens=8, mention_tokens=32)),
    ("wikimel", 1300, 2, dict(mention_objects=3, regions=3, entity_objects=1, entity_tokens=4, mention_tokens=32)),
], ids=["wd_om1_p7", "wd_om4_p1_oe2", "wm_om2_p16_oe2", "wm_b1300_c3_tiny"])
def test_object_and_region_counts_other_than_the_defaults(dataset, B, cands, kw):
    """The loader's shapes are configuration (args.py:53,57: 49 regions, top-3 mention objects, 1 entity object); the
    kernels take them from the batch like the reference's tensor code does: 1..4 mention objects, any region count,
    several entity objects (model.py:78-92 loops over both), at a batch that dispatches the full-size kernels too."""
    cfg = O.DrinConfig(num_candidates_model=cands + 1)
    batch = make_batch(dataset, B, 77, cands, **kw)
    sd = spread_weights(O.init_state(cfg, 0))
    s_ref, l_ref, g_ref = O.train_step_grads(sd, batch[:-1], batch[-1], cfg)
    model = drin_b200.Model(num_candidates_model=cands + 1)
    model.load_state_dict(sd)
    model = model.cuda()
    db = [t.cuda() for t in batch]
    scores = model(db[:-1])
    loss = drin_b200.TripletLoss(cfg.triplet_margin)(db[-1], scores)
    loss.backward()
    assert rel_err(scores.detach().cpu(), s_ref) < 1e-4
    assert abs(float(loss) - float(l_ref)) <= 1e-4 * abs(float(l_ref)) + 1e-7
    for k, p in model.named_parameters():
        if g_ref[k] is not None:
            assert rel_err(p.grad.cpu(), g_ref[k]) < 1e-4, k
