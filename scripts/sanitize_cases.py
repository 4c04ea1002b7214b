"""Workload for compute-sanitizer (scripts/sanitize.sh): train steps of the golden subset with every kernel variant
forced, plus at-scale steps that make the engine pick its full-size kernels itself (warp-per-mention row kernels,
sliced WikiMEL kernels + finish kernels, cta_group::2 GEMMs, column-wise first-layer backward), checked against the oracle
so a sanitizer-clean run is also a correct one.  Usage: python scripts/sanitize_cases.py [small|scale|all]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200 import _lib  # noqa: E402
from drin_b200.synthetic import make_batch, spread_weights  # noqa: E402
from oracle import drin_oracle as O  # noqa: E402

VARIANTS = ("score_bwd_variant", "score_fwd_variant", "layer_fwd_variant", "layer_bwd_variant")


def option(name, v):
    _lib.check(_lib.load().drin_debug_option(name.encode(), C.c_int32(v)), name)


def step(dataset, B, cands, seed, check=True, bf16=False, **cfg_kw):
    cfg = O.DrinConfig(num_candidates_model=cands + 1, **cfg_kw)
    kw = dict(entity_tokens=16, mention_tokens=32) if (dataset == "wikimel" and cands < 100) else {}
    batch = make_batch(dataset, B, seed, cands, **kw)
    sd = spread_weights(O.init_state(cfg, 0))
    m = drin_b200.Model(num_candidates_model=cands + 1, num_gcn_layers=cfg.num_gcn_layers,
                        gcn_edge_feature=cfg.gcn_edge_feature, gcn_edge_type=cfg.gcn_edge_type)
    m.load_state_dict(sd)
    m = m.cuda()
    feats = (0, 4, 5, 7, 9, 10)
    db = [t.cuda().to(torch.bfloat16) if (bf16 and i in feats) else t.cuda() for i, t in enumerate(batch)]
    tr = drin_b200.Trainer(m)
    loss = tr.step(db)
    tr.rank_scores(db)
    met = drin_b200.TopkAccuracy([1, 3])
    met.update(tr.last_scores, db[-1])
    torch.cuda.synchronize()
    if check and not bf16:
        s_ref, l_ref, _ = O.train_step_grads(sd, batch[:-1], batch[-1], cfg)
        err = float((tr.last_scores.cpu() - s_ref).abs().max())
        assert err < 1e-4 and abs(float(loss) - float(l_ref)) < 1e-4 * abs(float(l_ref)), (err, float(loss), float(l_ref))
    print(f"ok {dataset} B={B} C={cands + 1} {cfg_kw} bf16={bf16} loss={float(loss):.6f}", flush=True)


def small():
    for forced in (-1, 1, 2):
        for name in VARIANTS:
            option(name, forced)
        step("wikidiverse", 8, 10, 1)
        step("wikimel", 3, 100, 4)
        step("wikimel", 6, 5, 5)
    for name in VARIANTS:
        option(name, -1)
    step("wikidiverse", 8, 10, 12, gcn_edge_feature="vector")
    step("wikidiverse", 6, 10, 13, gcn_edge_feature="vector", num_gcn_layers=3)
    step("wikidiverse", 8, 10, 10, gcn_edge_type="static")
    step("wikidiverse", 8, 10, 2, bf16=True)
    step("wikimel", 3, 100, 4, bf16=True)


def scale():
    step("wikidiverse", 1280, 10, 21)
    step("wikimel", 148, 100, 22)
    step("wikidiverse", 1300, 10, 23, gcn_edge_feature="vector")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("small", "all"):
        small()
    if what in ("scale", "all"):
        scale()
    print("SANITIZE_CASES_DONE", flush=True)
