"""Generate tests/golden/*.pt from the UNMODIFIED reference -- run in the build container only:

    python oracle/make_golden.py

For every case the reference's own ``drin.model.Model`` (imported from /root/reference by
``oracle/ref_import.py``) and ``common.utils.TripletLoss`` / ``TopkAccuracy`` are run on a seeded
synthetic batch; the outputs are stored as small fixtures.  Inputs and weights are NOT stored: they
are regenerated from the seed by ``drin_b200.synthetic.make_batch`` / ``oracle.drin_oracle.init_state``
and guarded by float64 checksums kept in the fixture.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from drin_b200.synthetic import make_batch, spread_weights  # noqa: E402
from oracle import drin_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

CASES = [
    # name, dataset, B, cands, kwargs for make_batch, weight set, arg overrides
    dict(name="wd_b8_init", dataset="wikidiverse", B=8, cands=10, seed=0, weights="init"),
    dict(name="wd_b8_spread", dataset="wikidiverse", B=8, cands=10, seed=1, weights="spread"),
    dict(name="wd_b16_signed", dataset="wikidiverse", B=16, cands=10, seed=2, weights="spread",
         batch_kw=dict(signed_images=True)),
    dict(name="wd_b5_c4", dataset="wikidiverse", B=5, cands=3, seed=3, weights="spread"),
    dict(name="wm_b4_init", dataset="wikimel", B=4, cands=100, seed=0, weights="init"),
    dict(name="wm_b3_spread", dataset="wikimel", B=3, cands=100, seed=4, weights="spread"),
    dict(name="wm_b6_c6_le16", dataset="wikimel", B=6, cands=5, seed=5, weights="spread",
         batch_kw=dict(entity_tokens=16, mention_tokens=32), entity_tokens=16),
    dict(name="wd_b8_edge_off", dataset="wikidiverse", B=8, cands=10, seed=6, weights="spread",
         overrides=dict(gcn_edge_enabled=[1, 0, 1, 1])),
    dict(name="wd_b8_layers1", dataset="wikidiverse", B=8, cands=10, seed=7, weights="spread",
         overrides=dict(num_gcn_layers=1)),
    dict(name="wd_b8_margin005", dataset="wikidiverse", B=8, cands=10, seed=9, weights="spread",
         overrides=dict(triplet_margin=0.05)),
    dict(name="wd_b8_layers3", dataset="wikidiverse", B=8, cands=10, seed=8, weights="spread",
         overrides=dict(num_gcn_layers=3)),
    dict(name="wd_b8_static", dataset="wikidiverse", B=8, cands=10, seed=10, weights="spread",
         overrides=dict(gcn_edge_type="static")),
    dict(name="wd_b6_static_l3_mask", dataset="wikidiverse", B=6, cands=10, seed=11, weights="spread",
         overrides=dict(gcn_edge_type="static", num_gcn_layers=3, gcn_edge_enabled=[1, 1, 0, 1])),
    # gcn_edge_feature = "vector" (args.py:33): [B, C, D] edges, W_m edge update, D/2-wide W_u / W_v
    dict(name="wd_b8_vector", dataset="wikidiverse", B=8, cands=10, seed=12, weights="spread",
         overrides=dict(gcn_edge_feature="vector")),
    dict(name="wd_b6_vector_l3_mask", dataset="wikidiverse", B=6, cands=10, seed=13, weights="spread",
         overrides=dict(gcn_edge_feature="vector", num_gcn_layers=3, gcn_edge_enabled=[1, 0, 1, 1])),
    dict(name="wm_b3_c6_vector", dataset="wikimel", B=3, cands=5, seed=14, weights="spread",
         batch_kw=dict(entity_tokens=16, mention_tokens=32), entity_tokens=16,
         overrides=dict(gcn_edge_feature="vector")),
    dict(name="wd_b5_vector_static", dataset="wikidiverse", B=5, cands=10, seed=15, weights="spread",
         overrides=dict(gcn_edge_feature="vector", gcn_edge_type="static")),
    dict(name="wd_b4_vector_l1", dataset="wikidiverse", B=4, cands=10, seed=16, weights="spread",
         overrides=dict(gcn_edge_feature="vector", num_gcn_layers=1)),
]


def checksum(tensors):
    """Order-independent, exact: md5 of the raw bytes (a float sum depends on the thread count)."""
    import hashlib
    return [hashlib.md5(t.contiguous().numpy().tobytes()).hexdigest() for t in tensors]


def run_case(case):
    ov = case.get("overrides", {})
    m, u, a = ref_import.load(case["dataset"], case["cands"], case.get("entity_tokens", 64), **ov)
    cfg = O.DrinConfig(num_candidates_model=case["cands"] + 1,
                       num_gcn_layers=ov.get("num_gcn_layers", 2),
                       gcn_edge_enabled=tuple(ov.get("gcn_edge_enabled", (1, 1, 1, 1))),
                       gcn_edge_type=ov.get("gcn_edge_type", "dynamic"),
                       gcn_edge_feature=ov.get("gcn_edge_feature", "scaler"),
                       triplet_margin=ov.get("triplet_margin", 0.25))
    batch = make_batch(case["dataset"], case["B"], case["seed"], case["cands"], **case.get("batch_kw", {}))
    sd = O.init_state(cfg, seed=0)
    torch.manual_seed(0)
    model = m.Model()
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == O.state_dict_keys(cfg), "state_dict key order drifted"
    for k in ref_sd:   # init_state must reproduce the reference's seeded init bit for bit
        assert torch.equal(ref_sd[k], sd[k]), k
    if case["weights"] == "spread":
        sd = spread_weights(sd)
    model.load_state_dict(sd)
    y = batch[-1]
    scores = model(batch[:-1])
    loss = u.TripletLoss(a.triplet_margin)(y, scores)
    loss.backward()
    grads = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            grads[k] = None
        else:
            gflat = p.grad.flatten()
            grads[k] = dict(norm=float(gflat.double().norm()), sum=float(gflat.double().sum()),
                            head=gflat[:32].clone(), tail=gflat[-32:].clone())
    hits = {}
    for k in (1, 3, 5):
        if k <= case["cands"]:
            met = u.TopkAccuracy(k)
            met.update(scores.detach(), y)
            hits[k] = int(met.correct)
    with torch.no_grad():
        model2 = m.Model()
        model2.load_state_dict(sd)
        scores_nograd = model2(batch[:-1])
    assert torch.equal(scores_nograd, scores.detach())
    return dict(
        case={k: v for k, v in case.items()},
        torch_version=torch.__version__,
        input_checksum=checksum(batch),
        weight_checksum=checksum(list(sd.values())),
        scores=scores.detach().clone(),
        loss=float(loss),
        loss_tensor=loss.detach().clone(),
        grads=grads,
        topk_hits=hits,
    )


def loss_cases():
    """TripletLoss / TopkAccuracy of the reference on spread-out random scores (both hinge branches,
    cross-batch coupling, all-zero label rows, exact ties)."""
    _, u, _ = ref_import.load("wikidiverse", 10)
    out = []
    g = torch.Generator().manual_seed(123)
    for B, C, margin in ((7, 11, 0.25), (32, 11, 0.25), (5, 101, 0.25), (16, 6, 0.5), (1, 11, 0.25)):
        s = (torch.rand(B, C, generator=g) * 2 - 1).requires_grad_(True)
        with torch.no_grad():
            s[:, 1] = s[:, 0]                      # exact ties between two candidates
        ans = torch.randint(0, C, (B,), generator=g)
        onehot = torch.cat([torch.eye(C - 1, dtype=torch.uint8), torch.zeros(1, C - 1, dtype=torch.uint8)])
        y = onehot[ans]
        loss = u.TripletLoss(margin)(y, s)
        loss.backward()
        hits = {}
        for k in (1, 3, 5):
            met = u.TopkAccuracy(k)
            met.update(s.detach(), y)
            hits[k] = int(met.correct)
        out.append(dict(scores=s.detach().clone(), labels=y, margin=margin, loss=loss.detach().clone(),
                        dscores=s.grad.clone(), topk_hits=hits))
    return out


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])            # optional: regenerate only the named cases
    if not only:
        torch.save(loss_cases(), os.path.join(out_dir, "triplet_loss_cases.pt"))
    for case in CASES:
        if only and case["name"] not in only:
            continue
        fx = run_case(case)
        path = os.path.join(out_dir, case["name"] + ".pt")
        torch.save(fx, path)
        print(f"{case['name']:>18s}  loss={fx['loss']:.6f}  scores[{tuple(fx['scores'].shape)}] "
              f"range=({fx['scores'].min():.4f},{fx['scores'].max():.4f})  "
              f"none_grads={sum(v is None for v in fx['grads'].values())}  {os.path.getsize(path)} B")


if __name__ == "__main__":
    main()
