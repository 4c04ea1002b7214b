"""GPU bring-up: forward of the CUDA path vs the CPU oracle, stage by stage."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drin_b200 import engine as E  # noqa: E402
from drin_b200.synthetic import make_batch, spread_weights  # noqa: E402
from oracle import drin_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(dataset, B, cands, seed, layers=2, enabled=(1, 1, 1, 1), bf16=False, **kw):
    cfg = O.DrinConfig(num_candidates_model=cands + 1, num_gcn_layers=layers, gcn_edge_enabled=enabled)
    batch = make_batch(dataset, B, seed, cands, **kw)
    sd = spread_weights(O.init_state(cfg, 0))
    feats = (0, 4, 5, 7, 9, 10)
    if bf16:
        batch = [t.to(torch.bfloat16).float() if i in feats else t for i, t in enumerate(batch)]
    V = O.vertex_encode(sd, batch[:-1])
    tt, ii = O.edge_encode(batch[:-1])
    Ed = [tt, batch[13] / 100, batch[12] / 100, ii]
    ref = dict(x0=torch.cat([V[0], V[1], V[2].flatten(0, 1), V[3].flatten(0, 1)]), edges0=torch.stack([e.flatten() for e in Ed]))
    Vl, El = V, Ed
    per_layer = []
    for l in range(layers):
        Vl, El = O.gcn_layer(sd, l, cfg, Vl, El)
        per_layer.append((Vl, El))
    scores_ref = O.cosine(Vl[0].unsqueeze(1), Vl[2])

    eng = E.Engine(layers, enabled)
    dbatch = [t.cuda() for t in batch[:-1]]
    if bf16:
        dbatch = [t.to(torch.bfloat16) if i in feats else t for i, t in enumerate(dbatch)]
    dsd = {k: v.cuda() for k, v in sd.items()}
    scores, ctx = eng.forward(dbatch, dsd, training=False)
    torch.cuda.synchronize()
    out = dict(case=f"{dataset} B={B} C={cands+1} L={layers} bf16={bf16} en={enabled}")
    out["x0"] = rel(eng.debug_buffer(ctx, "x0"), ref["x0"])
    out["edges0"] = rel(eng.debug_buffer(ctx, "edges0"), ref["edges0"])
    for l in range(layers):
        h = eng.debug_buffer(ctx, "h", l).cpu()
        k = O.gcn_keys(l)
        act = F.gelu(F.layer_norm(h, (h.shape[-1],), sd[k["ln_w"]], sd[k["ln_b"]], 1e-5))
        Vr, Er = per_layer[l]
        if l < layers - 1:
            want = torch.cat([Vr[0], Vr[1], Vr[2].flatten(0, 1), Vr[3].flatten(0, 1)])
            out[f"edges_out{l}"] = rel(eng.debug_buffer(ctx, "edges_out", l), torch.stack([e.flatten() for e in Er]))
        else:
            want = torch.cat([Vr[0], Vr[2].flatten(0, 1)])
        out[f"act{l}"] = rel(act, want)
    out["scores"] = rel(scores, scores_ref)
    out["rank_equal"] = bool(torch.equal(O.ranking(scores.cpu()), O.ranking(scores_ref)))
    out["nan"] = int(torch.isnan(scores).sum())
    return out


if __name__ == "__main__":
    results = []
    cases = [
        dict(dataset="wikidiverse", B=8, cands=10, seed=1),
        dict(dataset="wikimel", B=3, cands=100, seed=4),
        dict(dataset="wikidiverse", B=37, cands=10, seed=2, signed_images=True),
        dict(dataset="wikidiverse", B=8, cands=10, seed=7, layers=1),
        dict(dataset="wikidiverse", B=8, cands=10, seed=8, layers=3),
        dict(dataset="wikidiverse", B=8, cands=10, seed=6, enabled=(1, 0, 1, 1)),
        dict(dataset="wikimel", B=6, cands=5, seed=5, entity_tokens=16, mention_tokens=32),
        dict(dataset="wikidiverse", B=16, cands=10, seed=1, bf16=True),
        dict(dataset="wikimel", B=3, cands=100, seed=4, bf16=True),
        dict(dataset="wikidiverse", B=300, cands=10, seed=3),
    ]
    for kw in cases:
        try:
            r = run(**kw)
        except Exception as e:  # noqa
            r = dict(case=str(kw), error=repr(e)[:800])
        print(json.dumps(r), flush=True)
        results.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "forward_check.json"), "w"), indent=1)
