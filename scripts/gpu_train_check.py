"""GPU bring-up: loss kernel, backward and fused train step of the CUDA path vs the CPU oracle."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200.synthetic import make_batch, spread_weights  # noqa: E402
from oracle import drin_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def loss_cases():
    out = []
    cases = torch.load(os.path.join(ROOT, "tests", "golden", "triplet_loss_cases.pt"), weights_only=False)
    for c in cases:
        s = c["scores"].cuda().requires_grad_(True)
        loss = drin_b200.TripletLoss(c["margin"])(c["labels"].cuda(), s)
        loss.backward()
        met = drin_b200.TopkAccuracy([1, 3, 5])
        met.update(s.detach(), c["labels"].cuda())
        out.append(dict(shape=list(c["scores"].shape), loss_err=abs(float(loss) - float(c["loss"])) / abs(float(c["loss"])),
                        dscore_err=rel(s.grad, c["dscores"]), hits=met.correct.tolist(), want_hits=[c["topk_hits"][k] for k in (1, 3, 5)]))
    return out


def train_case(dataset, B, cands, seed, layers=2, enabled=(1, 1, 1, 1), margin=0.25, steps=0, **kw):
    cfg = O.DrinConfig(num_candidates_model=cands + 1, num_gcn_layers=layers, gcn_edge_enabled=enabled, triplet_margin=margin)
    batch = make_batch(dataset, B, seed, cands, **kw)
    sd = spread_weights(O.init_state(cfg, 0))
    scores_ref, loss_ref, grads_ref = O.train_step_grads(sd, batch[:-1], batch[-1], cfg)
    model = drin_b200.Model(num_gcn_layers=layers, gcn_edge_enabled=enabled, num_candidates_model=cands + 1)
    model.load_state_dict(sd)
    model = model.cuda()
    dbatch = [t.cuda() for t in batch]
    scores = model(dbatch[:-1])
    loss = drin_b200.TripletLoss(margin)(dbatch[-1], scores)
    loss.backward()
    torch.cuda.synchronize()
    out = dict(case=f"{dataset} B={B} C={cands+1} L={layers} en={enabled} m={margin}")
    out["scores"] = rel(scores, scores_ref)
    out["loss"] = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
    gerr = {}
    for k, p in model.named_parameters():
        g = grads_ref[k]
        if g is None:
            gerr[k] = "None-ok" if p.grad is None else "SHOULD-BE-NONE"
        elif p.grad is None:
            gerr[k] = "MISSING"
        else:
            gerr[k] = rel(p.grad, g)
    out["grad_max"] = max(v for v in gerr.values() if isinstance(v, float))
    out["grad_bad"] = {k: v for k, v in gerr.items() if not (v == "None-ok" or (isinstance(v, float) and v < 1e-4))}
    # fused path: Trainer.forward_backward must give the same gradients; then Adam steps vs the oracle
    model2 = drin_b200.Model(num_gcn_layers=layers, gcn_edge_enabled=enabled, num_candidates_model=cands + 1)
    model2.load_state_dict(sd)
    model2 = model2.cuda()
    tr = drin_b200.Trainer(model2, lr=1e-3, margin=margin)
    l2 = tr.forward_backward(dbatch)
    out["fused_loss"] = abs(float(l2) - float(loss_ref)) / abs(float(loss_ref))
    fg = model2._grad_views()
    out["fused_grad_max"] = max(rel(fg[k], grads_ref[k]) for k in fg)
    if steps:
        params = {k: v.clone() for k, v in sd.items()}
        st = {}
        for i in range(steps):
            _, lr_, g_ = O.train_step_grads(params, batch[:-1], batch[-1], cfg)
            O.adam_step(params, g_, st)
            lg = tr.step(dbatch)
        sd2 = model2.state_dict()
        out["adam_param_err"] = max(rel(sd2[k], params[k]) for k in params)
        out["adam_last_loss"] = [float(lg), float(lr_)]
    return out


if __name__ == "__main__":
    res = dict(loss=loss_cases(), train=[])
    print(json.dumps(res["loss"]), flush=True)
    cases = [
        dict(dataset="wikidiverse", B=8, cands=10, seed=1, steps=3),
        dict(dataset="wikidiverse", B=8, cands=10, seed=9, margin=0.05),
        dict(dataset="wikimel", B=3, cands=100, seed=4),
        dict(dataset="wikidiverse", B=8, cands=10, seed=7, layers=1),
        dict(dataset="wikidiverse", B=8, cands=10, seed=8, layers=3),
        dict(dataset="wikidiverse", B=8, cands=10, seed=6, enabled=(1, 0, 1, 1)),
        dict(dataset="wikimel", B=6, cands=5, seed=5, entity_tokens=16, mention_tokens=32),
        dict(dataset="wikidiverse", B=67, cands=10, seed=3, signed_images=True),
    ]
    for kw in cases:
        try:
            r = train_case(**kw)
        except Exception as e:  # noqa
            import traceback
            r = dict(case=str(kw), error=traceback.format_exc()[-400:])
        print(json.dumps(r), flush=True)
        res["train"].append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "train_check.json"), "w"), indent=1)
