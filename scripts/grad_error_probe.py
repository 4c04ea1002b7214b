"""Per-tensor error statistics of one train step against the CPU oracle in float64 (the truth) and in fp32 (the
reference's own arithmetic): norm-wise error, and the worst element-wise ratio |a-b| / (1e-4 |b| + 1e-4 rms(b))."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200.synthetic import make_batch, spread_weights  # noqa: E402
from oracle import drin_oracle as O  # noqa: E402
from tests.helpers import elementwise_ratio  # noqa: E402


def stats(a, b):
    a, b = a.double(), b.double()
    rms = b.pow(2).mean().sqrt()
    d = (a - b).abs()
    return dict(norm=float(d.max() / b.abs().max()), ratio=float((d / (1e-4 * b.abs() + 1e-4 * rms)).max()),
                ratio_structured=float(elementwise_ratio(a, b, floor_mult=1.0).max()),
                viol=int((d > 1e-4 * b.abs() + 1e-4 * rms).sum()), rms_err_over_rms=float(d.pow(2).mean().sqrt() / rms))


def run(dataset, B, cands, seed):
    cfg = O.DrinConfig(num_candidates_model=cands + 1)
    batch = make_batch(dataset, B, seed, cands)
    sd = spread_weights(O.init_state(cfg, 0))
    s32, l32, g32 = O.train_step_grads(sd, batch[:-1], batch[-1], cfg)
    sd64 = {k: v.double() for k, v in sd.items()}
    b64 = [t.double() if t.is_floating_point() else t for t in batch]
    s64, l64, g64 = O.train_step_grads(sd64, b64[:-1], b64[-1], cfg)
    m = drin_b200.Model(num_candidates_model=cands + 1)
    m.load_state_dict(sd)
    m = m.cuda()
    db = [t.cuda() for t in batch]
    scores = m(db[:-1])
    loss = drin_b200.TripletLoss(cfg.triplet_margin)(db[-1], scores)
    loss.backward()
    out = {"scores": dict(ours_vs_64=stats(scores.detach().cpu(), s64), ref32_vs_64=stats(s32, s64))}
    for k, p in m.named_parameters():
        if g64[k] is None:
            continue
        out[k] = dict(ours_vs_64=stats(p.grad.cpu(), g64[k]), ref32_vs_64=stats(g32[k], g64[k]),
                      ours_vs_32=stats(p.grad.cpu(), g32[k]))
    print(f"== {dataset} B={B} C={cands + 1}")
    for k, v in out.items():
        o, r = v["ours_vs_64"], v["ref32_vs_64"]
        print(f"{k[-44:]:44s} ours: norm {o['norm']:.1e} ratio {o['ratio']:.2f} struct {o['ratio_structured']:.2f} viol {o['viol']} rmsErr {o['rms_err_over_rms']:.1e}"
              f" | ref32: norm {r['norm']:.1e} ratio {r['ratio']:.2f}", flush=True)
    return out


if __name__ == "__main__":
    res = {"wm148": run("wikimel", 148, 100, 179), "wd1280": run("wikidiverse", 1280, 10, 21)}
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "grad_error_probe.json"), "w"), indent=1)
