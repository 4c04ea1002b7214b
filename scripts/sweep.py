"""BASELINE.json configs[2] and configs[4] on one B200: WikiMEL-shaped ranking (100 candidates + gold slot) and the
sweep "text tokens 32 -> 128, candidates 10 -> 100", with the GEMM (tensor) and front-end (HBM) stage rooflines.

    python scripts/sweep.py [--out gpurun_out/sweep.json] [--quick]

Each point: synthetic features generated on the device, 3 warm-up + 5 timed iterations (CUDA events), per-stage
device time from the library's profile hooks (drin_profile_collect).  Inputs are sized >= 1 GiB so every
iteration streams from HBM, not L2.  Numbers are per GPU; ranking shards mentions with no communication.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from bench import algorithmic_bytes_per_mention, gemm_flops_per_mention  # noqa: E402
from drin_b200 import _lib  # noqa: E402
from drin_b200.synthetic import batch_bytes, make_batch  # noqa: E402

STAGES = ["gemm", "frontend", "gcn_fwd", "gcn_bwd", "score", "loss", "adam", "prep"]


def collect(lib):
    n = len(STAGES)
    ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    _lib.check(lib.drin_profile_collect(ms, fl, by, cnt), "drin_profile_collect")
    return {s: dict(ms=ms[i], flops=fl[i]) for i, s in enumerate(STAGES)}


def timed(fn, iters=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_point(lib, peaks, dataset, cands, Lm, Le, budget_bytes, train):
    wm = dataset == "wikimel"
    Cn = cands + 1
    per = (Lm * 768 + 49 * 2048 + 3 * 2048) * 4 + Cn * ((Le if wm else 1) * 768 + 2 * 2048) * 4
    B = max(64, min(4096, int(budget_bytes // per) // 64 * 64))
    kw = dict(mention_tokens=Lm)
    if wm:
        kw["entity_tokens"] = Le
    batch = make_batch(dataset, B, seed=7, num_candidates=cands, device="cuda", generate_on_device=True, **kw)
    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=Cn).cuda()
    tr = drin_b200.Trainer(model)
    out = dict(dataset=dataset, candidates=cands, mention_tokens=Lm, entity_tokens=Le if wm else 0, batch=B,
               input_mib=batch_bytes(batch) / 2**20)
    nbar = (4 + Le) / 2.0 if wm else 1.0          # generator: n ~ U{4..Le}
    abytes = algorithmic_bytes_per_mention(Cn, wm, Le=Le, nbar=nbar)
    for mode in (["rank", "train"] if train else ["rank"]):
        fn = (lambda: tr.rank_scores(batch)) if mode == "rank" else (lambda: tr.step(batch))
        ms = timed(fn)
        lib.drin_profile_enable(1)
        fn()
        torch.cuda.synchronize()
        collect(lib)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        prof = collect(lib)
        lib.drin_profile_enable(0)
        g, fe = prof["gemm"], prof["frontend"]
        gemm_tf = g["flops"] / (g["ms"] * 1e-3) / 1e12
        fe_gbs = abytes * B * 3 / (fe["ms"] * 1e-3) / 1e9
        out[mode] = dict(
            mentions_per_s=B / (ms * 1e-3), ms=ms,
            gemm_tflops_algorithmic=gemm_tf, gemm_frac_of_tensor_peak=gemm_tf / peaks["tensor"],
            frontend_gbs_algorithmic=fe_gbs, frontend_frac_of_hbm_peak=fe_gbs / peaks["hbm"],
            necessary_gemm_tflops=gemm_flops_per_mention(Cn, mode == "train") * B / (ms * 1e-3) / 1e12,
            stage_ms={k: v["ms"] / 3 for k, v in prof.items() if v["ms"] > 0})
    del batch, model, tr
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = _lib.load()
    peaks = {"tensor": 1377.5, "hbm": 6536.0}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks = {"tensor": p.get("bf16_tflops_sustained", 1377.5), "hbm": p.get("hbm_gbs", 6536.0)}
    except Exception:
        pass
    points = []
    # configs[2]: WikiMEL top-100 ranking shape; configs[4]: candidates 10 -> 100 and tokens 32 -> 128
    grid = [("wikimel", 100, 128, 64, True)]
    if not args.quick:
        grid += [("wikidiverse", c, 128, 0, True) for c in (10, 25, 50, 100)]
        grid += [("wikimel", c, 128, 64, False) for c in (10, 25, 50)]
        grid += [("wikimel", 100, lm, le, False) for lm, le in ((32, 32), (64, 64), (128, 128))]
        grid += [("wikidiverse", 10, lm, 0, False) for lm in (32, 64)]
    for ds, c, lm, le, train in grid:
        budget = (12 if ds == "wikimel" else 3) * 2**30     # WikiMEL rows are 22 MB / mention: keep B in the hundreds
        r = run_point(lib, peaks, ds, c, lm, le, budget, train)
        print(json.dumps(r), flush=True)
        points.append(r)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(dict(peaks=peaks, note="per-GPU, device-resident inputs, fp32-parity mode", points=points), fh, indent=1)
    small = small_batch_points()
    with open(args.out, "w") as fh:
        json.dump(dict(peaks=peaks, note="per-GPU, device-resident inputs, fp32-parity mode", points=points,
                       small_batch=small), fh, indent=1)


def small_batch_points():
    """The reference's own batch sizes (common/args.py:118,126 default 64; BASELINE configs[0] uses 32): one train step
    over a resident FeatureStore, eager (one C-ABI call per stage, ~60 launches) vs one captured CUDA graph."""
    from drin_b200.store import FeatureStore, synthetic_tables

    out = []
    for ds, cands in (("wikidiverse", 10), ("wikimel", 100)):
        n = 1024 if ds == "wikidiverse" else 256
        store = FeatureStore(ds, synthetic_tables(ds, n, 3, cands, device="cuda"), cands + 1, device="cuda")
        for B in (32, 64, 256):
            if B > n:
                continue
            torch.manual_seed(0)
            tr = drin_b200.Trainer(drin_b200.Model(num_candidates_model=cands + 1).cuda())
            idx = torch.randperm(n)[:B]
            ms_eager = timed(lambda: tr.step(store.select(idx)), iters=20, warmup=5)
            gs = drin_b200.GraphedStoreStep(tr, store, B)
            ms_graph = timed(lambda: gs.step(idx), iters=20, warmup=5)
            r = dict(dataset=ds, candidates=cands, batch=B, eager_ms=ms_eager, graph_ms=ms_graph,
                     eager_mentions_per_s=B / (ms_eager * 1e-3), graph_mentions_per_s=B / (ms_graph * 1e-3))
            print(json.dumps(r), flush=True)
            out.append(r)
            del gs, tr
        del store
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
