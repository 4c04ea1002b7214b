"""A/B of kernel launch options on one B200 (the test hooks of drin_debug_option):
    python scripts/ab.py [--edge-feature vector] [--batch 4096] --option vec_ctas_per_sm --values 0,2,3,4,5,6,8
prints ms/step and the per-stage device times of the train step for every value of the option."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200 import _lib  # noqa: E402
from drin_b200.synthetic import make_batch  # noqa: E402

STAGES = ["gemm", "frontend", "gcn_fwd", "gcn_bwd", "score", "loss", "adam", "prep"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--edge-feature", default="scaler")
    ap.add_argument("--option", default="vec_ctas_per_sm")
    ap.add_argument("--values", default="0,2,3,4,5,6,8")
    ap.add_argument("--reset", type=int, default=0, help="value that restores the default")
    args = ap.parse_args()
    lib = _lib.load()
    batch = make_batch("wikidiverse", args.batch, seed=1, num_candidates=10, device="cuda", generate_on_device=True)
    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=11, gcn_edge_feature=args.edge_feature, num_gcn_layers=args.layers).cuda()
    tr = drin_b200.Trainer(model)

    def opt(name, v):
        _lib.check(lib.drin_debug_option(name.encode(), C.c_int32(v)), name)

    def run(label):
        for _ in range(3):
            tr.step(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            tr.step(batch)
        e1.record()
        torch.cuda.synchronize()
        lib.drin_profile_enable(1)
        for _ in range(5):
            tr.step(batch)
        torch.cuda.synchronize()
        n = len(STAGES)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        lib.drin_profile_collect(ms, fl, by, cnt)
        lib.drin_profile_enable(0)
        print(f"{label:>26s}: {e0.elapsed_time(e1) / 10:.3f} ms/step  gcn_fwd {ms[2] / 5:.3f}  gcn_bwd {ms[3] / 5:.3f}  "
              f"score {ms[4] / 5:.3f}  gemm {ms[0] / 5:.3f}", flush=True)

    for v in [int(x) for x in args.values.split(",")]:
        opt(args.option, v)
        run(f"{args.option}={v}")
    opt(args.option, args.reset)


if __name__ == "__main__":
    main()
