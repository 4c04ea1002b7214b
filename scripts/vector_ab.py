"""A/B of the vector-edge kernels' launch options on one B200 (test hooks of drin_debug_option):
    python scripts/vector_ab.py [--batch 4096]
prints ms/step and the GCN forward / backward stage times for vec_ctas_per_sm in {auto, 2..8} and vec_bwd_width 2."""
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
    args = ap.parse_args()
    lib = _lib.load()
    batch = make_batch("wikidiverse", args.batch, seed=1, num_candidates=10, device="cuda", generate_on_device=True)
    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=11, gcn_edge_feature="vector", num_gcn_layers=args.layers).cuda()
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
        print(f"{label:>22s}: {e0.elapsed_time(e1) / 10:.3f} ms/step  gcn_fwd {ms[2] / 5:.3f}  gcn_bwd {ms[3] / 5:.3f}  "
              f"gemm {ms[0] / 5:.3f}", flush=True)

    run("auto")
    for k in (2, 3, 4, 5, 6, 8):
        opt("vec_ctas_per_sm", k)
        run(f"ctas_per_sm<={k}")
    opt("vec_ctas_per_sm", 0)
    opt("vec_bwd_width", 2)
    run("bwd 2 cols/thread")
    opt("vec_bwd_width", 0)


if __name__ == "__main__":
    main()
