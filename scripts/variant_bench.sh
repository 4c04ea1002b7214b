#!/bin/bash
# dev helper: stage times of the train step with a forced kernel variant:  variant_bench.sh <option> <value> [steps]
python - "$@" <<'PY'
import ctypes as C, json, sys, torch
sys.path.insert(0, ".")
import drin_b200
from drin_b200 import _lib
from drin_b200.synthetic import make_batch
lib = _lib.load()
opts = sys.argv[1:]
for i in range(0, len(opts) - 1, 2):
    _lib.check(lib.drin_debug_option(opts[i].encode(), C.c_int32(int(opts[i + 1]))), "opt")
B = 4096
batch = make_batch("wikidiverse", B, 1000, 10, device="cuda", generate_on_device=True)
torch.manual_seed(0)
tr = drin_b200.Trainer(drin_b200.Model(num_candidates_model=11).cuda())
for _ in range(3): tr.step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): tr.step(batch)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
names = ["gemm", "frontend", "gcn_fwd", "gcn_bwd", "score", "loss", "adam", "prep"]
lib.drin_profile_enable(1)
tr.step(batch); torch.cuda.synchronize()
n = 8
a, b, c, d = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
lib.drin_profile_collect(a, b, c, d)
for _ in range(5): tr.step(batch)
torch.cuda.synchronize()
lib.drin_profile_collect(a, b, c, d)
lib.drin_profile_enable(0)
with torch.no_grad():
    for _ in range(3): tr.rank_scores(batch)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): tr.rank_scores(batch)
    e1.record(); torch.cuda.synchronize()
print(opts, "train ms", round(ms, 3), "rank ms", round(e0.elapsed_time(e1) / 10, 3), {k: round(a[i] / 5, 3) for i, k in enumerate(names)})
PY
