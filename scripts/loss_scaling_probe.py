"""How does the sharded TripletLoss scale with the number of ranks?  (rows_local fixed at 4096, B_glob = N x 4096)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drin_b200.loss import triplet_loss_sharded  # noqa: E402

torch.manual_seed(0)
for world in (1, 2, 4, 8):
    B, C = 4096 * world, 11
    s = (torch.rand(B, C, device="cuda") * 0.4 + 0.5)
    y = torch.eye(C - 1, dtype=torch.uint8, device="cuda")[torch.randint(0, C - 1, (B,), device="cuda")]
    for _ in range(3):
        triplet_loss_sharded(s, y, 0.25, 4096 * (world - 1), 4096)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        triplet_loss_sharded(s, y, 0.25, 4096 * (world - 1), 4096)
    e1.record()
    torch.cuda.synchronize()
    print(f"world {world}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per loss call (B_glob = {B})", flush=True)
