"""torchrun-launched check of the NCCL data-parallel path: N ranks, each a shard of one global batch, must
reproduce the single-GPU full-batch step (loss, all-reduced gradients, parameters after Adam) and ranking."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200.synthetic import make_batch, spread_weights  # noqa: E402
from oracle import drin_oracle as O  # noqa: E402  (weights init only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # DRIN_DP_SAME_DEVICE=1: every rank on cuda:0 over gloo -- the same data-parallel protocol through the same CUDA
    # kernels on a box with ONE GPU (NCCL cannot put two ranks on one device)
    same_device = os.environ.get("DRIN_DP_SAME_DEVICE", "0") == "1"
    if same_device:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    dataset, cands = ("wikidiverse", 10)
    edge_feature = sys.argv[1] if len(sys.argv) > 1 else "scaler"
    B = 16 * world
    batch = make_batch(dataset, B, 21, cands)
    cfg = O.DrinConfig(num_candidates_model=cands + 1, triplet_margin=0.05, gcn_edge_feature=edge_feature)
    sd = spread_weights(O.init_state(cfg, 0))

    def model():
        m = drin_b200.Model(num_candidates_model=cands + 1, gcn_edge_feature=edge_feature)
        m.load_state_dict(sd)
        return m.to(dev)

    bl = B // world
    shard = [t[rank * bl:(rank + 1) * bl].contiguous().to(dev) for t in batch]
    tr = drin_b200.Trainer(model(), lr=1e-3, margin=cfg.triplet_margin)
    loss = tr.forward_backward(shard)
    g_dp = tr.model.flat_grads[:tr.model.n_live].clone()      # live gradients (the tail slot carries the summed loss)
    tr.opt.step()
    scores_all = tr.rank_scores(shard, gather=True)
    res = {}
    if rank == 0:
        full = [t.to(dev) for t in batch]
        ref = drin_b200.Trainer(model(), lr=1e-3, margin=cfg.triplet_margin, group=dist.new_group([0]) if False else None)
        # single-rank reference: bypass the process group by calling the engine pieces directly
        m = ref.model
        params = m._param_views()
        s, ctx = m._engine.forward(tuple(full[:-1]), params, training=True)
        from drin_b200.loss import triplet_loss_sharded
        l_ref, ds = triplet_loss_sharded(s, full[-1], cfg.triplet_margin)
        m._engine.backward(ctx, tuple(full[:-1]), params, ds, m._grad_views())
        g_ref = m.flat_grads[:m.n_live]
        m._flat_grads_valid = True                 # the engine pieces were called directly (what Trainer does)
        res["loss_err"] = abs(float(loss) - float(l_ref)) / abs(float(l_ref))
        res["grad_err"] = float((g_dp - g_ref).abs().max() / g_ref.abs().max())
        ref.opt.step()
        res["param_err_after_adam"] = float((tr.model.flat_params - m.flat_params).abs().max())
        s2, _ = m._engine.forward(tuple(full[:-1]), m._param_views(), training=False)
        res["rank_scores_err"] = float((scores_all - s2).abs().max())
        res["world"] = world
        print("DPCHECK " + json.dumps(res), flush=True)
    else:
        # other ranks must not hang: nothing collective happens in the rank-0-only reference above
        pass
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
