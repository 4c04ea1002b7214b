"""torchrun-launched check of ``drin_b200.fit`` under data parallelism: N ranks (NCCL, or gloo on one shared GPU with
DRIN_DP_SAME_DEVICE=1) run fit() with a GLOBAL batch size; the weights after training must equal a single-process fit()
of the same schedule (same seed, same global batches) up to fp32 re-association of the summed gradients."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200.store import FeatureStore, synthetic_tables  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    same_device = os.environ.get("DRIN_DP_SAME_DEVICE", "0") == "1"
    if same_device:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo" if same_device else "nccl", **({} if same_device else {"device_id": dev}))
    cands = 10

    def stores():
        out = []
        for n, seed in ((72, 1), (16, 2), (16, 3)):            # 72 = 4 x 16 + 8: a short last batch
            t = synthetic_tables("wikidiverse", n, seed=seed, num_candidates=cands, device=str(dev))
            out.append(FeatureStore("wikidiverse", t, cands + 1, device=dev))
        return out

    def model():
        torch.manual_seed(0)
        m = drin_b200.Model(num_candidates_model=cands + 1).to(dev)
        with torch.no_grad():
            for l in m.gcn_layers:
                l.w_h.weight.mul_(3.0)
        return m

    solo = [dist.new_group([r]) for r in range(world)]                 # collective: every rank creates every group
    kw = dict(batch_size=16, num_epoch=2, test_epoch_interval=1, seed=3, log=None)
    m_dp = model()
    hist_dp = drin_b200.fit(m_dp, *stores(), **kw)                      # every rank: rows rank::world of each global batch
    res = {}
    if rank == 0:
        m_one = model()
        # single-process reference: a group that contains only this rank short-circuits every collective in fit()
        hist_one = drin_b200.fit(m_one, *stores(), group=solo[0], **kw)
        res["param_err"] = float((m_dp.flat_params - m_one.flat_params).abs().max())
        res["loss_err"] = max(abs(a["loss"] - b["loss"]) / max(abs(b["loss"]), 1e-9)
                              for a, b in zip(hist_dp, hist_one) if a["type"] == b["type"] == "training")
        res["records"] = [r["type"] for r in hist_dp]
        res["topk_err"] = max(abs(a["topk"][k] - b["topk"][k]) for a, b in zip(hist_dp, hist_one) for k in a["topk"])
        res["world"] = world
        print("FITCHECK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
