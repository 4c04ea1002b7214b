"""Probe: does splitting a step into micro-batches on separate streams overlap the HBM-bound row kernels / front end
of one micro-batch with the tensor-bound GEMMs of another?  Python-level prototype (two engines, two streams)."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drin_b200  # noqa: E402
from drin_b200 import _lib, engine as E  # noqa: E402
from drin_b200.loss import triplet_loss_sharded  # noqa: E402
from drin_b200.synthetic import make_batch  # noqa: E402


def option(name, v):
    _lib.check(_lib.load().drin_debug_option(name.encode(), C.c_int32(v)), name)


def timed(fn, steps=12, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def probe(dataset, B, chunk_list, caps, out):
    cands = 10 if dataset == "wikidiverse" else 100
    Cn = cands + 1
    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=Cn).cuda()
    opt = drin_b200.FusedAdam(model)
    batch = make_batch(dataset, B, seed=1, num_candidates=cands, device="cuda", generate_on_device=True)
    params = model._param_views()
    for chunks in chunk_list:
        Bc = B // chunks
        parts = [[t[i * Bc:(i + 1) * Bc] for t in batch] for i in range(chunks)]
        engines = [E.Engine(2) for _ in range(chunks)]
        streams = [torch.cuda.Stream() for _ in range(chunks)]
        gbufs = [torch.zeros_like(model.flat_params) for _ in range(chunks)]

        def fork():
            ev = torch.cuda.Event()
            ev.record()
            for s in streams:
                s.wait_event(ev)

        def join():
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)

        def rank_step():
            fork()
            for i in range(chunks):
                with torch.cuda.stream(streams[i]):
                    engines[i].forward(tuple(parts[i][:14]), params, False)
            join()

        def train_step():
            fork()
            outs = []
            for i in range(chunks):
                with torch.cuda.stream(streams[i]):
                    outs.append(engines[i].forward(tuple(parts[i][:14]), params, True))
            join()
            scores = torch.cat([o[0] for o in outs]) if chunks > 1 else outs[0][0]
            loss, ds = triplet_loss_sharded(scores, batch[-1], 0.25)
            fork()
            for i in range(chunks):
                with torch.cuda.stream(streams[i]):
                    engines[i].backward(outs[i][1], tuple(parts[i][:14]), params, ds[i * Bc:(i + 1) * Bc],
                                        model._grad_views(gbufs[i]))
                    outs[i][1].release()
            join()
            if chunks > 1:
                torch.add(gbufs[0], gbufs[1], out=model.flat_grads)
                for gb in gbufs[2:]:
                    model.flat_grads.add_(gb)
            else:
                model.flat_grads.copy_(gbufs[0])
            model._flat_grads_valid = True
            opt.step()

        for cap in caps:
            option("gemm_sm_cap", cap)
            r = dict(dataset=dataset, B=B, chunks=chunks, gemm_sm_cap=cap, rank_ms=timed(rank_step), train_ms=timed(train_step))
            print(json.dumps(r), flush=True)
            out.append(r)
        option("gemm_sm_cap", 0)
        del engines, gbufs, parts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    out = []
    probe("wikidiverse", 4096, [1, 2, 4], [0, 132, 116], out)
    probe("wikimel", 576, [1, 2, 3, 4], [0, 132, 116], out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pipeline_probe.json"), "w"), indent=1)
