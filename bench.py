#!/usr/bin/env python
"""Benchmark of the DRIN hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--dataset wikidiverse|wikimel] [--batch B_per_gpu] [--precision fp32|bf16]
                    [--edge-feature scaler|vector]

One "step" = one pass of the hot path over one batch of synthetic features:
  train step = forward + TripletLoss + backward (+ gradient all-reduce at N > 1) + Adam   (headline)
  ranking    = forward under no_grad                                                      (reported beside it)
Workload at N = 1 is BASELINE.json configs[1]: DRIN training on WikiDiverse-shaped features
(10 candidates + gold slot per mention), fp32-parity mode.  Weak scaling: the per-GPU batch is fixed.

`value`  : mentions/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through the public API with HOST (pinned) buffers: every step copies its 15-tensor
           batch host->device and reads the loss back, inside the timed region.
`--impl reference` times the CPU port of the reference (oracle/, loop-faithful mode) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mentions/sec train step (fwd+loss+bwd+Adam); ranking reported beside it"
UNIT = "mentions/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dataset", default="wikidiverse", choices=["wikidiverse", "wikimel"])
    ap.add_argument("--batch", type=int, default=0, help="mentions per GPU per step (0: 4096 WikiDiverse, 512 WikiMEL)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--cpu-batch", type=int, default=512,
                    help="mentions per step of the CPU reference sample (larger batches favour the CPU: 553 m/s at 32, 872 at 512)")
    ap.add_argument("--edge-feature", default="scaler", choices=["scaler", "vector"],
                    help="gcn_edge_feature (args.py:33); the headline is the reference default, scalar edges")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# algorithmic work per mention
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_mention(C, wm, D=768, R=2048, P=49, Om=3, Oe=1, Le=64, span=3.0, nbar=34.0, feat_bytes=4):
    te = (nbar - 1) if wm else 1
    b = feat_bytes * (span * D + P * R + Om * R) + 4 * Om + 16
    b += C * (feat_bytes * (te * D + R + Oe * R) + 4 * (Oe + 2)) + (C - 1)
    if wm:
        b += 8 * C * Le
    return b


def gemm_flops_per_mention(C, train, D=768, R=2048, vector=False):
    u, r = 2 * D * D, 2 * R * D
    if vector:   # as the reference formulates it, per unit: projections u + r; layer 1: W_h 2u, W_u | W_v (D -> D/2) on both
        # kinds u, W_m on the 4 edge types 4u; layer 2: W_h on the text kind u
        return (1 + C) * ((2 * r + 26 * u) if train else (r + 9 * u))
    return (1 + C) * ((2 * r + 17 * u) if train else (r + 6 * u))     # reference-necessary


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6 and parts[0].isdigit():
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [int(s[0]) for s in self.samples]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": int(statistics.median(sm)), "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port, loop-faithful like the reference's Python loops)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(dataset, batch, steps, warmup, loops=True, edge_feature="scaler", kind="auto"):
    """Time the train step and the ranking forward of the reference's CPU path on all host cores.

    kind "reference": the UNMODIFIED reference (drin/model.py, common/utils.py TripletLoss, torch.optim.Adam -- the body of
    train.py:32-34,55-56) imported from oracle/_ref (staged by oracle/make_ref.py) or /root/reference;
    kind "port": the oracle restatement (loops=True keeps the reference's per-item Python loops).
    "auto" takes the reference when it is importable."""
    import torch
    from drin_b200.synthetic import make_batch
    from oracle import drin_oracle as O
    from oracle import ref_import

    torch.set_num_threads(os.cpu_count() or 1)
    cands = 10 if dataset == "wikidiverse" else 100
    b = make_batch(dataset, batch, 0, cands)
    if kind == "auto":
        kind = "reference" if ref_import.available() else "port"
    if kind == "reference":
        m, u, a = ref_import.load(dataset, cands, 64, gcn_edge_feature=edge_feature)
        torch.manual_seed(0)
        model = m.Model()
        loss_obj = u.TripletLoss(a.triplet_margin)
        opt = torch.optim.Adam(model.parameters(), lr=a.learning_rate)

        def step():
            y_hat = model(b[:-1])                       # train.py:32-34
            loss = loss_obj(b[-1], y_hat)
            opt.zero_grad(set_to_none=True)             # Lightning's automatic optimisation
            loss.backward()
            opt.step()
            return float(loss)                          # train.py:35

        def rank():
            with torch.no_grad():
                return model(b[:-1])
    else:
        cfg = O.DrinConfig(num_candidates_model=cands + 1, gcn_edge_feature=edge_feature)
        sd = O.init_state(cfg, 0)
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(leaves.values()), lr=1e-3)
        loss_fn = O.triplet_loss_loops if loops else O.triplet_loss

        def step():
            opt.zero_grad(set_to_none=True)
            s = O.forward(leaves, b[:-1], cfg, loops)
            loss = loss_fn(b[-1], s, cfg.triplet_margin)
            loss.backward()
            opt.step()
            return float(loss)

        def rank():
            with torch.no_grad():
                return O.forward(leaves, b[:-1], cfg, loops)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    t_train = (time.perf_counter() - t0) / steps
    rank()
    t0 = time.perf_counter()
    for _ in range(max(steps // 2, 1)):
        rank()
    t_rank = (time.perf_counter() - t0) / max(steps // 2, 1)
    return dict(train_mps=batch / t_train, rank_mps=batch / t_rank, ms_per_step=t_train * 1e3, cores=os.cpu_count(),
                threads=torch.get_num_threads(), kind=kind)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch if args.cpu_batch else 512
    r = cpu_reference_run(args.dataset, B, args.steps, args.warmup, loops=True, edge_feature=args.edge_feature)
    cands = 10 if args.dataset == "wikidiverse" else 100
    what = ("the UNMODIFIED reference (drin/model.py + common/utils.py TripletLoss + torch Adam, staged under oracle/_ref)"
            if r["kind"] == "reference" else "oracle port of the reference (per-item Python loops kept)")
    sample = (f"{what}, {args.dataset}-shaped batch of {B} mentions per step, fwd+TripletLoss+bwd+Adam, "
              f"{args.warmup} warm-up + {args.steps} timed steps, torch CPU {r['threads']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["train_mps"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DRIN train step, {args.dataset}-shaped synthetic features, C={cands + 1}",
                   "batch_per_step": B, "parallelism": "host CPU"},
        "cpu_baseline": {"value": r["train_mps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["train_mps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ranking": {"value": r["rank_mps"], "unit": UNIT},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import drin_b200
    from drin_b200 import _lib
    from drin_b200.synthetic import batch_bytes, make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lib.drin_launch_count.restype = C.c_longlong

    wm = args.dataset == "wikimel"
    cands = 100 if wm else 10
    Cn = cands + 1
    B = args.batch or (512 if wm else 4096)
    bf16 = args.precision == "bf16"
    feats = (0, 4, 5, 7, 9, 10)

    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=Cn, gcn_edge_feature=args.edge_feature).to(dev)
    trainer = drin_b200.Trainer(model, lr=1e-3, margin=0.25)
    batch = make_batch(args.dataset, B, seed=1000 + rank, num_candidates=cands, device=str(dev), generate_on_device=True)
    if bf16:
        batch = [t.to(torch.bfloat16) if i in feats else t for i, t in enumerate(batch)]
    in_bytes = batch_bytes(batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, after=None):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    stage_names = ["gemm", "frontend", "gcn_fwd", "gcn_bwd", "score", "loss", "adam", "prep"]

    def collect_profile():
        n = len(stage_names)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        _lib.check(lib.drin_profile_collect(ms, fl, by, cnt), "drin_profile_collect")
        return {s: dict(ms=ms[i], flops=fl[i], launches=cnt[i]) for i, s in enumerate(stage_names)}

    # ---------------- device-resident train step (the headline `value`) ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        trainer.step(batch)
    torch.cuda.synchronize()
    launches0 = lib.drin_launch_count()
    ms_train = timed(lambda: trainer.step(batch), args.steps, 0)          # the headline: no per-launch events
    launches = lib.drin_launch_count() - launches0
    # same K steps again with a CUDA-event pair around every launch (on the launching stream): per-stage device time
    lib.drin_profile_enable(1)
    trainer.step(batch)
    torch.cuda.synchronize()
    collect_profile()                                   # drop the first profiled step
    ms_train_profiled = timed(lambda: trainer.step(batch), args.steps, 0)
    prof = collect_profile()
    lib.drin_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms_train * 1e-3)

    # ---------------- ranking (forward under no_grad), device-resident ----------------
    ms_rank = timed(lambda: trainer.rank_scores(batch), args.steps, args.warmup)
    lib.drin_profile_enable(1)
    trainer.rank_scores(batch)
    torch.cuda.synchronize()
    prof_rank = collect_profile()
    lib.drin_profile_enable(0)

    # ---------------- end to end: host (pinned) batch -> device every step, loss read back ----------------
    e2e = None
    if not args.no_e2e:
        host = [t.cpu().pin_memory() for t in batch]
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def run_e2e(compact):
            feeder = drin_b200.HostFeeder(dev, slots=2, compact_spans=compact)
            state = {"slot": feeder.submit(host)}

            def e2e_step():
                # double buffered: the host gather + copy of step i+1 overlap the compute of step i; all of it is
                # inside the timed region, and the loss is read back every step (reference train.py:35)
                sid = state["slot"]
                dbatch = feeder.get(sid)
                loss = trainer.step(dbatch)
                feeder.release(sid)
                state["slot"] = feeder.submit(host)
                loss_host.copy_(loss.reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return float(loss_host)

            # K timed iterations must contain exactly K host->device copies: the copy of the first timed step was
            # prefetched during warm-up, so the region also waits for the copy submitted by its last iteration
            def drain():
                torch.cuda.current_stream().wait_event(feeder.slots[state["slot"]]["ready"])

            ms = timed(e2e_step, max(args.steps // 2, 3), 3, after=drain)
            nbytes = feeder.last_bytes
            del feeder
            return ms, nbytes

        ms_full, bytes_full = run_e2e(False)
        ms_e2e, bytes_e2e = run_e2e(True)
        e2e = {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bytes_e2e,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
               "note": ("pinned host batch -> device every step on a side stream (double buffered), loss read back; "
                        "HostFeeder copies only the bytes the path reads: the start:end span rows of "
                        "mention_text_feature are gathered on the host (the other rows are never read)"),
               "full_copy": {"value": world * B / (ms_full * 1e-3), "h2d_bytes_per_step": bytes_full,
                             "ms_per_step": ms_full, "note": "all 15 tensors copied verbatim"}}
        del host

        # ---------------- resident feature store: tables in HBM, only indices cross PCIe -------
        from drin_b200.store import FeatureStore, synthetic_tables
        n_store = 3 * B
        tables = synthetic_tables(args.dataset, n_store, seed=2000 + rank, num_candidates=cands, device=str(dev))
        store = FeatureStore(args.dataset, tables, Cn, device=dev,
                             feature_dtype=torch.bfloat16 if bf16 else torch.float32)
        del tables
        order = torch.randperm(n_store, generator=torch.Generator().manual_seed(rank)).pin_memory()
        cursor = {"k": 0}

        def store_step():
            k = cursor["k"]
            cursor["k"] = (k + 1) % 3
            loss = trainer.step(store.select(order[k * B:(k + 1) * B]))      # host indices -> device, 8 B / mention
            loss_host.copy_(loss.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(loss_host)

        ms_store = timed(store_step, max(args.steps // 2, 3), 3)
        e2e["resident_store"] = {
            "value": world * B / (ms_store * 1e-3), "unit": UNIT, "ms_per_step": ms_store,
            "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 4, "resident_bytes_per_gpu": store.nbytes(),
            "mentions_resident_per_gpu": n_store,
            "note": ("FeatureStore: the split's feature tables are uploaded once and stay in HBM; each step ships the "
                     "[B] mention indices from pinned host memory, the front-end kernel gathers rows by index "
                     "(drin/data.py:85-108 on device), loss read back every step")}
        del store

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel family (tcgen05 GEMM), live CUDA-event times ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json, sustained bf16)" if peaks else "fallback"
    g = prof["gemm"]
    gemm_tflops = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    passes = 1 if bf16 else 3
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json"))).get("traffic_bytes_per_launch")
    except Exception:
        pass
    total_ms = sum(v["ms"] for v in prof.values())
    roofline = {
        "bound": "tensor", "kernel": "gemm_tcgen05_kernel (all GEMM launches of the step)",
        "achieved": gemm_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": gemm_tflops / tensor_peak,
        "traffic": traffic, "peak_source": peak_src,
        "executed_tflops": gemm_tflops * passes,
        "note": ("achieved = algorithmic 2MNK flops of the launched GEMMs / their summed CUDA-event time; fp32-parity mode "
                 "issues 3 bf16 tensor passes per algorithmic flop (split-bf16), so executed = 3x achieved"),
        "gemm_launches_per_step": g["launches"] / args.steps, "gemm_share_of_step": g["ms"] / total_ms if total_ms else None,
    }
    fe = prof["frontend"]
    fe_bytes = algorithmic_bytes_per_mention(Cn, wm, feat_bytes=2 if bf16 else 4) * B * args.steps
    fe_gbs = fe_bytes / (fe["ms"] * 1e-3) / 1e9 if fe["ms"] > 0 else 0.0
    roofline_hbm = {"bound": "hbm", "kernel": "frontend_kernel", "achieved": fe_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": fe_gbs / hbm_peak, "traffic": None,
                    "note": "algorithmic input bytes (SURVEY 8d) / CUDA-event time of the front-end kernel"}
    rk_fl = prof_rank["gemm"]
    ranking = {"value": world * B / (ms_rank * 1e-3), "unit": UNIT, "ms_per_step": ms_rank,
               "gemm_tflops": rk_fl["flops"] / (rk_fl["ms"] * 1e-3) / 1e12 if rk_fl["ms"] > 0 else None,
               "frontend_gbs": (algorithmic_bytes_per_mention(Cn, wm, feat_bytes=2 if bf16 else 4) * B /
                                (prof_rank["frontend"]["ms"] * 1e-3) / 1e9) if prof_rank["frontend"]["ms"] > 0 else None}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.dataset, args.cpu_batch, 12, 2, loops=True, edge_feature=args.edge_feature)
        rv = cpu_reference_run(args.dataset, args.cpu_batch, 6, 1, loops=False, edge_feature=args.edge_feature, kind="port")
        cpu = {"value": r["train_mps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": (("the unmodified reference (oracle/_ref)" if r["kind"] == "reference" else
                           "oracle port (reference-style Python loops)") +
                          f", {args.dataset}-shaped batch of {args.cpu_batch}, "
                          f"fwd+loss+bwd+Adam, 2 warm-up + 12 timed steps, {r['threads']} torch threads"),
               "ranking_value": r["rank_mps"], "vectorised_port_value": rv["train_mps"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_train, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if bf16 else "f32 (GEMMs as 3-pass split-bf16 on tcgen05, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"DRIN training step, {args.dataset}-shaped synthetic features, C={Cn} candidate slots",
                   "gcn_edge_feature": args.edge_feature,
                   "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
                   "l2": f"inputs {in_bytes / 2**20:.0f} MiB per step per GPU, larger than the 126 MB L2 (no flush needed)",
                   "input_bytes_per_mention": in_bytes / B,
                   "loss_semantics": "global-batch TripletLoss (scores all-gathered), gradients summed"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "ranking": ranking,
        "stage_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        "ms_per_step_with_stage_events": ms_train_profiled,
        "necessary_gemm_tflops_of_step": gemm_flops_per_mention(Cn, True, vector=args.edge_feature == "vector") * B / (ms_train * 1e-3) / 1e12,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "RANK" not in os.environ:
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
