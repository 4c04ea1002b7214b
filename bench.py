#!/usr/bin/env python
"""Benchmark of the DRIN hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--dataset wikidiverse|wikimel] [--batch B_per_gpu] [--precision fp32|bf16]
                    [--edge-feature scaler|vector] [--no-legs] [--no-e2e] [--no-cpu-baseline]

One "step" = one pass of the hot path over one batch of synthetic features:
  train step = forward + TripletLoss + backward (+ gradient all-reduce at N > 1) + Adam   (headline)
  ranking    = forward under no_grad                                                      (reported beside it)
Headline workload = BASELINE.json configs[1]: DRIN training on WikiDiverse-shaped features (10 candidates + gold slot
per mention), fp32-parity mode, 4096 mentions per GPU per step.  Weak scaling: the per-GPU batch is fixed.

`value`  : mentions/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through the public API with HOST (pinned) buffers: every step copies its 15-tensor batch
           host->device and reads the loss back, inside the timed region; the measured host->device ceiling of the box
           (all ranks copying at once) is printed next to it.
`legs`   : the other BASELINE.json configurations, measured in the same run at the same N (compact):
           configs[2] WikiMEL-shaped ranking over 100 candidates + gold slot, mentions sharded over the ranks;
           configs[3] bf16-feature data-parallel training; configs[4] token / candidate sweep (N = 1 only).
`--impl reference` times the reference's own CPU implementation on the host cores: the UNMODIFIED reference staged under
oracle/_ref (oracle/make_ref.py) when present (kind "reference"), else the oracle port (kind "port").
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
FEATS = (0, 4, 5, 7, 9, 10)          # the feature tensors of the 14-tensor batch (bf16 in bf16 mode)
STAGES = ["gemm", "frontend", "gcn_fwd", "gcn_bwd", "score", "loss", "adam", "prep"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dataset", default="wikidiverse", choices=["wikidiverse", "wikimel"])
    ap.add_argument("--batch", type=int, default=0, help="mentions per GPU per step (0: 4096 WikiDiverse, 576 WikiMEL)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="mentions per step of the CPU reference sample (0: 512 -- the batch at which the reference is "
                         "fastest on the host: 2153 m/s at 512 against 1378 at our per-GPU batch of 4096 and ~550 at 32)")
    ap.add_argument("--edge-feature", default="scaler", choices=["scaler", "vector"],
                    help="gcn_edge_feature (args.py:33); the headline is the reference default, scalar edges")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the configs[2]/[3]/[4] legs")
    return ap.parse_args()


def default_batch(dataset: str) -> int:
    return 576 if dataset == "wikimel" else 4096


def batch_bytes_per_mention(dataset, bf16=False, D=768, R=2048, P=49, Om=3, Lm=128, Le=64) -> float:
    """Bytes of the loader's full 15-tensor layout per mention (what a verbatim copy moves)."""
    wm = dataset == "wikimel"
    C = 101 if wm else 11
    f = 2 if bf16 else 4
    b = f * (Lm * D + P * R + Om * R) + 8 * Lm + 16 + 4 * Om
    b += C * (f * ((Le if wm else 1) * D + 2 * R) + 4 * 3) + (C - 1)
    b += 8 * C * Le if wm else 8
    return float(b)


def workload_config(args, world: int, B: int) -> dict:
    """`config` of the JSON line -- the same dict for our arm and for the reference arm (the reference arm runs a bounded
    sample of it on the host, described in its cpu_baseline.sample)."""
    cands = 100 if args.dataset == "wikimel" else 10
    per = batch_bytes_per_mention(args.dataset, args.precision == "bf16")
    return {"workload": f"DRIN training step, {args.dataset}-shaped synthetic features, C={cands + 1} candidate slots",
            "gcn_edge_feature": args.edge_feature, "batch_per_gpu": B, "global_batch": world * B,
            "parallelism": f"dp{world}",
            "l2": f"inputs {per * B / 2**20:.0f} MiB per step per GPU, larger than the 126 MB L2 (no flush needed)",
            "input_bytes_per_mention": per,
            "loss_semantics": "global-batch TripletLoss (scores all-gathered), gradients summed"}


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

    def __init__(self, index: int, enabled: bool = True):
        self.samples, self.proc, self.index, self.enabled = [], None, index, enabled

    def start(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            # nvidia-smi needs a few hundred ms to come up: wait for its first line so that the samples cover the timed
            # region that follows (a 150 ms region would otherwise see one sample or none), then drop the idle samples
            t0 = time.perf_counter()
            while not self.samples and time.perf_counter() - t0 < 2.0:
                time.sleep(0.01)
            self.samples.clear()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6 and parts[0].isdigit():
                self.samples.append(parts)

    def stop(self):
        if not self.enabled:
            return None
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
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(dataset, batch, steps, warmup, loops=True, edge_feature="scaler", kind="auto", budget_s=None):
    """Time the train step and the ranking forward of the reference's CPU path on all host cores.

    kind "reference": the UNMODIFIED reference (drin/model.py, common/utils.py TripletLoss, torch.optim.Adam -- the body of
    train.py:32-34,55-56) imported from oracle/_ref (staged by oracle/make_ref.py) or /root/reference;
    kind "port": the oracle restatement (loops=True keeps the reference's per-item Python loops).
    "auto" takes the reference when it is importable.  budget_s: stop the timed loop early (>= 3 steps) once it is spent."""
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
            return float(loss.detach())                 # train.py:35

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
            return float(loss.detach())

        def rank():
            with torch.no_grad():
                return O.forward(leaves, b[:-1], cfg, loops)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if budget_s is not None and done >= 3 and time.perf_counter() - t0 > budget_s:
            break
    t_train = (time.perf_counter() - t0) / done
    rank()
    nr = max(done // 2, 1)
    t0 = time.perf_counter()
    for _ in range(nr):
        rank()
    t_rank = (time.perf_counter() - t0) / nr
    return dict(train_mps=batch / t_train, rank_mps=batch / t_rank, ms_per_step=t_train * 1e3, cores=os.cpu_count(),
                threads=torch.get_num_threads(), kind=kind, steps_timed=done)


def _cpu_what(kind):
    return ("the UNMODIFIED reference (drin/model.py + common/utils.py TripletLoss + torch.optim.Adam, staged under "
            "oracle/_ref)" if kind == "reference" else "oracle port of the reference (per-item Python loops kept)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch or min(512, args.batch or default_batch(args.dataset))
    r = cpu_reference_run(args.dataset, B, args.steps, args.warmup, loops=True, edge_feature=args.edge_feature,
                          budget_s=150.0)
    sample = (f"{_cpu_what(r['kind'])}: a bounded sample of the workload in `config` -- one {args.dataset}-shaped batch of "
              f"{B} mentions per step (the batch size at which the host is fastest; per-mention throughput is the metric), "
              f"fwd+TripletLoss+bwd+Adam, {args.warmup} warm-up + {r['steps_timed']} timed steps, torch CPU "
              f"{r['threads']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["train_mps"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus, args.batch or default_batch(args.dataset)),
        "cpu_baseline": {"value": r["train_mps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["train_mps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ranking": {"value": r["rank_mps"], "unit": UNIT},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from drin_b200 import _lib

        self.args, self.C, self.torch, self.dist, self._lib = args, C, torch, dist, _lib
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _lib.load()
        self.lib.drin_launch_count.restype = C.c_longlong
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # fallback = the profiling recipe's figures (B200_PROFILING.md) when the driver-written file is absent
        self.tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        self.hbm_peak = peaks.get("hbm_gbs", 6650.0)
        self.peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    # ---- helpers ----
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, after=None):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    def collect_profile(self):
        C, n = self.C, len(STAGES)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        self._lib.check(self.lib.drin_profile_collect(ms, fl, by, cnt), "drin_profile_collect")
        return {s: dict(ms=ms[i], flops=fl[i], launches=cnt[i]) for i, s in enumerate(STAGES)}

    def profiled(self, fn, steps):
        """Per-stage device time: CUDA-event pairs around every launch of the library, on the launching stream."""
        self.lib.drin_profile_enable(1)
        fn()
        self.torch.cuda.synchronize()
        self.collect_profile()                      # drop the first profiled step
        ms = self.timed(fn, steps, 0)
        prof = self.collect_profile()
        self.lib.drin_profile_enable(0)
        return ms, prof

    def make(self, dataset, B, bf16=False, edge_feature="scaler", seed=1000, **kw):
        import drin_b200
        from drin_b200.synthetic import make_batch
        torch = self.torch
        cands = kw.pop("cands", 100 if dataset == "wikimel" else 10)
        torch.manual_seed(0)
        model = drin_b200.Model(num_candidates_model=cands + 1, gcn_edge_feature=edge_feature).to(self.dev)
        trainer = drin_b200.Trainer(model, lr=1e-3, margin=0.25)
        batch = make_batch(dataset, B, seed=seed + self.rank, num_candidates=cands, device=str(self.dev),
                           generate_on_device=True, **kw)
        if bf16:
            batch = [t.to(torch.bfloat16) if i in FEATS else t for i, t in enumerate(batch)]
        return model, trainer, batch

    def h2d_ceiling(self, nbytes=1 << 30, reps=4):
        """Host->device bandwidth of one large pinned copy per rank, all ranks copying at the same time: the ceiling the
        e2e leg runs against on this box (PCIe link + host fabric shared by the GPUs)."""
        torch = self.torch
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        d = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        ms = self.timed(lambda: d.copy_(h, non_blocking=True), reps, 1)
        del h, d
        return self.world * nbytes / (ms * 1e-3) / 1e9

    # ---- legs ----
    def leg_wikimel_ranking(self, steps, warmup):
        """BASELINE configs[2]: ranking over 100 candidates + gold slot (WikiMEL top100 shape), mentions sharded over the
        ranks, no communication but the final gather of the scores (inside the timed region at N > 1)."""
        B = 576
        model, tr, batch = self.make("wikimel", B, seed=3000)
        sampler = ClockSampler(self.local, self.rank == 0).start()
        gather = self.world > 1
        ms = self.timed(lambda: tr.rank_scores(batch, gather=gather), steps, warmup)
        n = max(steps // 2, 3)
        _, prof = self.profiled(lambda: tr.rank_scores(batch, gather=gather), n)
        clocks = sampler.stop()
        abytes = algorithmic_bytes_per_mention(101, True)
        fe_gbs = abytes * B * n / (prof["frontend"]["ms"] * 1e-3) / 1e9
        g = prof["gemm"]
        out = {"config": "BASELINE configs[2]: DRIN ranking inference, WikiMEL top100 shape (C=101, Le=64), fp32-parity, "
                         f"{B} mentions per GPU per pass, mentions sharded over {self.world} GPU(s), final score gather "
                         + ("included" if gather else "n/a at N=1"),
               "value": self.world * B / (ms * 1e-3), "unit": UNIT, "ms_per_pass": ms, "batch_per_gpu": B,
               "path_hbm_gbs_per_gpu": abytes * B / (ms * 1e-3) / 1e9,
               "path_hbm_frac": abytes * B / (ms * 1e-3) / 1e9 / self.hbm_peak,
               "frontend_gbs": fe_gbs, "frontend_hbm_frac": fe_gbs / self.hbm_peak,
               "gemm_tflops": g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else None,
               "stage_ms_per_pass": {k: v["ms"] / n for k, v in prof.items() if v["ms"] > 0},
               "algorithmic_bytes_per_mention": abytes, "clocks": clocks,
               "l2": f"inputs {batch_bytes_per_mention('wikimel') * B / 2**30:.1f} GiB per pass per GPU, larger than the 126 MB "
                     "L2 (no flush needed)", "steps": steps, "warmup": warmup}
        del model, tr, batch
        self.torch.cuda.empty_cache()
        return out

    def leg_bf16_train(self, steps, warmup, dataset="wikidiverse"):
        """BASELINE configs[3]: data-parallel training with bf16 features (single-pass bf16 GEMMs), NCCL all-reduce at
        N > 1; WikiDiverse shape with 4096 mentions per GPU, WikiMEL shape (C=101, Le=64) with 576."""
        wm = dataset == "wikimel"
        B = 576 if wm else 4096
        model, tr, batch = self.make(dataset, B, bf16=True, seed=4000)
        sampler = ClockSampler(self.local, self.rank == 0).start()
        ms = self.timed(lambda: tr.step(batch), steps, warmup)
        n = max(steps // 2, 3)
        _, prof = self.profiled(lambda: tr.step(batch), n)
        clocks = sampler.stop()
        g = prof["gemm"]
        tf = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        out = {"config": "BASELINE configs[3]: DRIN data-parallel training, bf16 features, "
                         + ("WikiMEL shape (C=101, Le=64), " if wm else "WikiDiverse shape (C=11), ") +
                         f"{B} mentions per GPU per step, dp{self.world}" + (", NCCL grad all-reduce" if self.world > 1 else ""),
               "value": self.world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch_per_gpu": B, "dtype": "bf16",
               "gemm_tflops": tf, "gemm_frac_of_tensor_peak": tf / self.tensor_peak,
               "stage_ms_per_step": {k: v["ms"] / n for k, v in prof.items() if v["ms"] > 0},
               "tolerance": "scores/loss 2e-3, gradients 3e-2 against the fp32 reference on bf16-rounded features "
                            "(tests/test_gpu_parity_scale.py)", "clocks": clocks,
               "l2": f"inputs {batch_bytes_per_mention(dataset, True) * B / 2**30:.1f} GiB per step per GPU, larger than "
                     "the 126 MB L2 (no flush needed)", "steps": steps, "warmup": warmup}
        del model, tr, batch
        self.torch.cuda.empty_cache()
        return out

    def leg_sweep(self):
        """BASELINE configs[4] (N = 1): text tokens 32 -> 128 and candidates 10 -> 100; GEMM vs row-kernel stage split."""
        pts = []
        grid = [("wikidiverse", 10, 32, 0), ("wikidiverse", 100, 128, 0), ("wikimel", 10, 128, 64),
                ("wikimel", 100, 32, 32), ("wikimel", 100, 128, 128)]
        for ds, cands, Lm, Le in grid:
            wm = ds == "wikimel"
            per = (Lm * 768 + 49 * 2048 + 3 * 2048) * 4 + (cands + 1) * ((Le if wm else 1) * 768 + 2 * 2048) * 4
            B = max(64, min(4096, int((8 if wm else 3) * 2**30 // per) // 64 * 64))
            kw = dict(mention_tokens=Lm, cands=cands)
            if wm:
                kw["entity_tokens"] = Le
            model, tr, batch = self.make(ds, B, seed=5000, **kw)
            p = dict(dataset=ds, candidates=cands, mention_tokens=Lm, entity_tokens=Le, batch=B)
            nbar = (4 + Le) / 2.0 if wm else 1.0
            abytes = algorithmic_bytes_per_mention(cands + 1, wm, Le=Le, nbar=nbar)
            for mode, fn in (("rank", lambda: tr.rank_scores(batch)), ("train", lambda: tr.step(batch))):
                ms = self.timed(fn, 4, 2)
                _, prof = self.profiled(fn, 3)
                g, fe = prof["gemm"], prof["frontend"]
                rows = sum(prof[k]["ms"] for k in ("gcn_fwd", "gcn_bwd", "score"))
                tf = g["flops"] / (g["ms"] * 1e-3) / 1e12
                gbs = abytes * B * 3 / (fe["ms"] * 1e-3) / 1e9
                p[mode] = dict(mentions_per_s=B / (ms * 1e-3), ms=ms, gemm_tflops=tf,
                               gemm_frac_of_tensor_peak=tf / self.tensor_peak, frontend_gbs=gbs,
                               frontend_hbm_frac=gbs / self.hbm_peak,
                               stage_ms=dict(gemm=g["ms"] / 3, frontend=fe["ms"] / 3, gcn_rows_and_score=rows / 3))
            pts.append(p)
            del model, tr, batch
            self.torch.cuda.empty_cache()
        return {"config": "BASELINE configs[4]: token / candidate sweep, fp32-parity, 1 GPU, device-resident inputs",
                "points": pts}

    # ---- the run ----
    def run(self):
        args, torch, lib, world, rank = self.args, self.torch, self.lib, self.world, self.rank

        wm = args.dataset == "wikimel"
        Cn = (100 if wm else 10) + 1
        B = args.batch or default_batch(args.dataset)
        bf16 = args.precision == "bf16"
        model, trainer, batch = self.make(args.dataset, B, bf16, args.edge_feature)

        # ---------------- device-resident train step (the headline `value`) ----------------
        sampler = ClockSampler(self.local, rank == 0).start()
        for _ in range(args.warmup):
            trainer.step(batch)
        torch.cuda.synchronize()
        launches0 = lib.drin_launch_count()
        ms_train = self.timed(lambda: trainer.step(batch), args.steps, 0)       # the headline: no per-launch events
        launches = lib.drin_launch_count() - launches0
        ms_train_profiled, prof = self.profiled(lambda: trainer.step(batch), args.steps)
        clocks = sampler.stop()
        value = world * B / (ms_train * 1e-3)

        # ---------------- ranking (forward under no_grad), device-resident ----------------
        ms_rank = self.timed(lambda: trainer.rank_scores(batch), args.steps, args.warmup)
        _, prof_rank = self.profiled(lambda: trainer.rank_scores(batch), 3)

        # ---------------- end to end: host (pinned) batch -> device every step, loss read back ----------------
        e2e = None if args.no_e2e else self.e2e(trainer, batch, B, Cn, bf16)
        del batch
        torch.cuda.empty_cache()

        legs = None
        if not args.no_legs and args.dataset == "wikidiverse" and not bf16 and args.edge_feature == "scaler":
            k, w = max(args.steps // 2, 5), max(args.warmup, 3)
            del model, trainer
            torch.cuda.empty_cache()
            legs = {"wikimel_ranking": self.leg_wikimel_ranking(k, w), "bf16_train": self.leg_bf16_train(k, w),
                    "bf16_train_wikimel": self.leg_bf16_train(k, w, "wikimel")}
            if world == 1:
                legs["sweep"] = self.leg_sweep()

        if rank != 0:
            if world > 1:
                self.dist.destroy_process_group()
            return

        # ---------------- roofline of the dominant kernel family (tcgen05 GEMM), live CUDA-event times ----------------
        g = prof["gemm"]
        gemm_tflops = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        passes = 1 if bf16 else 3
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
        except Exception:
            pass
        total_ms = sum(v["ms"] for v in prof.values())
        roofline = {
            "bound": "tensor", "kernel": "gemm_tcgen05_kernel (all GEMM launches of the step)",
            "achieved": gemm_tflops, "peak": self.tensor_peak, "unit": "TFLOP/s", "frac": gemm_tflops / self.tensor_peak,
            "traffic": traffic.get("traffic_bytes_per_launch"), "traffic_kernel": traffic.get("kernel"),
            "traffic_algorithmic_bytes": traffic.get("algorithmic_bytes_per_launch"),
            "peak_source": self.peak_src + ", sustained bf16",
            "executed_tflops": gemm_tflops * passes,
            "executed_frac_of_sustained_peak": gemm_tflops * passes / self.tensor_peak,
            "note": ("achieved = algorithmic 2MNK flops of the launched GEMMs / their summed CUDA-event time; fp32-parity "
                     "mode issues 3 bf16 tensor passes per algorithmic flop (split-bf16), so executed = 3x achieved and the "
                     "ceiling of `frac` is 1/3; `traffic` is the ncu DRAM read+write of ONE launch of the largest GEMM of the "
                     "step (`traffic_kernel`), next to that launch's algorithmic bytes"),
            "gemm_launches_per_step": g["launches"] / args.steps,
            "gemm_share_of_step": g["ms"] / total_ms if total_ms else None,
        }
        fe = prof["frontend"]
        feat_b = 2 if bf16 else 4
        fe_bytes = algorithmic_bytes_per_mention(Cn, wm, feat_bytes=feat_b) * B * args.steps
        fe_gbs = fe_bytes / (fe["ms"] * 1e-3) / 1e9 if fe["ms"] > 0 else 0.0
        roofline_hbm = {"bound": "hbm", "kernel": "frontend_kernel", "achieved": fe_gbs, "peak": self.hbm_peak,
                        "unit": "GB/s", "frac": fe_gbs / self.hbm_peak, "traffic": None,
                        "note": "algorithmic input bytes (SURVEY 8d) / CUDA-event time of the front-end kernel"}
        rk = prof_rank["gemm"]
        ranking = {"value": world * B / (ms_rank * 1e-3), "unit": UNIT, "ms_per_step": ms_rank,
                   "gemm_tflops": rk["flops"] / (rk["ms"] * 1e-3) / 1e12 if rk["ms"] > 0 else None,
                   "frontend_gbs": (algorithmic_bytes_per_mention(Cn, wm, feat_bytes=feat_b) * B * 3 /
                                    (prof_rank["frontend"]["ms"] * 1e-3) / 1e9) if prof_rank["frontend"]["ms"] > 0 else None}

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb = args.cpu_batch or 512
            r = cpu_reference_run(args.dataset, cb, 12, 2, loops=True, edge_feature=args.edge_feature, budget_s=20.0)
            big = cpu_reference_run(args.dataset, B, 3, 1, loops=True, edge_feature=args.edge_feature, budget_s=12.0)
            best = max((r, big), key=lambda x: x["train_mps"])
            cpu = {"value": best["train_mps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                   "sample": (f"{_cpu_what(r['kind'])}, {args.dataset}-shaped batches, fwd+loss+bwd+Adam, {r['threads']} torch "
                              f"threads; best of batch {cb} ({r['steps_timed']} timed steps: {r['train_mps']:.0f} m/s) and "
                              f"batch {B} = our per-GPU batch ({big['steps_timed']} timed steps: {big['train_mps']:.0f} m/s)"),
                   "value_by_batch": {str(cb): r["train_mps"], str(B): big["train_mps"]},
                   "ranking_value": max(r["rank_mps"], big["rank_mps"])}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_train, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if bf16 else "f32 (GEMMs as 3-pass split-bf16 on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": workload_config(args, world, B),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "ranking": ranking,
            "stage_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "ms_per_step_with_stage_events": ms_train_profiled,
            "necessary_gemm_tflops_of_step": (gemm_flops_per_mention(Cn, True, vector=args.edge_feature == "vector") * B /
                                              (ms_train * 1e-3) / 1e12),
            "legs": legs,
        }
        print(json.dumps(line), flush=True)
        if world > 1:
            self.dist.destroy_process_group()

    def e2e(self, trainer, batch, B, Cn, bf16):
        import drin_b200
        args, torch, world, dev = self.args, self.torch, self.world, self.dev
        steps, warm = args.steps, max(args.warmup, 3)
        host = [t.cpu().pin_memory() for t in batch]
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def run_e2e(compact, k):
            feeder = drin_b200.HostFeeder(dev, slots=2, compact_spans=compact)
            state = {"slot": feeder.submit(host)}

            def e2e_step():
                # double buffered: the host gather + copy of step i+1 overlap the compute of step i; all of it is inside
                # the timed region, and the loss is read back every step (reference train.py:35)
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

            ms = self.timed(e2e_step, k, warm, after=drain)
            nbytes = feeder.last_bytes
            del feeder
            return ms, nbytes

        ms_full, bytes_full = run_e2e(False, max(steps // 4, 3))     # verbatim copy: reported beside, fewer iterations
        ms_e2e, bytes_e2e = run_e2e(True, steps)                     # the e2e value: the CLI's --steps
        peak = self.h2d_ceiling()
        gbs = world * bytes_e2e / (ms_e2e * 1e-3) / 1e9
        e2e = {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bytes_e2e,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "steps": steps, "warmup": warm,
               "h2d_gbs": gbs, "h2d_peak_gbs": peak, "h2d_frac_of_peak": gbs / peak,
               "note": ("pinned host batch -> device every step on a side stream (double buffered), loss read back; "
                        "HostFeeder copies only the bytes the path reads: the start:end span rows of mention_text_feature are "
                        "gathered on the host (the other rows are never read).  h2d_gbs = aggregate host->device rate of the "
                        "leg over all ranks; h2d_peak_gbs = one 1 GiB pinned copy per rank, all ranks at once, measured in "
                        "this run: the leg is bound by the host->device fabric of the box, not by the GPUs"),
               "full_copy": {"value": world * B / (ms_full * 1e-3), "h2d_bytes_per_step": bytes_full,
                             "ms_per_step": ms_full, "note": "all 15 tensors copied verbatim"}}
        del host

        # ---------------- resident feature store: tables in HBM, only indices cross PCIe -------
        from drin_b200.store import FeatureStore, synthetic_tables
        n_store = 3 * B
        tables = synthetic_tables(args.dataset, n_store, seed=2000 + self.rank, num_candidates=Cn - 1, device=str(dev))
        store = FeatureStore(args.dataset, tables, Cn, device=dev,
                             feature_dtype=torch.bfloat16 if bf16 else torch.float32)
        del tables
        order = torch.randperm(n_store, generator=torch.Generator().manual_seed(self.rank)).pin_memory()
        cursor = {"k": 0}

        def store_step():
            k = cursor["k"]
            cursor["k"] = (k + 1) % 3
            loss = trainer.step(store.select(order[k * B:(k + 1) * B]))      # host indices -> device, 8 B / mention
            loss_host.copy_(loss.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(loss_host)

        ms_store = self.timed(store_step, steps, warm)
        e2e["resident_store"] = {
            "value": world * B / (ms_store * 1e-3), "unit": UNIT, "ms_per_step": ms_store,
            "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 4, "resident_bytes_per_gpu": store.nbytes(),
            "mentions_resident_per_gpu": n_store,
            "note": ("FeatureStore: the split's feature tables are uploaded once and stay in HBM; each step ships the "
                     "[B] mention indices from pinned host memory, the front-end kernel gathers rows by index "
                     "(drin/data.py:85-108 on device), loss read back every step")}
        del store
        torch.cuda.empty_cache()
        return e2e


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
    Bench(args).run()


if __name__ == "__main__":
    main()
