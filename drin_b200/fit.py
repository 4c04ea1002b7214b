"""The caller of the hot path: the reference's ``main()`` / Lightning loop (upstream train.py:125-147 with
``MELModel`` train.py:20-56 and ``EpochLogger`` train.py:59-99) over resident ``FeatureStore`` splits, without
Lightning and without a host synchronisation per step.

Semantics kept from the reference:
  * ``num_epoch // test_epoch_interval`` blocks; every block builds a NEW optimizer (upstream creates a fresh
    ``pl.Trainer`` per block, train.py:141-143, so Adam's moments restart) and ends with a pass over the test split;
  * an epoch = one pass over the shuffled train split in batches of ``batch_size`` (``DataLoader(shuffle=True)``,
    drin/data.py:155; the last batch may be short), followed by a pass over the validation split;
  * per step ``y_hat = model(batch[:-1]); loss = TripletLoss(y, y_hat)`` + Adam (train.py:32-34,55-56); the metrics are
    the threshold top-k accuracies of common/utils.py:60-66 divided by ``1 - acc_correction[type]`` (train.py:38).
What differs: the loss / hit counters stay on the device and are read once per epoch (the reference prints them every
step, which forces a sync per step); full batches replay ONE captured CUDA graph (``GraphedStoreStep``); the shuffle
order comes from ``torch.randperm`` under ``seed`` (Lightning's sampler order is not reproduced).

Data parallel (``group`` / an initialised process group): ``batch_size`` is the GLOBAL batch -- every rank passes the
same stores and takes rows ``rank::world`` of each batch; the loss is the reference's at the global batch size.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .loss import TopkAccuracy
from .model import Model
from .store import FeatureStore
from .trainer import Evaluator, GraphedStoreStep, Trainer, _dist_on


def _batches(n: int, batch_size: int, perm: Optional[torch.Tensor]) -> List[torch.Tensor]:
    order = perm if perm is not None else torch.arange(n)
    return list(order.split(batch_size))


def evaluate(model: Model, store: FeatureStore, batch_size: int, margin: float, top_k: Sequence[int],
             acc_correction: float = 0.0, keep_scores: bool = False) -> Evaluator:
    """One pass over a split (validation_step / test_step of train.py:49-53): returns the filled Evaluator."""
    ev = Evaluator(model, margin, top_k, keep_scores=keep_scores)
    for idx in _batches(len(store), batch_size, None):
        ev.step(store.select(idx))
    return ev


def fit(model: Model, train: FeatureStore, valid: Optional[FeatureStore] = None, test: Optional[FeatureStore] = None, *,
        batch_size: int = 64, num_epoch: int = 30, test_epoch_interval: int = 10, lr: float = 1e-3, margin: float = 0.25,
        top_k: Sequence[int] = (1, 3, 5), acc_correction: Sequence[float] = (0.0, 0.0, 0.0), seed: int = 0,
        shuffle: bool = True, graph: bool = True, group=None, result_file=None,
        log: Optional[Callable[[str], None]] = print) -> List[Dict]:
    """Train like upstream ``main()``; returns one record per epoch (and per test pass) with loss and top-k accuracies."""
    world = dist.get_world_size(group) if _dist_on(group) else 1
    rank = dist.get_rank(group) if _dist_on(group) else 0
    if batch_size % world:
        raise ValueError(f"the global batch size {batch_size} must be a multiple of the number of ranks {world}")
    gen = torch.Generator().manual_seed(seed)                      # same order on every rank
    history: List[Dict] = []
    own_file = None
    if isinstance(result_file, (str, bytes)) or hasattr(result_file, "__fspath__"):
        # the reference opens test-result.txt once and every test pass appends to it (train.py:16-17,40-43,94-96)
        own_file = result_file = open(result_file, "w") if rank == 0 else None
    say = log if (log is not None and rank == 0) else (lambda s: None)
    blocks = max(num_epoch // test_epoch_interval, 1)
    epoch = 0
    for block in range(blocks):
        trainer = Trainer(model, lr=lr, margin=margin, group=group)          # fresh Adam state per block (train.py:141-143)
        graphed = None
        if graph and world == 1 and len(train) >= batch_size:
            graphed = GraphedStoreStep(trainer, train, batch_size)
        for _ in range(min(test_epoch_interval, num_epoch - epoch)):
            epoch += 1
            perm = torch.randperm(len(train), generator=gen) if shuffle else None
            metric = TopkAccuracy(list(top_k), device=train.device)
            loss_sum = torch.zeros(1, dtype=torch.float32, device=train.device)
            steps = 0
            for idx in _batches(len(train), batch_size, perm):
                if idx.numel() % world:
                    idx = idx[:idx.numel() - idx.numel() % world]           # equal shards (drops < world mentions)
                    if idx.numel() == 0:
                        continue
                local = idx[rank::world] if world > 1 else idx
                if graphed is not None and idx.numel() == batch_size:
                    loss, scores = graphed.step(local), graphed.scores
                    labels = train.onehot[train.answer[graphed.idx]]
                else:
                    sel = train.select(local)
                    loss, scores, labels = trainer.step(sel), trainer.last_scores, sel.labels
                loss_sum += loss.reshape(1)
                metric.update(scores, labels)                               # train.py:36-37, device-side counters
                steps += 1
            correct, total = metric.correct, metric.total
            if world > 1:                                                   # every rank counted its own shard
                correct = correct.clone()
                dist.all_reduce(correct, op=dist.ReduceOp.SUM, group=group)
                total *= world
            acc = (correct.double() / max(total, 1) / (1.0 - acc_correction[0])).tolist()
            rec = {"epoch": epoch, "type": "training", "loss": float(loss_sum) / max(steps, 1), "steps": steps,
                   "topk": dict(zip(metric.top_k, acc))}
            history.append(rec)
            say(f"***** Epoch {epoch}/{num_epoch} - training - loss: {rec['loss']:.5f}\t" +
                "\t".join(f"top-{k}: {a:.5f}" for k, a in rec["topk"].items()))
            if valid is not None:
                r = evaluate(model, valid, batch_size, margin, top_k).compute(acc_correction[1])
                rec = {"epoch": epoch, "type": "validating", "loss": r["loss"], "topk": r["topk"]}
                history.append(rec)
                say(f"***** Epoch {epoch}/{num_epoch} - validating - loss: {rec['loss']:.5f}\t" +
                    "\t".join(f"top-{k}: {a:.5f}" for k, a in rec["topk"].items()))
        if test is not None:
            ev = evaluate(model, test, batch_size, margin, top_k, keep_scores=result_file is not None)
            r = ev.compute(acc_correction[2])
            if result_file is not None and rank == 0:
                result_file.write("==========  Test ==========\n")          # train.py:94-96
                ev.write_results(result_file, batch_size)
            rec = {"epoch": epoch, "type": "testing", "loss": r["loss"], "topk": r["topk"]}
            history.append(rec)
            say(f"***** Epoch {epoch}/{num_epoch} - testing - loss: {rec['loss']:.5f}\t" +
                "\t".join(f"top-{k}: {a:.5f}" for k, a in rec["topk"].items()))
        del graphed
    if own_file is not None:
        own_file.close()
    say("Training completed")
    return history
