"""``drin_b200.fit`` = the reference's ``main()`` loop (train.py:125-147): blocks of epochs with a fresh Adam per block,
shuffled train batches (last one short), validation per epoch, a test pass per block, top-k accuracy with the
acc_correction division (train.py:38) -- here against a hand-written loop of eager steps in the same order."""
import io

import pytest
import torch

import drin_b200
from drin_b200.store import FeatureStore, synthetic_tables

pytestmark = pytest.mark.gpu


def _stores(cands=10):
    out = []
    for n, seed in ((88, 1), (24, 2), (20, 3)):
        t = synthetic_tables("wikidiverse", n, seed=seed, num_candidates=cands, device="cuda")
        out.append(FeatureStore("wikidiverse", t, cands + 1, device="cuda"))
    return out


def _model():
    torch.manual_seed(0)
    m = drin_b200.Model(num_candidates_model=11).cuda()
    with torch.no_grad():
        for l in m.gcn_layers:
            l.w_h.weight.mul_(3.0)
    return m


@pytest.mark.parametrize("graph", [True, False], ids=["graphed", "eager"])
def test_fit_equals_hand_written_reference_loop(graph, tmp_path):
    train, valid, test = _stores()
    B, epochs, interval, corr = 16, 4, 2, (0.1, 0.2, 0.0)
    m1, m2 = _model(), _model()
    buf = io.StringIO()
    hist = drin_b200.fit(m1, train, valid, test, batch_size=B, num_epoch=epochs, test_epoch_interval=interval, seed=5,
                         acc_correction=corr, graph=graph, result_file=buf, log=None)
    # the same schedule by hand: eager steps, one optimizer per block, same permutations
    gen = torch.Generator().manual_seed(5)
    losses, k = [], 0
    for block in range(epochs // interval):
        tr = drin_b200.Trainer(m2, lr=1e-3, margin=0.25)
        for _ in range(interval):
            tot, n = 0.0, 0
            for idx in torch.randperm(len(train), generator=gen).split(B):       # 88 = 5 x 16 + 8: short last batch
                tot += float(tr.step(train.select(idx)))
                n += 1
            losses.append(tot / n)
    assert torch.equal(m1.flat_params, m2.flat_params)                           # graph replay == eager, same order
    tr_recs = [r for r in hist if r["type"] == "training"]
    assert [r["epoch"] for r in tr_recs] == [1, 2, 3, 4] and all(r["steps"] == 6 for r in tr_recs)
    assert all(abs(r["loss"] - l) <= 1e-6 * abs(l) for r, l in zip(tr_recs, losses))
    assert tr_recs[-1]["loss"] < tr_recs[0]["loss"]
    assert [r["type"] for r in hist] == ["training", "validating"] * 2 + ["testing"] + ["training", "validating"] * 2 + ["testing"]
    # metrics: Evaluator on the final weights reproduces the last test record; acc_correction divides the accuracy
    ev = drin_b200.evaluate(m2, test, B, 0.25, (1, 3, 5)).compute(corr[2])
    assert hist[-1]["topk"] == ev["topk"] and abs(hist[-1]["loss"] - ev["loss"]) < 1e-7
    va = drin_b200.evaluate(m2, valid, B, 0.25, (1, 3, 5))
    assert va.compute(corr[1])["topk"][1] == pytest.approx(va.compute(0.0)["topk"][1] / (1 - corr[1]))
    # result file: one banner per test pass, "{index}:\t[scores]\n{labels}\n" per mention (train.py:40-43,94-96)
    text = buf.getvalue()
    assert text.count("==========  Test ==========") == 2 and text.count(":\t[") == 2 * len(test)
    if graph:       # a path instead of a file object: opened once, every test pass appends (train.py:16-17)
        path = tmp_path / "test-result.txt"
        drin_b200.fit(_model(), train, None, test, batch_size=B, num_epoch=2, test_epoch_interval=1, seed=5, graph=False,
                      result_file=str(path), log=None)
        assert path.read_text().count("==========  Test ==========") == 2
