"""Workspace ownership of the autograd path (engine.py:295, trainer.py:208,308): every training forward owns
its saved activations until its backward has run, like the reference nn.Module."""
import pytest
import torch

import drin_b200
from drin_b200.synthetic import make_batch, spread_weights
from oracle import drin_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _setup(B1=12, B2=7):
    cfg = O.DrinConfig(num_candidates_model=11)
    sd = spread_weights(O.init_state(cfg, 0))
    m = drin_b200.Model(num_candidates_model=11)
    m.load_state_dict(sd)
    return cfg, sd, m.cuda(), make_batch("wikidiverse", B1, 71, 10), make_batch("wikidiverse", B2, 72, 10)


def _oracle_sum_grads(cfg, sd, b1, b2):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    l = O.triplet_loss(b1[-1], O.forward(leaves, b1[:-1], cfg), cfg.triplet_margin) + \
        O.triplet_loss(b2[-1], O.forward(leaves, b2[:-1], cfg), cfg.triplet_margin)
    l.backward()
    return float(l), {k: v.grad for k, v in leaves.items()}


def test_two_forwards_before_one_backward():
    """loss(model(b1)) + loss(model(b2)): the second (smaller) forward must not overwrite the first one's activations."""
    cfg, sd, m, b1, b2 = _setup()
    l_ref, g_ref = _oracle_sum_grads(cfg, sd, b1, b2)
    d1, d2 = [t.cuda() for t in b1], [t.cuda() for t in b2]
    lf = drin_b200.TripletLoss(cfg.triplet_margin)
    loss = lf(d1[-1], m(d1[:-1])) + lf(d2[-1], m(d2[:-1]))
    loss.backward()
    assert abs(float(loss) - l_ref) <= 1e-4 * abs(l_ref)
    for k, p in m.named_parameters():
        if g_ref[k] is not None:
            assert rel_err(p.grad.cpu(), g_ref[k]) < 1e-4, k
    assert len(m._engine.pool.entries) == 2


def test_no_grad_forward_between_forward_and_backward():
    cfg, sd, m, b1, b2 = _setup(12, 12)
    d1, d2 = [t.cuda() for t in b1], [t.cuda() for t in b2]
    _, l_ref, g_ref = O.train_step_grads(sd, b1[:-1], b1[-1], cfg)
    scores = m(d1[:-1])
    with torch.no_grad():
        m(d2[:-1])                                   # a validation forward of the same size in between
    loss = drin_b200.TripletLoss(cfg.triplet_margin)(d1[-1], scores)
    loss.backward()
    assert abs(float(loss) - float(l_ref)) <= 1e-4 * abs(float(l_ref))
    for k, p in m.named_parameters():
        if g_ref[k] is not None:
            assert rel_err(p.grad.cpu(), g_ref[k]) < 1e-4, k


def test_train_loop_keeps_one_workspace_and_stale_backward_raises():
    cfg, sd, m, b1, b2 = _setup(12, 12)
    d1 = [t.cuda() for t in b1]
    lf = drin_b200.TripletLoss(cfg.triplet_margin)
    first = None
    for _ in range(4):                               # steady state: the finished lease is reclaimed by the next forward
        loss = lf(d1[-1], m(d1[:-1]))
        loss.backward(retain_graph=first is None)
        first = first if first is not None else loss
    assert len(m._engine.pool.entries) == 1
    with pytest.raises(RuntimeError, match="activations of this forward are gone"):
        first.backward()                             # its workspace was taken over by a later forward: loud, not wrong


def test_graphed_step_survives_other_batch_sizes():
    """GraphedStoreStep keeps the workspace and loss scratch it captured alive: an Evaluator step with another batch size
    (larger, so the engine re-plans) in between must not disturb replays (trainer.py:208)."""
    from drin_b200.store import FeatureStore, synthetic_tables
    cands, n = 10, 96
    tables = synthetic_tables("wikidiverse", n, seed=5, num_candidates=cands, device="cuda")
    store = FeatureStore("wikidiverse", tables, cands + 1, device="cuda")
    torch.manual_seed(0)
    m1, m2 = drin_b200.Model(num_candidates_model=cands + 1).cuda(), drin_b200.Model(num_candidates_model=cands + 1).cuda()
    m2.load_state_dict(m1.state_dict())
    t1, t2 = drin_b200.Trainer(m1), drin_b200.Trainer(m2)
    g = drin_b200.GraphedStoreStep(t1, store, 16)
    ev = drin_b200.Evaluator(m1)
    order = torch.randperm(n, generator=torch.Generator().manual_seed(1))
    for i in range(4):
        idx = order[i * 16:(i + 1) * 16]
        l_graph = g.step(idx).clone()
        l_eager = t2.step(store.select(idx.cuda()))
        assert torch.equal(l_graph, l_eager.reshape_as(l_graph))
        ev.step(store.select(order[:64 + 8 * i].cuda()))      # bigger problem: new workspace, new loss scratch
    assert torch.equal(m1.flat_params, m2.flat_params)
