"""Drop-in proof with the reference's OWN glue: ``drin_b200.Model()`` constructed with no arguments under
the reference's real ``common.args`` (train.py:13-14,136), scored by the reference's own ``TripletLoss``
(common/utils.py:26-43, plain autograd on the CUDA scores) and stepped by ``torch.optim.Adam`` (train.py:55-56) -- and
the same through the fused ``FusedAdam``.  The reference modules come from ``oracle/_ref`` (staged by
``oracle/make_ref.py``) or, in the build container, from /root/reference; skipped when neither exists."""
import pytest
import torch

import drin_b200
from drin_b200.synthetic import make_batch
from oracle import ref_import
from tests.helpers import rel_err

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference files not staged (python oracle/make_ref.py)")


@needs_ref
@pytest.mark.parametrize("overrides", [dict(), dict(gcn_edge_feature="vector", num_gcn_layers=3),
                                       dict(gcn_edge_type="static", gcn_edge_enabled=[1, 0, 1, 1])],
                         ids=["default", "vector_l3", "static_masked"])
def test_no_arg_constructor_reads_the_real_common_args(overrides):
    """Model() takes every hyper-parameter from the reference's flat config module, like upstream's star-imports, and
    creates the same parameters in the same order (bit-identical under the same seed)."""
    m_ref, u_ref, a = ref_import.load("wikimel", 100, 64, **overrides)
    torch.manual_seed(0)
    ref = m_ref.Model()
    torch.manual_seed(0)
    ours = drin_b200.Model()                        # no arguments: train.py:136
    assert ours.num_candidates_model == a.num_candidates_model == 101
    assert ours.num_gcn_layers == a.num_gcn_layers
    assert ours.vector_edges == (a.gcn_edge_feature == "vector") and ours.static_edges == (a.gcn_edge_type == "static")
    assert tuple(ours.cfg["gcn_edge_enabled"]) == tuple(a.gcn_edge_enabled)
    sd_ref, sd = ref.state_dict(), ours.state_dict()
    assert list(sd.keys()) == list(sd_ref.keys())
    assert all(torch.equal(sd[k], sd_ref[k]) for k in sd)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("fused_adam", [False, True], ids=["torch_adam", "fused_adam"])
def test_one_train_step_with_the_reference_loss_and_optimizer(fused_adam):
    """The documented "plain autograd" integration: reference loss object on our CUDA scores,
    optimizer over our parameters.  One step must land on the reference's own post-step weights."""
    m_ref, u_ref, a = ref_import.load("wikidiverse", 10)
    batch = make_batch("wikidiverse", 16, 7, 10)
    torch.manual_seed(0)
    ref = m_ref.Model()
    torch.manual_seed(0)
    ours = drin_b200.Model().cuda()
    with torch.no_grad():                           # spread the scores past the margin (both hinge branches active)
        for mod in (ref, ours):
            for l in mod.gcn_layers:
                l.w_h.weight.mul_(3.0)
    loss_ref_fn = u_ref.TripletLoss(a.triplet_margin)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=a.learning_rate)
    y_hat_ref = ref(batch[:-1])
    loss_ref = loss_ref_fn(batch[-1], y_hat_ref)
    opt_ref.zero_grad()
    loss_ref.backward()
    opt_ref.step()

    db = [t.cuda() for t in batch]
    opt = drin_b200.FusedAdam(ours, lr=a.learning_rate) if fused_adam else torch.optim.Adam(ours.parameters(),
                                                                                           lr=a.learning_rate)
    y_hat = ours(db[:-1])                           # train.py:33
    loss = loss_ref_fn(db[-1], y_hat)               # the reference's own TripletLoss, ATen ops on the GPU
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert rel_err(y_hat.detach().cpu(), y_hat_ref.detach()) < 1e-4
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    for (k, p), (_, r) in zip(ours.named_parameters(), ref.named_parameters()):
        assert (p.grad is None) == (r.grad is None), k
        if r.grad is not None:
            assert rel_err(p.grad.cpu(), r.grad) < 1e-4, k
        # Adam's first step moves every live weight by lr * g / (|g| + eps): entries whose gradient is within noise of zero
        # (|g| ~ eps = 1e-8) may move differently by up to 2 lr; everything else must land on the reference's weights
        diff = (p.detach().cpu() - r.detach()).abs()
        assert float(diff.max()) <= 2.0 * a.learning_rate + 1e-6, k
        assert float((diff > 1e-5).double().mean()) < 2e-3, k
