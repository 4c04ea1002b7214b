"""CPU-only checks: the C-ABI library loads and exports every symbol include/drin_b200.h declares, the
host-side mirror of the reference interface (constructor, state_dict, batch validation), no CPU fallback."""
import ctypes as C
import os
import re

import pytest
import torch

import drin_b200
from drin_b200 import _lib, engine as E
from drin_b200.synthetic import make_batch
from oracle import drin_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "drin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drin_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/drin_b200.h but not exported"
    assert lib.drin_version() >= 100


def test_workspace_plan_and_config_errors_need_no_gpu():
    eng = E.Engine(2)
    pb = E.Problem(4096, 11, 128, 0, 49, 3, 1, 768, 2048, E.FP32)
    infer, train = eng.workspace_bytes(eng.config(pb, False)), eng.workspace_bytes(eng.config(pb, True))
    assert 0 < infer < train < 8 << 30
    bad = eng.config(E.Problem(4, 11, 128, 0, 49, 3, 1, 512, 2048, E.FP32), False)
    with pytest.raises(RuntimeError, match="gcn_embed_dim"):
        eng.workspace_bytes(bad)


def test_vector_edge_workspace_plan():
    """Workspace plan of the vector-edge path (csrc/engine.cu: plan_vector): the two-layer model needs no [4BC, D] edge
    matrices thanks to the first-layer shortcut, every further layer adds them; static edges and one layer fall back to
    the scalar plan exactly."""
    pb = E.Problem(4096, 11, 128, 0, 49, 3, 1, 768, 2048, E.FP32)
    edge = 4 * 4096 * 11 * 768                      # elements of one [4BC, D] matrix

    def need(layers, training, **kw):
        eng = E.Engine(layers, **kw)
        return eng.workspace_bytes(eng.config(pb, training))

    s2, v2, v3, v4 = need(2, True), need(2, True, vector_edges=True), need(3, True, vector_edges=True), need(4, True, vector_edges=True)
    assert s2 < v2 < s2 + 4 * edge * 4              # two layers: row-sized extras only (no m / q / dm / dq: 6 edge matrices)
    assert v3 - v2 > 3 * edge * 4                   # m planes + q per general layer, dm + dq planes once
    assert v4 - v3 > 2 * edge * 4
    assert need(2, True, vector_edges=True, static_edges=True) == need(2, True, static_edges=True)
    assert need(1, True, vector_edges=True) == need(1, True)
    assert need(2, False, vector_edges=True) < v2 < 16 << 30


def test_model_is_a_drop_in_for_the_reference_constructor():
    torch.manual_seed(0)
    m = drin_b200.Model()
    sd = O.init_state(O.DrinConfig(), 0)             # pinned bit-for-bit to the reference init by make_golden.py
    assert list(m.state_dict().keys()) == O.state_dict_keys(O.DrinConfig())
    assert all(torch.equal(sd[k], v) for k, v in m.state_dict().items())
    assert sum(p.numel() for p in m.parameters()) == 7_875_072
    # parameters are views of one flat buffer and stay so across load_state_dict
    m.load_state_dict({k: v * 2 for k, v in sd.items()})
    assert torch.equal(m.flat_params[:768 * 768].view(768, 768), sd[O.K_MT_W] * 2)
    assert int(m.dead_mask().sum()) == 2 * (768 * 768 + 768)


def test_flat_layout_keeps_dead_parameters_behind_the_allreduce_bucket():
    """Flat buffer = [live parameters | 4 tail floats | dead parameters]: the data-parallel all-reduce bucket is the
    leading n_live + 4 floats (26.8 MB) and carries no dead zeros; state_dict order is unchanged."""
    for kw, n_dead in ((dict(), 2 * (768 * 768 + 768)), (dict(gcn_edge_type="static"), 4 * (768 * 768 + 768)),
                       (dict(gcn_edge_feature="vector"), 768 * 768 + 768 + 2 * (384 * 768 + 384))):
        m = drin_b200.Model(**kw)
        total = sum(p.numel() for p in m.parameters())
        assert m.n_live == total - n_dead and m.flat_params.numel() == total + m.TAIL
        assert m.flat_grads_bucket.numel() == m.n_live + m.TAIL
        assert m.flat_grads_bucket.data_ptr() == m.flat_grads.data_ptr()
        dead = set(m._dead)
        for k, (o, n, shape) in m._offsets.items():
            assert (o >= m.n_live + m.TAIL) if k in dead else (o + n <= m.n_live), k
        assert int(m.skip_mask().sum()) == n_dead + m.TAIL and int(m.dead_mask().sum()) == n_dead
        assert m.layer_grad_offset() == 2 * (768 * 768 + 768) + 2 * (768 * 2048 + 768)
        # parameters are still views of the flat buffer, in state_dict order for the caller
        sd = {k: torch.full_like(v, float(i)) for i, (k, v) in enumerate(m.state_dict().items())}
        m.load_state_dict(sd)
        for i, k in enumerate(sd):
            o, n, _ = m._offsets[k]
            assert float(m.flat_params[o]) == float(i) == float(m.flat_params[o + n - 1]), k


def test_unsupported_configurations_fail_loudly():
    for kw in (dict(gcn_edge_feature="matrix"), dict(gcn_edge_type="learned"), dict(gcn_vertex_activation="relu"),
               dict(gcn_embed_dim=512)):
        with pytest.raises(NotImplementedError):
            drin_b200.Model(**kw)


def test_no_cpu_fallback():
    m = drin_b200.Model()
    batch = make_batch("wikidiverse", 2, 0)
    with pytest.raises(RuntimeError):
        m(batch[:-1])
    with pytest.raises(RuntimeError):
        drin_b200.TripletLoss(0.25)(batch[-1], torch.zeros(2, 11))


def test_install_as_reference_module():
    import sys
    drin_b200.install_as_reference_module()
    from drin import model as model_module          # what reference train.py:13-14 does
    assert model_module.Model is drin_b200.Model
    del sys.modules["drin.model"], sys.modules["drin"]


def test_param_key_tables():
    assert E.param_keys(2) == O.state_dict_keys(O.DrinConfig())
    assert E.dead_param_keys(2) == ["gcn_layers.1.w_u.weight", "gcn_layers.1.w_u.bias",
                                    "gcn_layers.1.w_v.weight", "gcn_layers.1.w_v.bias"]
    # gcn_edge_type="static": no layer has an edge update (drin/model.py:135-136)
    assert len(E.dead_param_keys(3, static_edges=True)) == 12
    assert int(drin_b200.Model(gcn_edge_type="static").dead_mask().sum()) == 4 * (768 * 768 + 768)


def test_vector_edge_model_is_a_drop_in_for_the_reference_constructor():
    """gcn_edge_feature="vector" (args.py:33): w_m becomes a Linear created right after w_h, w_u / w_v map D -> D/2
    (drin/model.py:111-116).  Same keys, creation order and seeded weights as the reference (pinned by make_golden.py)."""
    cfg = O.DrinConfig(gcn_edge_feature="vector")
    torch.manual_seed(0)
    m = drin_b200.Model(gcn_edge_feature="vector")
    sd = O.init_state(cfg, 0)
    assert list(m.state_dict().keys()) == O.state_dict_keys(cfg) == E.param_keys(2, vector_edges=True)
    assert all(torch.equal(sd[k], v) for k, v in m.state_dict().items())
    shapes = E.param_shapes(2, vector_edges=True)
    assert shapes["gcn_layers.0.w_u.weight"] == (384, 768) and shapes["gcn_layers.0.w_m.weight"] == (768, 768)
    assert all(tuple(v.shape) == shapes[k] for k, v in sd.items())
    # the last layer's edge update is dead code: w_m, w_u, w_v of layer 1 never get a gradient
    dead = E.dead_param_keys(2, vector_edges=True)
    assert dead == [f"gcn_layers.1.{s}" for s in ("w_m.weight", "w_m.bias", "w_u.weight", "w_u.bias", "w_v.weight",
                                                  "w_v.bias")]
    assert int(m.dead_mask().sum()) == 768 * 768 + 768 + 2 * (384 * 768 + 384)


def test_workspace_pool_lease_semantics():
    """Host logic of Engine.pool (no GPU needed): a live lease is never handed out again, a finished ('done') lease is
    reclaimed by the next acquire and becomes invalid, released leases are reused, too-small free workspaces are dropped."""
    pool = E.WorkspacePool()
    dev = torch.device("cpu")
    a = pool.acquire(1000, dev)
    b = pool.acquire(500, dev)                      # a is live: a second workspace
    assert a.entry is not b.entry and len(pool.entries) == 2 and a.valid() and b.valid()
    b.release()
    c = pool.acquire(400, dev)                      # reuses the released one (smallest that fits)
    assert c.entry is b.entry and not b.valid() and c.valid()
    a.state = "done"                                # what Engine.backward sets
    assert a.valid()                                # a second backward right away is still fine
    c.release()
    d = pool.acquire(900, dev)                      # the free 500-byte one is too small: the done lease is reclaimed
    assert d.entry is a.entry and not a.valid() and d.valid()
    del d                                           # a dropped lease (autograd graph freed) frees its workspace
    import gc
    gc.collect()
    e = pool.acquire(2000, dev)                     # nothing fits: the free ones are dropped, one new workspace
    assert len(pool.entries) == 1 and pool.entries[0]["ws"].numel() == 2000 and e.valid()


def test_loss_scratch_cache_is_keyed_and_bounded():
    """ADVICE r1 (loss.py:36): the scratch of one (B, C) must survive calls with other shapes (a CUDA graph holds raw
    pointers into it); the cache is a small LRU, not a single slot."""
    from drin_b200 import loss as L
    L._scratch.clear()
    a = L._loss_scratch(64, 11, "cpu")
    b = L._loss_scratch(48, 11, "cpu")
    assert L._loss_scratch(64, 11, "cpu") is a and L._loss_scratch(48, 11, "cpu") is b     # alternating shapes: no churn
    for i in range(L._SCRATCH_KEEP + 3):
        L._loss_scratch(100 + i, 11, "cpu")
    assert len(L._scratch) == L._SCRATCH_KEEP and len(L.scratch_tensors()) == L._SCRATCH_KEEP
    assert a.numel() > 0                       # evicted from the cache, still alive for whoever kept a reference
    L._scratch.clear()
