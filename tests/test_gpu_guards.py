"""Memory-safety checks of the hand-written kernels WITHOUT compute-sanitizer (the GPU pool refuses it: "runs under it
have left GPUs needing a reset", profiles/r02_sanitizer_unavailable.md).  Own instrumentation instead:

  * guard bands: with ``drin_debug_option("workspace_guard", n)`` every buffer of the workspace plan is followed by n
    untouchable bytes.  The whole workspace is filled with 0xFF (NaN in fp32 and bf16), a train step + a ranking forward
    run, and every band must still be 0xFF afterwards            -> no out-of-bounds WRITE between workspace buffers;
  * poison equivalence: scores, loss and every gradient of that run must be finite and BIT-IDENTICAL to a run whose
    workspace was zero-filled                                    -> no READ of workspace memory the step did not write
    itself (uninitialised or stale data would have to be both NaN-proof and zero-proof to go unnoticed);
  * external buffers: the dead-parameter slots of the gradient buffer and a tail behind it keep their pattern.

Covers the golden-size cases with every kernel variant forced, and at-scale steps where the engine picks the warp-per-
mention / sliced / column-wise kernels and cta_group::2 GEMM tiles by itself."""
import ctypes as C

import pytest
import torch

import drin_b200
from drin_b200 import _lib
from drin_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu
VARIANTS = ("score_bwd_variant", "score_fwd_variant", "layer_fwd_variant", "layer_bwd_variant")
FEATS = (0, 4, 5, 7, 9, 10)


def _option(name, v):
    _lib.check(_lib.load().drin_debug_option(name.encode(), C.c_int32(v)), name)


def _guard_regions(eng, cfg):
    lib = _lib.load()
    n, g = C.c_int32(0), C.c_size_t(0)
    _lib.check(lib.drin_debug_guard_regions(C.byref(cfg), None, C.c_int32(0), C.byref(n), C.byref(g)), "guard_regions")
    offs = (C.c_size_t * n.value)()
    _lib.check(lib.drin_debug_guard_regions(C.byref(cfg), offs, n, C.byref(n), C.byref(g)), "guard_regions")
    return list(offs), g.value


def _run(dataset, B, cands, fill, bf16=False, forced=-1, **model_kw):
    batch = make_batch(dataset, B, 91, cands, device="cuda", generate_on_device=True,
                       **(dict(entity_tokens=16, mention_tokens=32) if (dataset == "wikimel" and cands < 100) else {}))
    if bf16:
        batch = [t.to(torch.bfloat16) if i in FEATS else t for i, t in enumerate(batch)]
    torch.manual_seed(0)
    model = drin_b200.Model(num_candidates_model=cands + 1, **model_kw).cuda()
    with torch.no_grad():
        for l in model.gcn_layers:
            l.w_h.weight.mul_(3.0)
    tr = drin_b200.Trainer(model)
    for name in VARIANTS:
        _option(name, forced)
    try:
        tr.forward_backward(batch)                      # allocates the workspace and the loss scratch
        torch.cuda.synchronize()
        eng = model._engine
        assert len(eng.pool.entries) == 1
        ws = eng.pool.entries[0]["ws"]
        ws.fill_(fill)
        model.flat_grads.view(torch.uint8).fill_(0xA5)  # dead-parameter and tail slots must keep this pattern
        loss = tr.forward_backward(batch)
        scores = tr.last_scores.clone()
        grads = model.flat_grads.clone()
        ranked = tr.rank_scores(batch).clone()
        torch.cuda.synchronize()
        from drin_b200 import engine as E
        pb = E.inspect_batch(tuple(batch[:-1]), cands + 1)
        offs, guard = _guard_regions(eng, eng.config(pb, True))
        return dict(loss=loss.clone(), scores=scores, grads=grads, ranked=ranked, ws=ws, offs=offs, guard=guard,
                    dead=model.skip_mask().bool())
    finally:
        for name in VARIANTS:
            _option(name, -1)


CASES = [
    ("wikidiverse", 8, 10, False, -1, {}), ("wikidiverse", 8, 10, False, 1, {}), ("wikidiverse", 8, 10, False, 2, {}),
    ("wikimel", 3, 100, False, -1, {}), ("wikimel", 3, 100, False, 1, {}), ("wikimel", 6, 5, False, 2, {}),
    ("wikidiverse", 8, 10, False, -1, dict(gcn_edge_feature="vector")),
    ("wikidiverse", 6, 10, False, -1, dict(gcn_edge_feature="vector", num_gcn_layers=3)),
    ("wikidiverse", 8, 10, False, -1, dict(gcn_edge_type="static")),
    ("wikidiverse", 8, 10, True, -1, {}), ("wikimel", 3, 100, True, -1, {}),
    ("wikidiverse", 1280, 10, False, -1, {}), ("wikimel", 148, 100, False, -1, {}), ("wikimel", 37, 100, False, -1, {}),
    ("wikidiverse", 1300, 10, False, -1, dict(gcn_edge_feature="vector")), ("wikidiverse", 1280, 10, True, -1, {}),
]


@pytest.mark.parametrize("dataset,B,cands,bf16,forced,kw", CASES,
                         ids=[f"{c[0][:5]}-B{c[1]}-C{c[2] + 1}{'-bf16' if c[3] else ''}-v{c[4]}" +
                              "".join(f"-{v}" for v in c[5].values()) for c in CASES])
def test_guard_bands_and_poison_equivalence(dataset, B, cands, bf16, forced, kw):
    _option("workspace_guard", 1024)
    try:
        poisoned = _run(dataset, B, cands, 0xFF, bf16, forced, **kw)
        zeroed = _run(dataset, B, cands, 0x00, bf16, forced, **kw)
    finally:
        _option("workspace_guard", 0)
    ws, guard = poisoned["ws"], poisoned["guard"]
    assert guard == 1024 and len(poisoned["offs"]) > 20
    touched = [off for off in poisoned["offs"] if not bool((ws[off:off + guard] == 0xFF).all())]
    assert not touched, f"{len(touched)} guard band(s) were written, first at workspace offset {touched[0]}"
    for k in ("loss", "scores", "grads", "ranked"):
        a, b = poisoned[k], zeroed[k]
        if k == "grads":                                # dead slots: pattern untouched in both runs
            dead = poisoned["dead"]
            assert torch.equal(a[dead].view(torch.uint8), torch.full_like(a[dead].view(torch.uint8), 0xA5))
            a, b = a[~dead], b[~dead]
        assert bool(torch.isfinite(a).all()), f"{k}: non-finite values with a poisoned workspace"
        assert torch.equal(a, b), f"{k}: depends on the previous contents of the workspace"
