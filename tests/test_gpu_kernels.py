"""Stage-level parity through the C ABI: tcgen05 GEMM (3 layouts x 2 modes), front end, loss, metric, Adam."""
import ctypes as C
import os

import pytest
import torch

import drin_b200
from drin_b200 import _lib, engine as E
from drin_b200.synthetic import make_batch
from oracle import drin_oracle as O
from tests.helpers import GOLDEN_DIR, rel_err

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _gemm(layout, planes, M, N, K, ksplit=1, bias=False, want_planes=False, reference=0):
    lib = _lib.load()
    torch.manual_seed(M + N + K)
    a = torch.randn((K, M) if layout == 2 else (M, K), device="cuda")
    b = torch.randn((N, K) if layout == 0 else (K, N), device="cuda")

    def split(x):
        hi = x.to(torch.bfloat16)
        return hi, ((x - hi.float()).to(torch.bfloat16) if planes == 2 else None)

    (a_hi, a_lo), (b_hi, b_lo) = split(a), split(b)
    a_eff = a_hi.double() + (a_lo.double() if a_lo is not None else 0)
    b_eff = b_hi.double() + (b_lo.double() if b_lo is not None else 0)
    ref = (a_eff.t() if layout == 2 else a_eff) @ (b_eff.t() if layout == 0 else b_eff)
    bias_t = torch.randn(N, device="cuda") if bias else None
    if bias:
        ref = ref + bias_t.double()
    out = torch.full((M, N), float("nan"), device="cuda")
    oh = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda") if want_planes else None
    ol = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda") if want_planes else None
    partial = torch.empty(ksplit * M * N, device="cuda") if ksplit > 1 else None
    st = lib.drin_gemm(C.c_int32(layout), _p(a_hi), _p(a_lo), C.c_int32(a.shape[1]), _p(b_hi), _p(b_lo),
                       C.c_int32(b.shape[1]), C.c_int64(M), C.c_int32(N), C.c_int64(K), _p(out), C.c_int32(N),
                       _p(bias_t), _p(oh), _p(ol), C.c_int32(N), C.c_int32(ksplit), _p(partial), C.c_int32(reference),
                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(st, "drin_gemm")
    torch.cuda.synchronize()
    err = rel_err(out, ref)
    perr = rel_err(oh.double() + ol.double(), ref) if want_planes else None
    return err, perr


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("planes", [1, 2])
def test_gemm_layouts(layout, planes):
    """NT / NN / TN, ragged M (TMA zero fill + masked epilogue), K tails, bias, plane outputs."""
    tol = 2e-5 if planes == 2 else 5e-6     # operands are pre-rounded: only accumulation order differs
    for (M, N, K) in ((128, 256, 64), (300, 768, 768), (77, 768, 2048), (768, 768, 1000)):
        if layout == 2:
            M = (M + 7) // 8 * 8          # A is stored [K, M]: TMA needs a 16-byte row pitch
        err, perr = _gemm(layout, planes, M, N, K, bias=(layout == 0), want_planes=(planes == 2 and layout != 2))
        assert err < tol, (layout, planes, M, N, K, err)
        if perr is not None:
            assert perr < 3e-5


def test_gemm_half_width_shapes_of_the_vector_edge_update():
    """W_u / W_v of the vector-edge update map 768 -> 384 (drin/model.py:113-116): N = 384 ends inside a 256-column
    tile (ragged N tile, TMA zero fill of the missing B rows), K = 384 is six k-blocks; both cta_group variants."""
    for planes in (1, 2):
        tol = 2e-5 if planes == 2 else 5e-6
        for layout, (M, N, K), ks in ((0, (16, 384, 768), 1), (0, (176, 384, 768), 1), (0, (1000, 384, 768), 1),
                                      (1, (16, 768, 384), 1), (1, (520, 768, 384), 1),
                                      (2, (384, 768, 16), 1), (2, (384, 768, 1100), 8), (2, (384, 768, 176), 2)):
            err, _ = _gemm(layout, planes, M, N, K, ksplit=ks, bias=(layout == 0))
            assert err < tol, (layout, planes, M, N, K, err)


def test_gemm_split_k_is_deterministic_and_exact():
    e1, _ = _gemm(2, 2, 768, 2048, 5000, ksplit=3)
    e2, _ = _gemm(2, 2, 768, 768, 9000, ksplit=8)
    assert e1 < 2e-5 and e2 < 2e-5


def test_gemm_reference_kernel_agrees():
    err, _ = _gemm(0, 2, 200, 512, 512, reference=1)
    assert err < 5e-6


@pytest.mark.parametrize("dataset,kw", [("wikidiverse", {}), ("wikimel", {}),
                                        ("wikimel", dict(num_candidates=5, entity_tokens=16, mention_tokens=32))])
def test_frontend_stage(dataset, kw):
    """drin_frontend vs the oracle: span mean, region mean, entity pooling, tt / ii edges, CLIP edges / 100."""
    B = 5
    batch = make_batch(dataset, B, 11, **kw)
    db = [t.cuda() for t in batch[:-1]]
    pb = E.inspect_batch(db)
    eng = E.Engine(2)
    cfg = eng.config(pb, False)
    span = torch.empty(pb.B, pb.D, device="cuda")
    mim = torch.empty(pb.B, pb.R, device="cuda")
    ep = torch.empty(pb.B * pb.C, pb.D, device="cuda")
    edges = torch.empty(4, pb.B * pb.C, device="cuda")
    ins = eng.inputs(db)
    _lib.check(eng.lib.drin_frontend(C.byref(cfg), C.byref(ins), _p(span), _p(mim), _p(ep), _p(edges),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "drin_frontend")
    torch.cuda.synchronize()
    b = batch[:-1]
    assert rel_err(span.cpu(), O.span_mean(b[0], b[2], b[3])) < 1e-6
    assert rel_err(mim.cpu(), b[4].mean(-2)) < 1e-6
    assert rel_err(ep.cpu(), O.entity_text_pool(b[7], b[8]).flatten(0, 1)) < 1e-6
    tt, ii = O.edge_encode(b)
    want = torch.stack([tt.flatten(), (b[13] / 100).flatten(), (b[12] / 100).flatten(), ii.flatten()])
    assert rel_err(edges.cpu(), want) < 2e-6
    # zero object scores are legal: ii = 0 / (0 + 1e-9) = 0
    z = [t.clone() for t in db]
    z[6].zero_()
    ins = eng.inputs(z)
    _lib.check(eng.lib.drin_frontend(C.byref(cfg), C.byref(ins), _p(span), _p(mim), _p(ep), _p(edges),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "drin_frontend")
    assert float(edges[3].abs().max()) == 0.0


def test_triplet_loss_and_topk_against_reference_fixtures():
    cases = torch.load(os.path.join(GOLDEN_DIR, "triplet_loss_cases.pt"), weights_only=False)
    for c in cases:
        s = c["scores"].cuda().requires_grad_(True)
        loss = drin_b200.TripletLoss(c["margin"])(c["labels"].cuda(), s)
        loss.backward()
        assert abs(float(loss) - float(c["loss"])) <= 1e-6 * max(1.0, abs(float(c["loss"])))
        assert rel_err(s.grad.cpu(), c["dscores"]) < 1e-5
        assert float(s.grad[:, -1].abs().max()) == 0.0
        met = drin_b200.TopkAccuracy([1, 3, 5])
        met.update(s.detach(), c["labels"].cuda())
        assert met.correct.tolist() == [c["topk_hits"][k] for k in (1, 3, 5)]


def test_sharded_loss_equals_global_loss():
    """Data-parallel form: shares of the loss add up, local gradient rows equal the global ones (bitwise)."""
    from drin_b200.loss import triplet_loss_sharded
    g = torch.Generator().manual_seed(5)
    B, Cn = 64, 11
    s = (torch.rand(B, Cn, generator=g) * 2 - 1).cuda()
    y = torch.eye(Cn - 1, dtype=torch.uint8)[torch.randint(0, Cn - 1, (B,), generator=g)].cuda()
    full_loss, full_d = triplet_loss_sharded(s, y, 0.25)
    full_loss, full_d = full_loss.clone(), full_d.clone()
    parts, shares = [], []
    for r in range(4):
        l, d = triplet_loss_sharded(s, y, 0.25, r * 16, 16)
        parts.append(d.clone())
        shares.append(float(l))
    assert torch.equal(torch.cat(parts), full_d)
    assert abs(sum(shares) - float(full_loss)) < 1e-6 * abs(float(full_loss))
    ref = O.triplet_loss(y.cpu(), s.cpu(), 0.25)
    assert abs(float(full_loss) - float(ref)) < 1e-6 * abs(float(ref))


def test_loss_across_sort_segments_matches_closed_form():
    """Global batches above 4096 mentions span several sort segments; shards need not be aligned to them.  Checked
    against the oracle's closed form (exact integer counts -> tight tolerance), with exact ties and all-zero label rows."""
    from drin_b200.loss import triplet_loss_sharded
    g = torch.Generator().manual_seed(9)
    B, Cn = 5000, 11
    s = torch.rand(B, Cn, generator=g) * 2 - 1
    s[:, 1] = s[:, 0]
    ans = torch.randint(0, Cn, (B,), generator=g)
    y = torch.cat([torch.eye(Cn - 1, dtype=torch.uint8), torch.zeros(1, Cn - 1, dtype=torch.uint8)])[ans]
    sd, yd = s.cuda(), y.cuda()
    shares, parts = [], []
    for row0, rows in ((0, 3000), (3000, 1500), (4500, 500)):          # the middle shard straddles segments 0 and 1
        l, d = triplet_loss_sharded(sd, yd, 0.25, row0, rows)
        want_l, want_d = O.triplet_sharded(s, y, 0.25, row0, rows)
        assert abs(float(l) - float(want_l)) <= 1e-6 * abs(float(want_l))
        assert rel_err(d.cpu(), want_d) < 1e-6
        shares.append(float(l))
        parts.append(d.clone())
    full_l, full_d = triplet_loss_sharded(sd, yd, 0.25)
    assert torch.equal(torch.cat(parts), full_d)
    assert abs(sum(shares) - float(full_l)) < 1e-6 * abs(float(full_l))


def test_fused_adam_matches_torch_adam():
    torch.manual_seed(0)
    model = drin_b200.Model().cuda()
    ref = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    names = [n for n, _ in model.named_parameters()]
    dead = set(E.dead_param_keys(2))
    opt_ref = torch.optim.Adam(ref, lr=1e-3)
    opt = drin_b200.FusedAdam(model, lr=1e-3)
    for step in range(3):
        g = torch.randn_like(model.flat_params) * (10.0 ** (-step))
        model.flat_grads.copy_(g)
        model._flat_grads_valid = True          # what Trainer.forward_backward sets after writing the flat buffer
        for n, r, (o, cnt, shape) in zip(names, ref, [model._offsets[n] for n in names]):
            r.grad = None if n in dead else g[o:o + cnt].view(shape).clone()
        opt.step()
        opt_ref.step()
    for n, p, r in zip(names, model.parameters(), ref):
        assert torch.allclose(p.detach(), r.detach(), rtol=2e-6, atol=1e-8), n


@pytest.mark.parametrize("variant", [-1, 1], ids=["auto", "warp-kernels"])
@pytest.mark.parametrize("dataset,B,cands,layers,kw", [("wikidiverse", 9, 10, 2, {}), ("wikidiverse", 8, 10, 3, {}),
                                                       ("wikimel", 3, 100, 2, {})])
def test_forward_intermediates_stage_by_stage(dataset, B, cands, layers, kw, variant):
    """Every stage of Model.forward against the oracle through the workspace test hook (drin_debug_buffer):
    projections x0 (model.py:26-46), edge list (model.py:201-204), per layer the activated vertices
    gelu(LN(h)) (model.py:128) and the dynamic edge update (model.py:131-134), then the scores."""
    import torch.nn.functional as F

    from drin_b200.synthetic import spread_weights
    for name in ("score_fwd_variant", "layer_fwd_variant"):
        _lib.check(_lib.load().drin_debug_option(name.encode(), C.c_int32(variant)), "drin_debug_option")
    try:
        cfg = O.DrinConfig(num_candidates_model=cands + 1, num_gcn_layers=layers)
        batch = make_batch(dataset, B, 31, cands, **kw)
        sd = spread_weights(O.init_state(cfg, 0))
        V = O.vertex_encode(sd, batch[:-1])
        tt, ii = O.edge_encode(batch[:-1])
        Ed = [tt, batch[13] / 100, batch[12] / 100, ii]
        eng = E.Engine(layers)
        scores, ctx = eng.forward([t.cuda() for t in batch[:-1]], {k: v.cuda() for k, v in sd.items()}, training=False)
        x0 = torch.cat([V[0], V[1], V[2].flatten(0, 1), V[3].flatten(0, 1)])
        assert rel_err(eng.debug_buffer(ctx, "x0").cpu(), x0) < 1e-5
        assert rel_err(eng.debug_buffer(ctx, "edges0").cpu(), torch.stack([e.flatten() for e in Ed])) < 2e-6
        Vl, El = V, Ed
        for l in range(layers):
            Vl, El = O.gcn_layer(sd, l, cfg, Vl, El)
            h = eng.debug_buffer(ctx, "h", l).cpu()
            k = O.gcn_keys(l)
            act = F.gelu(F.layer_norm(h, (h.shape[-1],), sd[k["ln_w"]], sd[k["ln_b"]], 1e-5))
            if l < layers - 1:
                want = torch.cat([Vl[0], Vl[1], Vl[2].flatten(0, 1), Vl[3].flatten(0, 1)])
                assert rel_err(eng.debug_buffer(ctx, "edges_out", l).cpu(), torch.stack([e.flatten() for e in El])) < 1e-5
            else:
                want = torch.cat([Vl[0], Vl[2].flatten(0, 1)])       # the last layer only updates mt / et
            assert rel_err(act, want) < 2e-5, l
        assert rel_err(scores.cpu(), O.cosine(Vl[0].unsqueeze(1), Vl[2])) < 1e-5
    finally:
        for name in ("score_fwd_variant", "layer_fwd_variant"):
            _lib.check(_lib.load().drin_debug_option(name.encode(), C.c_int32(-1)), "drin_debug_option")


@pytest.mark.parametrize("dataset,B,cands,layers,mask,kw", [
    ("wikidiverse", 9, 10, 2, (1, 1, 1, 1), {}), ("wikidiverse", 5, 10, 3, (1, 0, 1, 1), {}),
    ("wikidiverse", 4, 6, 4, (1, 1, 0.5, 1), {}),
    ("wikimel", 3, 5, 2, (1, 1, 1, 1), dict(entity_tokens=16, mention_tokens=32))])
def test_vector_edge_forward_stage_by_stage(dataset, B, cands, layers, mask, kw):
    """gcn_edge_feature="vector" (drin/model.py:112-116,133,139-152): per layer the activated vertices entering it
    (xa), W_v v (fv), the PRE-sigmoid edge outputs q = W_m(cat[fu, fv] + e) + b_m and the W_h outputs, then scores."""
    import torch.nn.functional as F

    from drin_b200.synthetic import spread_weights
    cfg = O.DrinConfig(num_candidates_model=cands + 1, num_gcn_layers=layers, gcn_edge_feature="vector",
                       gcn_edge_enabled=mask)
    batch = make_batch(dataset, B, 41, cands, **kw)
    sd = spread_weights(O.init_state(cfg, 0))
    V = O.vertex_encode(sd, batch[:-1])
    tt, ii = O.edge_encode(batch[:-1])
    El = [e.unsqueeze(-1).expand(-1, -1, 768) for e in (tt, batch[13] / 100, batch[12] / 100, ii)]
    eng = E.Engine(layers, mask, vector_edges=True)
    scores, ctx = eng.forward([t.cuda() for t in batch[:-1]], {k: v.cuda() for k, v in sd.items()}, training=False)
    Vl = V
    for l in range(layers):
        k = O.gcn_keys(l)
        xa = torch.cat([Vl[0], Vl[1], Vl[2].flatten(0, 1), Vl[3].flatten(0, 1)])
        assert rel_err(eng.debug_buffer(ctx, "xa", l).cpu(), xa) < 2e-5, l
        Vn, En = O.gcn_layer(sd, l, cfg, Vl, El)
        if l < layers - 1:
            fv = torch.cat([F.linear(Vl[2], sd[k["w_v"]], sd[k["b_v"]]).flatten(0, 1),
                            F.linear(Vl[3], sd[k["w_v"]], sd[k["b_v"]]).flatten(0, 1)])
            fu = torch.cat([F.linear(Vl[0], sd[k["w_u"]], sd[k["b_u"]]), F.linear(Vl[1], sd[k["w_u"]], sd[k["b_u"]])])
            Cn = cands + 1
            if l == 0:
                # first layer: edge outputs in affine form  q_k = A_u + Bv_v + e_k (W_m 1)  (csrc/gcn_vec.cu header)
                w_m, b_m = sd[k["w_m"]], sd[k["b_m"]]
                A = eng.debug_buffer(ctx, "edge_a", l).cpu()
                Bv = eng.debug_buffer(ctx, "edge_bv", l).cpu()
                w1 = eng.debug_buffer(ctx, "edge_w1", l).cpu()[0]
                assert rel_err(A, fu @ w_m[:, :384].t() + b_m) < 2e-5
                assert rel_err(Bv, fv @ w_m[:, 384:].t()) < 2e-5
                assert rel_err(w1, w_m.sum(1)) < 1e-6
                A, Bv = A.view(2, B, 1, 768), Bv.view(2, B, Cn, 768)
                e_in = torch.stack([e[..., 0] * m for e, m in zip(El, mask)])             # masked scalar input edges
                got_e = torch.stack([torch.sigmoid(A[kk >> 1] + Bv[kk & 1] + e_in[kk].unsqueeze(-1) * w1)
                                     for kk in range(4)])
            else:
                assert rel_err(eng.debug_buffer(ctx, "fv", l).cpu(), fv) < 2e-5, l
                assert rel_err(eng.debug_buffer(ctx, "fu", l).cpu(), fu) < 2e-5, l
                q = eng.debug_buffer(ctx, "q", l).cpu()
                got_e = torch.sigmoid(q).view(B, Cn, 4, 768).permute(2, 0, 1, 3)   # rows are candidate-major (r * 4 + k)
            assert rel_err(got_e, torch.stack(En)) < 2e-5, l
        h = eng.debug_buffer(ctx, "h", l).cpu()
        act = F.gelu(F.layer_norm(h, (h.shape[-1],), sd[k["ln_w"]], sd[k["ln_b"]], 1e-5))
        if l < layers - 1:
            want = torch.cat([Vn[0], Vn[1], Vn[2].flatten(0, 1), Vn[3].flatten(0, 1)])
        else:
            want = torch.cat([Vn[0], Vn[2].flatten(0, 1)])
        assert rel_err(act, want) < 2e-5, l
        Vl, El = Vn, En
    assert rel_err(scores.cpu(), O.cosine(Vl[0].unsqueeze(1), Vl[2])) < 1e-5
