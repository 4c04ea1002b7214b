"""CPU check (float64 torch, no GPU) of the algebra behind the "fold W_h of the first layer into the projections" step
listed in DESIGN.md section 9 -- planning aid for the next round, not product code.

First GCN layer, scalar edges (drin/model.py:121-153), candidate rows r = (b, c):
    et = epool W_et^T + b_et,  ei = eimg W_ei^T + b_ei                       (projections, model.py:26-46)
    z_et = et + e0 mt[b] + e2 mi[b],   z_ei = ei + e1 mt[b] + e3 mi[b]       (messages to the candidates)
    h = z W_h^T + b_h
The current backward computes dz = dh W_h for all 2BC candidate rows (2u per candidate) and
    dx_et = dz_et + (e0/C) dz_mt[b] + (e2/C) dz_mi[b] + dd0 g_mt[b] + dd2 g_mi[b]      (dd = dsigma / D, edge update)
then dW_et = dx_et^T epool (u), db_et = sum_r dx_et, dxm = dz_m + sum_c (e0 dz_et + e1 dz_ei), dg = sum_c dd x.
Every one of these is linear in dz_cand, so the [2BC, D] x [D, D] data-gradient GEMM can be replaced by pooled
per-mention sums and [D, D] x [D, D] products:
    dW_et = W_h^T (dh_et^T epool) + dzm_mt^T P0 + dzm_mi^T P2 + g_mt^T Q0 + g_mi^T Q2
    db_et = (sum_r dh_et) W_h + sum_b (dzm_mt s_ec0 + dzm_mi s_ec2 + g_mt s_dd0 + g_mi s_dd2)
    dxm_mt = dzm_mt + (sum_c e0 dh_et + e1 dh_ei) W_h
    dg_mt = (sum_c dd0 epool_c) W_et^T + b_et sum_c dd0 + (sum_c dd1 eimg_c) W_ei^T + b_ei sum_c dd1
with P_k[b] = sum_c (e_k/C) epool_c, Q_k[b] = sum_c dd_k epool_c (and the same on eimg for the image kind).
This script verifies those identities against the direct formulas on random data."""
import torch

torch.manual_seed(0)
dt = torch.float64
B, C, D, R = 5, 7, 48, 80
epool, eimg = torch.randn(B, C, D, dtype=dt), torch.randn(B, C, R, dtype=dt)
W_et, b_et = torch.randn(D, D, dtype=dt), torch.randn(D, dtype=dt)
W_ei, b_ei = torch.randn(D, R, dtype=dt), torch.randn(D, dtype=dt)
W_h = torch.randn(D, D, dtype=dt)
e = torch.rand(4, B, C, dtype=dt)                       # masked input edges tt, ti, it, ii
dd = torch.randn(4, B, C, dtype=dt) / D                 # dsigma / D of the edge update
g_mt, g_mi = torch.randn(B, D, dtype=dt), torch.randn(B, D, dtype=dt)
dh_et, dh_ei = torch.randn(B, C, D, dtype=dt), torch.randn(B, C, D, dtype=dt)
dh_mt, dh_mi = torch.randn(B, D, dtype=dt), torch.randn(B, D, dtype=dt)
et, ei = epool @ W_et.t() + b_et, eimg @ W_ei.t() + b_ei

# ---- direct (what the kernels do today) ----
dz_et, dz_ei, dzm_mt, dzm_mi = dh_et @ W_h, dh_ei @ W_h, dh_mt @ W_h, dh_mi @ W_h
dx_et = (dz_et + (e[0] / C).unsqueeze(-1) * dzm_mt.unsqueeze(1) + (e[2] / C).unsqueeze(-1) * dzm_mi.unsqueeze(1)
         + dd[0].unsqueeze(-1) * g_mt.unsqueeze(1) + dd[2].unsqueeze(-1) * g_mi.unsqueeze(1))
dx_ei = (dz_ei + (e[1] / C).unsqueeze(-1) * dzm_mt.unsqueeze(1) + (e[3] / C).unsqueeze(-1) * dzm_mi.unsqueeze(1)
         + dd[1].unsqueeze(-1) * g_mt.unsqueeze(1) + dd[3].unsqueeze(-1) * g_mi.unsqueeze(1))
dW_et = torch.einsum("bcd,bck->dk", dx_et, epool)
dW_ei = torch.einsum("bcd,bck->dk", dx_ei, eimg)
db_et, db_ei = dx_et.sum((0, 1)), dx_ei.sum((0, 1))
dxm_mt = dzm_mt + (e[0].unsqueeze(-1) * dz_et + e[1].unsqueeze(-1) * dz_ei).sum(1)
dxm_mi = dzm_mi + (e[2].unsqueeze(-1) * dz_et + e[3].unsqueeze(-1) * dz_ei).sum(1)
dg_mt = (dd[0].unsqueeze(-1) * et + dd[1].unsqueeze(-1) * ei).sum(1)
dg_mi = (dd[2].unsqueeze(-1) * et + dd[3].unsqueeze(-1) * ei).sum(1)

# ---- folded: no [BC, D] x [D, D] data-gradient product ----
pool = lambda w, x: (w.unsqueeze(-1) * x).sum(1)                        # [B, C] x [B, C, K] -> [B, K]
G_et = torch.einsum("bcd,bck->dk", dh_et, epool)                        # the weight-gradient-sized GEMM that remains
G_ei = torch.einsum("bcd,bck->dk", dh_ei, eimg)
f_dW_et = (W_h.t() @ G_et + dzm_mt.t() @ pool(e[0] / C, epool) + dzm_mi.t() @ pool(e[2] / C, epool)
           + g_mt.t() @ pool(dd[0], epool) + g_mi.t() @ pool(dd[2], epool))
f_dW_ei = (W_h.t() @ G_ei + dzm_mt.t() @ pool(e[1] / C, eimg) + dzm_mi.t() @ pool(e[3] / C, eimg)
           + g_mt.t() @ pool(dd[1], eimg) + g_mi.t() @ pool(dd[3], eimg))
rowsum = lambda w: w.sum(1, keepdim=True)
f_db_et = (dh_et.sum((0, 1)) @ W_h + (dzm_mt * rowsum(e[0] / C) + dzm_mi * rowsum(e[2] / C) + g_mt * rowsum(dd[0])
                                      + g_mi * rowsum(dd[2])).sum(0))
f_db_ei = (dh_ei.sum((0, 1)) @ W_h + (dzm_mt * rowsum(e[1] / C) + dzm_mi * rowsum(e[3] / C) + g_mt * rowsum(dd[1])
                                      + g_mi * rowsum(dd[3])).sum(0))
f_dxm_mt = dzm_mt + (pool(e[0], dh_et) + pool(e[1], dh_ei)) @ W_h
f_dxm_mi = dzm_mi + (pool(e[2], dh_et) + pool(e[3], dh_ei)) @ W_h
f_dg_mt = (pool(dd[0], epool) @ W_et.t() + b_et * rowsum(dd[0]) + pool(dd[1], eimg) @ W_ei.t() + b_ei * rowsum(dd[1]))
f_dg_mi = (pool(dd[2], epool) @ W_et.t() + b_et * rowsum(dd[2]) + pool(dd[3], eimg) @ W_ei.t() + b_ei * rowsum(dd[3]))

# ---- forward fold: h_et without materialising z_et ----
mt, mi, b_h = torch.randn(B, D, dtype=dt), torch.randn(B, D, dtype=dt), torch.randn(D, dtype=dt)
h_et = (et + e[0].unsqueeze(-1) * mt.unsqueeze(1) + e[2].unsqueeze(-1) * mi.unsqueeze(1)) @ W_h.t() + b_h
f_h_et = (epool @ (W_h @ W_et).t() + e[0].unsqueeze(-1) * (mt @ W_h.t()).unsqueeze(1)
          + e[2].unsqueeze(-1) * (mi @ W_h.t()).unsqueeze(1) + (W_h @ b_et + b_h))

for name, a, b in (("dW_et", dW_et, f_dW_et), ("dW_ei", dW_ei, f_dW_ei), ("db_et", db_et, f_db_et),
                   ("db_ei", db_ei, f_db_ei), ("dxm_mt", dxm_mt, f_dxm_mt), ("dxm_mi", dxm_mi, f_dxm_mi),
                   ("dg_mt", dg_mt, f_dg_mt), ("dg_mi", dg_mi, f_dg_mi), ("h_et", h_et, f_h_et)):
    err = float((a - b).abs().max() / a.abs().max())
    print(f"{name:8s} rel err {err:.2e}")
    assert err < 1e-12, name
print("fold identities hold")
