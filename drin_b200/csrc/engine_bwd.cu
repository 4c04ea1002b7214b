// Backward orchestration of the DRIN hot path (what loss.backward() does after reference train.py:34).
// Mirrors engine.cu: per layer one data-gradient GEMM (dZ = dH W_h), one split-K weight-gradient GEMM
// (dW_h = dH^T Z) and the fused memory-bound kernels of gcn_bwd.cu; the small mention-side GEMMs carry
// the W_u / W_v gradients of the dynamic edge update.
#include "engine.cuh"

namespace drin {

static Operand op(const Planes& p, long long rows, int cols, long long row_offset = 0) {
  Operand o;
  o.hi = p.hi + row_offset * cols;
  o.lo = p.lo ? p.lo + row_offset * cols : nullptr;
  o.rows = rows;
  o.cols = cols;
  o.ld = cols;
  return o;
}

// Deferred reductions of one backward pass: split-K slices and per-CTA column sums are written to private regions
// of two arenas and reduced by two launches at the very end (instead of 8 + 5 small launches in between).
static int g_defer_reductions = 1;       // test / A-B hook: 0 runs every reduction right after its producer
void debug_set_defer_reductions(int v) { g_defer_reductions = v < 0 ? 1 : v; }

struct Deferred {
  SplitKJobs splitk;
  ColsumJobs colsum;
  cudaStream_t stream;
  int D;
  // column sums: queue (default) or reduce immediately
  int add_colsum(const float* src0, int ctas0, const float* src1, int ctas1, int nvec, float* out0, float* out1, float* out2) {
    if (!g_defer_reductions) return colsum_reduce(stream, src0, ctas0, src1, ctas1, nvec, D, out0, out1, out2);
    return colsum.add(src0, ctas0, src1, ctas1, nvec, out0, out1, out2) ? fail(DRIN_ERR_ARG, "internal: too many column-sum jobs") : DRIN_OK;
  }
  float* partial_next;
  float* partial_end;
  float* colsum_next;
  float* colsum_end;
  float* take_colsum(size_t n) {
    float* p = colsum_next;
    colsum_next += n;
    return colsum_next <= colsum_end ? p : nullptr;
  }
};

// C[M,N] = A[K,M]^T B[K,N], contraction split so that tiles * slices ~ one wave
static int weight_grad(cudaStream_t s, const Workspace& ws, Deferred& df, const Operand& A, const Operand& B, int M, int N,
                       long long K, float* out) {
  if (!out) return fail(DRIN_ERR_ARG, "gradient buffer is null");
  GemmEpilogue ep;
  ep.C = out;
  ep.ldc = N;
  const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
  int ksplit = 148 / tiles;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > ws.ksplit) ksplit = ws.ksplit;
  float* region = df.partial_next;
  df.partial_next += (size_t)ksplit * M * N;
  if (df.partial_next > df.partial_end) return fail(DRIN_ERR_WORKSPACE, "internal: split-K arena exhausted");
  return gemm_tcgen05(s, GEMM_TN, A, B, M, N, K, ep, ksplit, region, g_defer_reductions ? &df.splitk : nullptr);
}

// W_m[:, :H] / W_m[:, H:] as the [K = D, N = H] operand of an NN GEMM (dA W_m[:, :H], dBv W_m[:, H:])
static Operand wm_half_kn(const Planes& w_m, int D, int half) {
  Operand o;
  const int H = D / 2;
  o.hi = w_m.hi + half * H;
  o.lo = w_m.lo ? w_m.lo + half * H : nullptr;
  o.rows = D;
  o.cols = H;
  o.ld = D;
  return o;
}

// Backward of the vector-edge layers (forward_vector_layers in engine.cu), layer by layer from the scores down:
//   dZ = dH W_h, dW_h = dH^T Z;  [edge update] dM = dQ W_m, dW_m = dQ^T M;  vec_layer_bwd (messages, sigmoid, cat);
//   [edge update] dW_u = dFu^T Xm, dW_v = dFv^T Xv, dX += dFu W_u | dFv W_v;  vec_rows_bwd (GELU + LayerNorm).
static int backward_vector_layers(const drin_config& c, const drin_params& p, Workspace& ws, const float* dscores,
                                  const drin_params& grads, Deferred& df, cudaStream_t stream) {
  const long long B = c.batch, C = c.candidates, BC = B * C;
  const int D = c.embed_dim, H = D / 2, L = c.gcn_layers;
  const size_t part_floats = (size_t)ws.colsum_ctas * 3 * D;
  StridedColsumJobs cs;          // bias / LayerNorm gradients: per-CTA partials of every layer, reduced by one launch
  auto colsum = [&](const float* src, int count, long long stride, int n, float* out) {
    if (!out) return fail(DRIN_ERR_ARG, "gradient buffer is null");
    return cs.add(src, count, stride, n, out) ? fail(DRIN_ERR_ARG, "internal: too many column-sum jobs") : (int)DRIN_OK;
  };
  for (int l = L - 1; l >= 0; --l) {
    const LayerWs& lw = ws.layer[l];
    const drin_layer_params& lg = grads.layer[l];
    float* vec_part = ws.vec_part + (size_t)l * vec_layer_ctas() * 3 * D;
    float* rows_part = ws.rows_part + (size_t)l * vec_rows_ctas() * 4 * D;
    if (l == L - 1) {
      const drin_layer_params& lp = p.layer[l];
      ScoreBwdArgs sa{};
      sa.B = c.batch; sa.C = c.candidates; sa.D = D;
      sa.h_mt = lw.h; sa.h_et = lw.h + B * D; sa.gamma = lp.ln_w; sa.beta = lp.ln_b; sa.dscores = dscores;
      float* part = df.take_colsum(part_floats);
      float* part2 = ws.slices > 1 ? df.take_colsum(part_floats) : nullptr;
      if (!part || (ws.slices > 1 && !part2)) return fail(DRIN_ERR_WORKSPACE, "internal: column-sum arena exhausted");
      sa.dh_hi = ws.dh.hi; sa.dh_lo = ws.dh.lo; sa.partials = part;
      sa.slices = ws.slices; sa.slice_part = ws.slice_part; sa.partials2 = part2;
      bool used2 = false;
      DRIN_TRY(score_bwd(stream, sa, &used2));
      DRIN_TRY(df.add_colsum(part, backward_ctas(), used2 ? part2 : nullptr, used2 ? backward_ctas() : 0, 3, lg.ln_w, lg.ln_b, lg.b_h));
    }
    {
      GemmEpilogue ez;
      ez.C = ws.dz; ez.ldc = D;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dh, lw.rows, D), op(lw.w_h, D, D), lw.rows, D, D, ez));
      DRIN_TRY(weight_grad(stream, ws, df, op(ws.dh, lw.rows, D), op(lw.z, lw.rows, D), D, D, lw.rows, lg.w_h));
    }
    const bool general = lw.dyn && !lw.affine;
    if (general) {     // dq_p holds the gradient w.r.t. this layer's pre-sigmoid edge outputs (written by layer l + 1)
      GemmEpilogue em;
      em.C = ws.dm; em.ldc = D;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dq_p, 4 * BC, D), op(lw.w_m, D, D), 4 * BC, D, D, em));
      DRIN_TRY(weight_grad(stream, ws, df, op(ws.dq_p, 4 * BC, D), op(lw.m_p, 4 * BC, D), D, D, 4 * BC, lg.w_m));
    }
    VecLayerArgs va{};
    va.B = c.batch; va.C = c.candidates; va.D = D; va.full = lw.full; va.dyn = general;
    for (int k = 0; k < 4; ++k) va.en[k] = c.edge_enabled[k];
    va.xa = lw.xa;
    if (l == 0) {
      va.e_scalar = ws.edges0;
    } else if (ws.layer[l - 1].affine) {       // the layer below is the first layer: its edge outputs are A_u + Bv_v + e w1
      const LayerWs& pw = ws.layer[l - 1];
      va.e_scalar = ws.edges0; va.edge_a = pw.edge_a; va.edge_bv = pw.edge_bv; va.edge_w1 = pw.edge_w1;
      va.da_hi = ws.da_p.hi; va.da_lo = ws.da_p.lo;
      va.dbv_hi = ws.dbv_p.hi; va.dbv_lo = ws.dbv_p.lo;
    } else {
      va.q_in = ws.layer[l - 1].q;
      va.dq_hi = ws.dq_p.hi; va.dq_lo = ws.dq_p.lo;      // safe: dm / dW_m above already consumed dq_p
    }
    va.dz = ws.dz;
    va.dxa = ws.dxa;
    if (general) {
      va.dm = ws.dm;
      va.dfu_hi = ws.dfu_p.hi; va.dfu_lo = ws.dfu_p.lo;
      va.dfv_hi = ws.dfv_p.hi; va.dfv_lo = ws.dfv_p.lo;
    }
    va.partials = vec_part;
    int prow = 0;
    DRIN_TRY(vec_layer_bwd(stream, va, &prow));
    if (l > 0) {
      DRIN_TRY(colsum(vec_part, prow, 3 * D, D, grads.layer[l - 1].b_m));
      if (ws.layer[l - 1].affine) DRIN_TRY(colsum(vec_part + 2 * D, prow, 3 * D, D, ws.dw1));
    }
    if (general) {
      DRIN_TRY(colsum(vec_part + D, prow, 3 * D, H, lg.b_u));
      DRIN_TRY(colsum(vec_part + D + H, prow, 3 * D, H, lg.b_v));
    }
    if (lw.affine) {
      // first layer: dA, dBv (written by the layer above) -> dFu, dFv and the two halves of dW_m; no per-edge-type GEMM
      GemmEpilogue ef;
      ef.ld_planes = H;
      ef.out_hi = ws.dfu_p.hi; ef.out_lo = ws.dfu_p.lo;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.da_p, 2 * B, D), wm_half_kn(lw.w_m, D, 0), 2 * B, H, D, ef));
      ef.out_hi = ws.dfv_p.hi; ef.out_lo = ws.dfv_p.lo;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dbv_p, 2 * BC, D), wm_half_kn(lw.w_m, D, 1), 2 * BC, H, D, ef));
      DRIN_TRY(weight_grad(stream, ws, df, op(ws.da_p, 2 * B, D), op(lw.fu_p, 2 * B, H), D, H, 2 * B, ws.dwm_a));
      DRIN_TRY(weight_grad(stream, ws, df, op(ws.dbv_p, 2 * BC, D), op(lw.fv_p, 2 * BC, H), D, H, 2 * BC, ws.dwm_b));
    }
    if (lw.dyn) {
      DRIN_TRY(weight_grad(stream, ws, df, op(ws.dfu_p, 2 * B, H), op(lw.xa_p, 2 * B, D), H, D, 2 * B, lg.w_u));
      DRIN_TRY(weight_grad(stream, ws, df, op(ws.dfv_p, 2 * BC, H), op(lw.xa_p, 2 * BC, D, 2 * B), H, D, 2 * BC, lg.w_v));
      GemmEpilogue ex;
      ex.C = ws.dxuv; ex.ldc = D;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dfu_p, 2 * B, H), op(lw.w_u, H, D), 2 * B, D, H, ex));
      ex.C = ws.dxuv + 2 * B * D;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dfv_p, 2 * BC, H), op(lw.w_v, H, D), 2 * BC, D, H, ex));
    }
    VecRowsBwdArgs ra{};
    ra.B = B; ra.BC = BC; ra.D = D;
    ra.d0 = ws.dxa;
    ra.d1 = lw.dyn ? ws.dxuv : nullptr;
    ra.partials = rows_part;
    const long long pstride = 4 * D;
    if (l > 0) {
      const drin_layer_params& pg = grads.layer[l - 1];
      ra.h_prev = ws.layer[l - 1].h;
      ra.ln_gamma = p.layer[l - 1].ln_w; ra.ln_beta = p.layer[l - 1].ln_b;
      ra.out_hi = ws.dh.hi; ra.out_lo = ws.dh.lo;
      DRIN_TRY(vec_rows_bwd(stream, ra));
      DRIN_TRY(colsum(rows_part, vec_rows_ctas(), pstride, D, pg.ln_w));
      DRIN_TRY(colsum(rows_part + D, vec_rows_ctas(), pstride, D, pg.ln_b));
      DRIN_TRY(colsum(rows_part + 2 * D, vec_rows_ctas(), pstride, D, pg.b_h));
    } else {
      ra.out_hi = ws.dx0.hi; ra.out_lo = ws.dx0.lo;
      DRIN_TRY(vec_rows_bwd(stream, ra));
      DRIN_TRY(colsum(rows_part, vec_rows_ctas(), pstride, D, grads.b_mt));
      DRIN_TRY(colsum(rows_part + D, vec_rows_ctas(), pstride, D, grads.b_mi));
      DRIN_TRY(colsum(rows_part + 2 * D, vec_rows_ctas(), pstride, D, grads.b_et));
      DRIN_TRY(colsum(rows_part + 3 * D, vec_rows_ctas(), pstride, D, grads.b_ei));
    }
  }
  DRIN_TRY(strided_colsum_multi(stream, cs));
  if (ws.layer[0].affine) {
    // dW_m, b_u, b_v of the first layer need the finished split-K sums and column sums: reduce what is queued, then assemble
    DRIN_TRY(splitk_reduce_multi(stream, df.splitk));
    df.splitk.count = 0;
    const drin_layer_params& g0 = grads.layer[0];
    DRIN_TRY(wm_fixup(stream, ws.dwm_a, ws.dwm_b, ws.dw1, g0.b_m, p.layer[0].w_m, D, g0.w_m, g0.b_u, g0.b_v));
  }
  return DRIN_OK;
}

int backward(const drin_config& c, const drin_inputs& in, const drin_params& p, void* workspace,
             size_t workspace_bytes, const float* dscores, const drin_params& grads, cudaStream_t stream,
             cudaEvent_t layers_done) {
  if (!c.training) return fail(DRIN_ERR_ARG, "drin_backward needs a config with training = 1");
  Workspace ws;
  DRIN_TRY(plan_workspace(c, &in, workspace, ws));
  if (!workspace || workspace_bytes < ws.bytes)
    return fail(DRIN_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, ws.bytes);
  if (!dscores) return fail(DRIN_ERR_ARG, "dscores is null");
  if (ws.colsum_ctas != backward_ctas()) return fail(DRIN_ERR_ARG, "internal: partial-sum grid mismatch");
  const long long B = c.batch, C = c.candidates, BC = B * C;
  const int D = c.embed_dim, R = c.resnet_dim, L = c.gcn_layers;
  SideStream side(stream);
  Deferred df;
  df.stream = stream; df.D = D;
  df.partial_next = ws.partial; df.partial_end = ws.partial + ws.partial_floats;
  df.colsum_next = ws.colsum; df.colsum_end = ws.colsum + ws.colsum_floats;
  const size_t part_floats = (size_t)ws.colsum_ctas * 3 * D;
  float* dedges[2] = {ws.dedges, ws.dedges + 4 * BC};

  if (ws.vec) DRIN_TRY(backward_vector_layers(c, p, ws, dscores, grads, df, stream));
  for (int l = ws.vec ? -1 : L - 1; l >= 0; --l) {
    const LayerWs& lw = ws.layer[l];
    const drin_layer_params& lp = p.layer[l];
    const drin_layer_params& lg = grads.layer[l];
    if (l == L - 1) {
      ScoreBwdArgs sa{};
      sa.B = c.batch; sa.C = c.candidates; sa.D = D;
      sa.h_mt = lw.h; sa.h_et = lw.h + B * D; sa.gamma = lp.ln_w; sa.beta = lp.ln_b; sa.dscores = dscores;
      float* part = df.take_colsum(part_floats);
      if (!part) return fail(DRIN_ERR_WORKSPACE, "internal: column-sum arena exhausted");
      sa.dh_hi = ws.dh.hi; sa.dh_lo = ws.dh.lo; sa.partials = part;
      float* part2 = nullptr;
      if (ws.slices > 1) {
        part2 = df.take_colsum(part_floats);
        if (!part2) return fail(DRIN_ERR_WORKSPACE, "internal: column-sum arena exhausted");
      }
      sa.slices = ws.slices; sa.slice_part = ws.slice_part; sa.partials2 = part2;
      bool used2 = false;
      DRIN_TRY(score_bwd(stream, sa, &used2));
      DRIN_TRY(df.add_colsum(part, backward_ctas(), used2 ? part2 : nullptr, used2 ? backward_ctas() : 0, 3, lg.ln_w, lg.ln_b, lg.b_h));
    }
    // dZ = dH W_h ; dW_h = dH^T Z
    {
      // the two consumers of dH are independent: dW_h runs on the side stream next to dZ; joined before the layer
      // kernel overwrites the dh planes with the gradient of the layer below
      GemmEpilogue ez;
      ez.C = ws.dz; ez.ldc = D;
      cudaStream_t ss = side.fork();
      DRIN_TRY(weight_grad(ss, ws, df, op(ws.dh, lw.rows, D), op(lw.z, lw.rows, D), D, D, lw.rows, lg.w_h));
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dh, lw.rows, D), op(lw.w_h, D, D), lw.rows, D, D, ez));
      DRIN_TRY(side.join());
    }
    LayerBwdArgs la{};
    la.B = c.batch; la.C = c.candidates; la.D = D; la.full = lw.full;
    for (int k = 0; k < 4; ++k) la.en[k] = c.edge_enabled[k];
    la.xm = lw.xm;
    if (l == 0) {
      la.x_et = ws.x0 + 2 * B * D;
      la.x_ei = ws.x0 + (2 * B + BC) * D;
      la.edges_in = ws.edges0;
      la.dcand_hi = ws.dx0.hi; la.dcand_lo = ws.dx0.lo;
      la.dedges_in = nullptr;
    } else {
      const LayerWs& pw = ws.layer[l - 1];
      la.x_et = pw.h + 2 * B * D;
      la.x_ei = pw.h + (2 * B + BC) * D;
      la.ln_gamma = p.layer[l - 1].ln_w;
      la.ln_beta = p.layer[l - 1].ln_b;
      la.dcand_hi = ws.dh.hi; la.dcand_lo = ws.dh.lo;
      if (pw.dyn) {
        la.edges_in = pw.edges_out;
        la.dedges_in = dedges[l % 2];
      } else {   // static edges are inputs: same masked view as the forward pass, no edge gradient
        la.edges_in = ws.edges0;
        for (int k = 0; k < 4; ++k) la.en[k] = powf(c.edge_enabled[k], (float)(l + 1));
        la.dedges_in = nullptr;
      }
    }
    la.dz = ws.dz;
    if (lw.dyn) {
      la.g = lw.g;
      la.edges_out = lw.edges_out;
      la.dedges_out = dedges[(l + 1) % 2];
      la.dg_hi = ws.dg_p.hi; la.dg_lo = ws.dg_p.lo;
      la.dbeta = ws.dbeta;
    }
    la.dxm = ws.dxm;
    float* partA = df.take_colsum(part_floats);
    float* partB = df.take_colsum(part_floats);
    float* partC = df.take_colsum(part_floats);
    if (!partA || !partB || !partC) return fail(DRIN_ERR_WORKSPACE, "internal: column-sum arena exhausted");
    la.partials = partA;
    la.slices = ws.slices; la.slice_part = ws.slice_part; la.slice_dbeta = ws.slice_dbeta;
    int rowsA = layer_bwd_ctas();
    la.partial_rows = &rowsA;
    DRIN_TRY(gcn_layer_bwd(stream, la));

    if (lw.dyn) {
      // edge-update weights: fu = W_u xm + b_u, g = fu W_v, beta = fu . b_v
      // four 2B-row GEMMs, each a third of the machine: the two weight gradients go to the side stream (dW_v needs
      // only dg, dW_u the finished dfu planes), the data-gradient chain dfu -> dxu stays on the main stream
      GemmEpilogue ef;
      ef.C = ws.dfu; ef.ldc = D;
      cudaStream_t ss = side.fork();
      DRIN_TRY(weight_grad(ss, ws, df, op(lw.fu_p, 2 * B, D), op(ws.dg_p, 2 * B, D), D, D, 2 * B, lg.w_v));
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(ws.dg_p, 2 * B, D), op(lw.w_v, D, D), 2 * B, D, D, ef));
      DRIN_TRY(dfu_finish(stream, D, ws.dfu, ws.dbeta, lp.b_v, lw.fu, 2 * B, ws.dfu_p.hi, ws.dfu_p.lo, partC));
      DRIN_TRY(df.add_colsum(partC, backward_ctas(), nullptr, 0, 2, lg.b_u, lg.b_v, nullptr));
      ss = side.fork();                                    // the side stream now also waits for the dfu planes
      DRIN_TRY(weight_grad(ss, ws, df, op(ws.dfu_p, 2 * B, D), op(lw.xm_p, 2 * B, D), D, D, 2 * B, lg.w_u));
      GemmEpilogue ex;
      ex.C = ws.dxu; ex.ldc = D;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NN, op(ws.dfu_p, 2 * B, D), op(lw.w_u, D, D), 2 * B, D, D, ex));
    }
    MentionBwdArgs ma{};
    ma.B = c.batch; ma.D = D;
    ma.dxm = ws.dxm;
    ma.dxu = lw.dyn ? ws.dxu : nullptr;
    if (l > 0) {
      ma.h_prev = ws.layer[l - 1].h;
      ma.ln_gamma = p.layer[l - 1].ln_w;
      ma.ln_beta = p.layer[l - 1].ln_b;
      ma.out_hi = ws.dh.hi; ma.out_lo = ws.dh.lo;
    } else {
      ma.out_hi = ws.dx0.hi; ma.out_lo = ws.dx0.lo;
    }
    ma.partials = partB;
    DRIN_TRY(mention_bwd_finish(stream, ma));
    DRIN_TRY(side.join());                                 // dg / dfu planes are rewritten by the layer below
    if (l > 0) {
      const drin_layer_params& pg = grads.layer[l - 1];
      DRIN_TRY(df.add_colsum(partA, rowsA, partB, backward_ctas(), 3, pg.ln_w, pg.ln_b, pg.b_h));
    } else {
      DRIN_TRY(df.add_colsum(partA, rowsA, nullptr, 0, 3, grads.b_et, grads.b_ei, nullptr));
      DRIN_TRY(df.add_colsum(partB, backward_ctas(), nullptr, 0, 3, grads.b_mt, grads.b_mi, nullptr));
    }
  }

  if (layers_done) {
    // a data-parallel caller wants to start reducing the layer gradients now: finish them (two launches) and signal
    DRIN_TRY(splitk_reduce_multi(stream, df.splitk));
    DRIN_TRY(colsum_reduce_multi(stream, df.colsum, D));
    df.splitk.count = 0;
    df.colsum.count = 0;
    DRIN_CUDA(cudaEventRecord(layers_done, stream));
  }

  // input projections: dW = dX0^T A (no data gradient: the cached features are constants); the two mention-side ones
  // (K = B rows) run beside the two candidate-side ones
  {
    cudaStream_t ss = side.fork();
    DRIN_TRY(weight_grad(ss, ws, df, op(ws.dx0, B, D, 0), op(ws.span, B, D), D, D, B, grads.w_mt));
    DRIN_TRY(weight_grad(ss, ws, df, op(ws.dx0, B, D, B), op(ws.mim, B, R), D, R, B, grads.w_mi));
    DRIN_TRY(weight_grad(stream, ws, df, op(ws.dx0, BC, D, 2 * B), op(ws.epool, BC, D), D, D, BC, grads.w_et));
    DRIN_TRY(weight_grad(stream, ws, df, op(ws.dx0, BC, D, 2 * B + BC), op(ws.eimg, BC, R), D, R, BC, grads.w_ei));
    DRIN_TRY(side.join());
  }
  // every split-K slice and every per-CTA column partial of the pass is reduced here, in two launches
  DRIN_TRY(splitk_reduce_multi(stream, df.splitk));
  DRIN_TRY(colsum_reduce_multi(stream, df.colsum, D));
  return DRIN_OK;
}

}  // namespace drin
