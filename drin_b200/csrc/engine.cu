// Workspace plan + forward orchestration of the DRIN hot path (reference drin/model.py:164-209).
#include "engine.cuh"

#include <vector>

namespace drin {

int check_config(const drin_config& c) {
  if (c.batch <= 0 || c.candidates <= 1) return fail(DRIN_ERR_ARG, "config: batch=%d candidates=%d", c.batch, c.candidates);
  if (c.embed_dim != 768) return fail(DRIN_ERR_ARG, "config: gcn_embed_dim=%d (kernels are built for 768)", c.embed_dim);
  if (c.resnet_dim <= 0 || c.resnet_dim % 256) return fail(DRIN_ERR_ARG, "config: resnet_dim=%d must be a multiple of 256", c.resnet_dim);
  if (c.gcn_layers < 1 || c.gcn_layers > DRIN_MAX_LAYERS) return fail(DRIN_ERR_ARG, "config: gcn_layers=%d (1..%d)", c.gcn_layers, DRIN_MAX_LAYERS);
  if (c.mention_tokens <= 0 || c.regions <= 0 || c.entity_tokens < 0) return fail(DRIN_ERR_ARG, "config: bad token/region counts");
  if (c.mention_objects < 1 || c.mention_objects > 4 || c.entity_objects < 1) return fail(DRIN_ERR_ARG, "config: mention_objects=%d entity_objects=%d", c.mention_objects, c.entity_objects);
  if (c.precision != DRIN_FP32 && c.precision != DRIN_BF16) return fail(DRIN_ERR_ARG, "config: precision=%d", c.precision);
  return DRIN_OK;
}

// Test hook (no compute-sanitizer on the GPU pool): with a guard size set, every buffer of the workspace plan is
// followed by `guard` bytes that no kernel may touch.  A test fills the whole workspace with a poison pattern, runs a
// step and checks (a) every guard band still holds the pattern (no out-of-bounds write between buffers), (b) results
// equal those of a zero-filled workspace bit for bit (no read of memory the step did not write itself).
static size_t g_guard_bytes = 0;
static std::vector<size_t> g_guard_offsets;      // of the most recent plan (planning is host-side and cheap)
void debug_set_workspace_guard(int bytes) { g_guard_bytes = bytes > 0 ? align_up((size_t)bytes, 256) : 0; }
size_t debug_workspace_guard_bytes() { return g_guard_bytes; }
const std::vector<size_t>& debug_workspace_guard_offsets() { return g_guard_offsets; }

// ---- side stream (see engine.cuh) ----
static bool g_side_stream = true;
void debug_set_side_stream(int v) { g_side_stream = v != 0; }
namespace {
struct SideState {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  bool ok = false;
};
SideState g_side[64];
}  // namespace

SideStream::SideStream(cudaStream_t main) : main_(main) {
  if (!g_side_stream || prof::enabled()) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  SideState& st = g_side[dev];
  if (!st.ok) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;   // create outside captures
    if (cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&st.fork_ev, cudaEventDisableTiming) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&st.join_ev, cudaEventDisableTiming) != cudaSuccess) return;
    st.ok = true;
  }
  side_ = st.stream;
  fork_ev_ = st.fork_ev;
  join_ev_ = st.join_ev;
}

cudaStream_t SideStream::fork() {
  if (!side_) return main_;
  if (cudaEventRecord(fork_ev_, main_) != cudaSuccess || cudaStreamWaitEvent(side_, fork_ev_, 0) != cudaSuccess) {
    cudaGetLastError();
    return main_;                      // cannot fork: stay serial (correct, just not concurrent)
  }
  forked_ = true;
  return side_;
}

int SideStream::join() {
  if (!side_ || !forked_) return DRIN_OK;
  DRIN_CUDA(cudaEventRecord(join_ev_, side_));
  DRIN_CUDA(cudaStreamWaitEvent(main_, join_ev_, 0));
  forked_ = false;
  return DRIN_OK;
}

namespace {
struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<char*>(b)) { g_guard_offsets.clear(); }
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    if (g_guard_bytes) {
      off = align_up(off, 256);
      g_guard_offsets.push_back(off);
      off += g_guard_bytes;
    }
    return p;
  }
  Planes planes(size_t n, bool split) {
    Planes p;
    p.hi = take<bf16>(n);
    if (split) p.lo = take<bf16>(n);
    return p;
  }
};
}  // namespace

// Vector-edge layers (gcn_vec.cu): everything after x0.
static int plan_vector(const drin_config& c, Bump& m, Workspace& ws) {
  const bool split = c.precision == DRIN_FP32;
  const size_t B = c.batch, C = c.candidates, D = c.embed_dim, R = c.resnet_dim, BC = B * C, H = D / 2;
  const size_t rows_all = 2 * B + 2 * BC;
  const int L = c.gcn_layers;
  ws.x0_p = m.planes(rows_all * D, split);
  ws.xm0_p = ws.x0_p;
  for (int l = 0; l < L; ++l) {
    LayerWs& lw = ws.layer[l];
    lw.full = l < L - 1;
    lw.dyn = lw.full;
    lw.rows = lw.full ? (long long)rows_all : (long long)(B + BC);
    lw.w_h = m.planes(D * D, split);
    if (lw.dyn) {
      lw.w_u = m.planes(H * D, split);
      lw.w_v = m.planes(H * D, split);
      lw.w_m = m.planes(D * D, split);
    }
    if (l == 0) {
      lw.xa = ws.x0;
      lw.xa_p = ws.x0_p;
    } else {
      lw.xa = m.take<float>(rows_all * D);
      if (lw.dyn) lw.xa_p = m.planes(rows_all * D, split);
    }
    lw.xm = lw.xa;
    lw.affine = lw.dyn && l == 0;
    if (lw.affine) {
      lw.fu_p = m.planes(2 * B * H, split);
      lw.fv_p = m.planes(2 * BC * H, split);
      lw.edge_a = m.take<float>(2 * B * D);
      lw.edge_bv = m.take<float>(2 * BC * D);
      lw.edge_w1 = m.take<float>(D);
    } else if (lw.dyn) {
      lw.fu = m.take<float>(2 * B * H);
      lw.fv = m.take<float>(2 * BC * H);
      lw.m_p = m.planes(4 * BC * D, split);
      lw.q = m.take<float>(4 * BC * D);
    }
    lw.z = m.planes((size_t)lw.rows * D, split);
    lw.h = m.take<float>((size_t)lw.rows * D);
  }
  if (c.training) {
    ws.dh = m.planes(rows_all * D, split);
    ws.dz = m.take<float>(rows_all * D);
    ws.dxa = m.take<float>(rows_all * D);
    ws.dxuv = m.take<float>(rows_all * D);
    if (L > 2) {            // general edge updates exist only above the first layer
      ws.dm = m.take<float>(4 * BC * D);
      ws.dq_p = m.planes(4 * BC * D, split);
    }
    ws.dfu_p = m.planes(2 * B * H, split);
    ws.dfv_p = m.planes(2 * BC * H, split);
    ws.da_p = m.planes(2 * B * D, split);
    ws.dbv_p = m.planes(2 * BC * D, split);
    ws.dwm_a = m.take<float>(D * H);
    ws.dwm_b = m.take<float>(D * H);
    ws.dw1 = m.take<float>(D);
    ws.dx0 = m.planes(rows_all * D, split);
    if (ws.slices > 1) {
      ws.slice_part = m.take<float>(B * ws.slices * 4 * D);
      ws.slice_dbeta = m.take<float>(B * ws.slices * 2);
    }
    ws.ksplit = 8;
    ws.partial_floats = (size_t)(3 * L + 2) * ws.ksplit * D * D + (size_t)2 * 3 * D * R;
    ws.partial = m.take<float>(ws.partial_floats);
    ws.colsum_ctas = backward_ctas();
    ws.colsum_floats = (size_t)2 * ws.colsum_ctas * 3 * D;        // score_bwd (and its sliced mention finish)
    ws.colsum = m.take<float>(ws.colsum_floats);
    ws.vec_part = m.take<float>((size_t)L * vec_layer_ctas() * 3 * D);
    ws.rows_part = m.take<float>((size_t)L * vec_rows_ctas() * 4 * D);
  }
  ws.bytes = align_up(m.off, 256);
  return DRIN_OK;
}

int plan_workspace(const drin_config& c, const drin_inputs* in, void* base, Workspace& ws) {
  DRIN_TRY(check_config(c));
  const bool split = c.precision == DRIN_FP32;
  const size_t B = c.batch, C = c.candidates, D = c.embed_dim, R = c.resnet_dim, BC = B * C;
  const int L = c.gcn_layers;
  Bump m(base);
  ws.w_mt = m.planes(D * D, split);
  ws.w_et = m.planes(D * D, split);
  ws.w_mi = m.planes(D * R, split);
  ws.w_ei = m.planes(D * R, split);
  ws.span = m.planes(B * D, split);
  ws.mim = m.planes(B * R, split);
  if (split || c.entity_tokens > 0 || c.indexed) {
    ws.epool = m.planes(BC * D, split);
  } else {   // bf16 WikiDiverse: the entity text rows are the GEMM operand as they are
    ws.epool = Planes();
    if (in) ws.epool.hi = const_cast<bf16*>(static_cast<const bf16*>(in->entity_text_feature));
  }
  if (split || c.indexed) {
    ws.eimg = m.planes(BC * R, split);
  } else {
    ws.eimg = Planes();
    if (in) ws.eimg.hi = const_cast<bf16*>(static_cast<const bf16*>(in->entity_image_feature));
  }
  ws.edges0 = m.take<float>(4 * BC);
  ws.slices = row_kernel_slices(c.batch, c.candidates);
  if (ws.slices > 1) ws.acc_part = m.take<float>(B * ws.slices * 2 * D);
  ws.x0 = m.take<float>((2 * B + 2 * BC) * D);
  ws.vec = vector_path(c);
  if (ws.vec) return plan_vector(c, m, ws);
  ws.xm0_p = m.planes(2 * B * D, split);
  for (int l = 0; l < L; ++l) {
    LayerWs& lw = ws.layer[l];
    lw.full = l < L - 1;
    lw.dyn = lw.full && !c.static_edges;
    lw.rows = lw.full ? (long long)(2 * B + 2 * BC) : (long long)(B + BC);
    lw.w_h = m.planes(D * D, split);
    if (lw.dyn) {
      lw.w_u = m.planes(D * D, split);
      lw.w_v = m.planes(D * D, split);
    }
    if (l == 0) {
      lw.xm = ws.x0;
      lw.xm_p = ws.xm0_p;
    } else {
      lw.xm = m.take<float>(2 * B * D);
      if (lw.dyn) lw.xm_p = m.planes(2 * B * D, split);
    }
    if (lw.dyn) {
      lw.fu = m.take<float>(2 * B * D);
      lw.fu_p = m.planes(2 * B * D, split);
      lw.g = m.take<float>(2 * B * D);
      lw.beta_u = m.take<float>(2 * B);
      lw.edges_out = m.take<float>(4 * BC);
    }
    lw.z = m.planes((size_t)lw.rows * D, split);
    lw.h = m.take<float>((size_t)lw.rows * D);
  }
  if (c.training) {
    const size_t rows = 2 * B + 2 * BC;
    ws.dh = m.planes(rows * D, split);
    ws.dz = m.take<float>(rows * D);
    ws.dxm = m.take<float>(2 * B * D);
    ws.dxu = m.take<float>(2 * B * D);
    ws.dg = m.take<float>(2 * B * D);
    ws.dg_p = m.planes(2 * B * D, split);
    ws.dbeta = m.take<float>(2 * B);
    ws.dfu = m.take<float>(2 * B * D);
    ws.dfu_p = m.planes(2 * B * D, split);
    ws.dedges = m.take<float>(2 * 4 * BC);           // ping-pong between layers
    ws.dx0 = m.planes(rows * D, split);
    if (ws.slices > 1) {
      ws.slice_part = m.take<float>(B * ws.slices * 4 * D);      // also holds the [D + 32] slice records of score_bwd
      ws.slice_dbeta = m.take<float>(B * ws.slices * 2);
    }
    // split-K: enough slices to fill the machine for the [D, D] weight gradients (18 tiles each).  Every
    // weight-gradient GEMM of a pass owns a region of the arena, so that all reductions can run in ONE launch at the end.
    ws.ksplit = 8;
    ws.partial_floats = (size_t)(3 * L + 2) * ws.ksplit * D * D + (size_t)2 * 3 * D * R;
    ws.partial = m.take<float>(ws.partial_floats);
    ws.colsum_ctas = backward_ctas();
    ws.colsum_floats = (size_t)(3 * L + 3) * ws.colsum_ctas * 3 * D;     // same idea for the column-sum partials
    ws.colsum = m.take<float>(ws.colsum_floats);
  }
  ws.bytes = align_up(m.off, 256);
  return DRIN_OK;
}

static Operand op(const Planes& p, long long rows, int cols, long long row_offset = 0) {
  Operand o;
  o.hi = p.hi + row_offset * cols;
  o.lo = p.lo ? p.lo + row_offset * cols : nullptr;
  o.rows = rows;
  o.cols = cols;
  o.ld = cols;
  return o;
}

// W_m[:, :H] (half = 0) or W_m[:, H:] (half = 1) as a [D, H] operand inside the [D, D] planes (leading dimension D)
static Operand wm_half(const Planes& w_m, int D, int half) {
  Operand o;
  const int H = D / 2;
  o.hi = w_m.hi + half * H;
  o.lo = w_m.lo ? w_m.lo + half * H : nullptr;
  o.rows = D;
  o.cols = H;
  o.ld = D;
  return o;
}

// GCN layers + scoring with vector edges (drin/model.py:205-209 with gcn_edge_feature == "vector").
static int forward_vector_layers(const drin_config& c, const drin_params& p, Workspace& ws, float* scores,
                                 cudaStream_t stream) {
  const long long B = c.batch, C = c.candidates, BC = B * C;
  const int D = c.embed_dim, H = D / 2, L = c.gcn_layers;
  for (int l = 0; l < L; ++l) {
    LayerWs& lw = ws.layer[l];
    const drin_layer_params& lp = p.layer[l];
    VecLayerArgs va{};
    va.B = c.batch; va.C = c.candidates; va.D = D; va.full = lw.full; va.dyn = lw.dyn;
    for (int k = 0; k < 4; ++k) va.en[k] = c.edge_enabled[k];
    if (l == 0) {
      va.e_scalar = ws.edges0;
    } else {
      const LayerWs& pw = ws.layer[l - 1];
      const drin_layer_params& pp = p.layer[l - 1];
      DRIN_TRY(mention_ln(stream, D, pw.h, 2 * B + 2 * BC, pp.ln_w, pp.ln_b, lw.xa, lw.dyn ? lw.xa_p.hi : nullptr,
                          lw.dyn ? lw.xa_p.lo : nullptr));
      if (pw.affine) {
        va.e_scalar = ws.edges0; va.edge_a = pw.edge_a; va.edge_bv = pw.edge_bv; va.edge_w1 = pw.edge_w1;
      } else {
        va.q_in = pw.q;
      }
    }
    va.xa = lw.xa;
    va.dyn = lw.dyn && !lw.affine;
    if (lw.dyn) {
      // model.py:149: fu = W_u u for the 2B mention vertices, fv = W_v v for the 2BC candidate vertices (D -> D/2)
      GemmEpilogue ep;
      ep.ldc = H; ep.ld_planes = H;
      ep.C = lw.fu; ep.bias = lp.b_u; ep.out_hi = lw.fu_p.hi; ep.out_lo = lw.fu_p.lo;     // affine: planes, else fp32
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.xa_p, 2 * B, D), op(lw.w_u, H, D), 2 * B, H, D, ep));
      ep.C = lw.fv; ep.bias = lp.b_v; ep.out_hi = lw.fv_p.hi; ep.out_lo = lw.fv_p.lo;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.xa_p, 2 * BC, D, 2 * B), op(lw.w_v, H, D), 2 * BC, H, D, ep));
      if (lw.affine) {
        // first layer: W_m(cat[fu, fv] + e 1) + b_m = (fu W_m[:, :H]^T + b_m) + fv W_m[:, H:]^T + e (W_m 1); nothing per edge type
        GemmEpilogue ea;
        ea.ldc = D; ea.C = lw.edge_a; ea.bias = lp.b_m;
        DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.fu_p, 2 * B, H), wm_half(lw.w_m, D, 0), 2 * B, D, H, ea));
        ea.C = lw.edge_bv; ea.bias = nullptr;
        DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.fv_p, 2 * BC, H), wm_half(lw.w_m, D, 1), 2 * BC, D, H, ea));
        DRIN_TRY(rowsum(stream, lp.w_m, D, D, lw.edge_w1));
      } else {
        va.fu = lw.fu; va.fv = lw.fv;
        va.m_hi = lw.m_p.hi; va.m_lo = lw.m_p.lo;
      }
    }
    va.z_hi = lw.z.hi; va.z_lo = lw.z.lo;
    DRIN_TRY(vec_layer_fwd(stream, va));
    GemmEpilogue eh;
    eh.ldc = D; eh.C = lw.h; eh.bias = lp.b_h;
    DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.z, lw.rows, D), op(lw.w_h, D, D), lw.rows, D, D, eh));
    if (va.dyn) {   // model.py:133: q = W_m(cat[fu, fv] + e) + b_m for the four edge types at once; sigmoid in the consumer
      GemmEpilogue em;
      em.ldc = D; em.C = lw.q; em.bias = lp.b_m;
      DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.m_p, 4 * BC, D), op(lw.w_m, D, D), 4 * BC, D, D, em));
    }
  }
  const LayerWs& last = ws.layer[L - 1];
  return score_fwd(stream, D, last.h, last.h + B * D, p.layer[L - 1].ln_w, p.layer[L - 1].ln_b, c.batch, c.candidates,
                   scores);
}

int forward(const drin_config& c, const drin_inputs& in, const drin_params& p, void* workspace, size_t workspace_bytes,
            float* scores, cudaStream_t stream) {
  Workspace ws;
  DRIN_TRY(plan_workspace(c, &in, workspace, ws));
  if (!workspace || workspace_bytes < ws.bytes)
    return fail(DRIN_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, ws.bytes);
  if (!scores) return fail(DRIN_ERR_ARG, "scores is null");
  const bool bf16_in = c.precision == DRIN_BF16;
  const long long B = c.batch, C = c.candidates, BC = B * C;
  const int D = c.embed_dim, R = c.resnet_dim, L = c.gcn_layers;

  SideStream side(stream);
  // ---- weights -> bf16 planes (they change every optimizer step): one launch for all matrices ----
  {
    SplitJobs jobs;
    auto add = [&](const float* w, const Planes& pl, size_t n) { jobs.job[jobs.count++] = SplitJob{w, pl.hi, pl.lo, (long long)(n / 4)}; };
    add(p.w_mt, ws.w_mt, (size_t)D * D);
    add(p.w_et, ws.w_et, (size_t)D * D);
    add(p.w_mi, ws.w_mi, (size_t)D * R);
    add(p.w_ei, ws.w_ei, (size_t)D * R);
    for (int l = 0; l < L; ++l) {
      add(p.layer[l].w_h, ws.layer[l].w_h, (size_t)D * D);
      if (ws.layer[l].dyn) {
        const size_t uv = ws.vec ? (size_t)(D / 2) * D : (size_t)D * D;
        add(p.layer[l].w_u, ws.layer[l].w_u, uv);
        add(p.layer[l].w_v, ws.layer[l].w_v, uv);
        if (ws.vec) {
          if (!p.layer[l].w_m || !p.layer[l].b_m) return fail(DRIN_ERR_ARG, "vector edges: layer %d has no w_m / b_m", l);
          add(p.layer[l].w_m, ws.layer[l].w_m, (size_t)D * D);
        }
      }
    }
    // the weight planes are not needed before the first GEMM: produced on the side stream, beside the front end
    DRIN_TRY(split_planes_multi(side.fork(), jobs));
  }

  // ---- front end: pooling, edges, projection operands ----
  FrontendArgs fa{};
  fa.B = c.batch; fa.C = c.candidates; fa.Lm = c.mention_tokens; fa.Le = c.entity_tokens; fa.P = c.regions;
  fa.Om = c.mention_objects; fa.Oe = c.entity_objects; fa.D = D; fa.R = R;
  fa.mtf = in.mention_text_feature; fa.start = (const long long*)in.mention_start_pos;
  fa.end = (const long long*)in.mention_end_pos; fa.mif = in.mention_image_feature;
  fa.mof = in.mention_object_feature; fa.mos = in.mention_object_score; fa.etf = in.entity_text_feature;
  fa.emask = (const long long*)in.entity_text_mask; fa.eif = in.entity_image_feature;
  fa.eof = in.entity_object_feature; fa.eos = in.entity_object_score; fa.miet = in.miet_similarity;
  fa.mtei = in.mtei_similarity;
  if (c.indexed) {
    if (!in.mention_index) return fail(DRIN_ERR_ARG, "indexed inputs need drin_inputs.mention_index");
    if (c.entity_tokens > 0 && !in.entity_index)
      return fail(DRIN_ERR_ARG, "indexed WikiMEL-layout inputs need drin_inputs.entity_index (entity tables are per entity)");
    fa.mention_index = (const long long*)in.mention_index;
    fa.entity_index = (const long long*)in.entity_index;
  }
  fa.span_hi = ws.span.hi; fa.span_lo = ws.span.lo; fa.mim_hi = ws.mim.hi; fa.mim_lo = ws.mim.lo;
  const bool own_epool = !bf16_in || c.entity_tokens > 0 || c.indexed;
  const bool own_eimg = !bf16_in || c.indexed;
  fa.ep_hi = own_epool ? ws.epool.hi : nullptr; fa.ep_lo = own_epool ? ws.epool.lo : nullptr;
  fa.ei_hi = own_eimg ? ws.eimg.hi : nullptr; fa.ei_lo = own_eimg ? ws.eimg.lo : nullptr;
  fa.edges = ws.edges0;
  DRIN_TRY(frontend(stream, fa, bf16_in));
  DRIN_TRY(side.join());

  // ---- input projections (ghmfc.py:66-69,250; model.py:42,45): x0 = [mt; mi; et; ei] ----
  // The mention-side chain (W_mt, W_mi and, for a dynamic first layer, fu = W_u xm, beta = fu . b_v, g = fu W_v) only
  // has 2B rows per GEMM: it runs on the side stream, in the shadow of the two candidate-side projections.
  auto fu_chain = [&](cudaStream_t s, int l) -> int {      // model.py:149-150 for layer l (the W_v GEMM is folded)
    LayerWs& lw = ws.layer[l];
    const drin_layer_params& lp = p.layer[l];
    GemmEpilogue ep;
    ep.ldc = D; ep.ld_planes = D;
    ep.C = lw.fu; ep.bias = lp.b_u; ep.out_hi = lw.fu_p.hi; ep.out_lo = lw.fu_p.lo;
    DRIN_TRY(gemm_tcgen05(s, GEMM_NT, op(lw.xm_p, 2 * B, D), op(lw.w_u, D, D), 2 * B, D, D, ep));
    DRIN_TRY(rowdot(s, D, lw.fu, 2 * B, lp.b_v, lw.beta_u));
    GemmEpilogue eg;
    eg.ldc = D; eg.C = lw.g;
    return gemm_tcgen05(s, GEMM_NN, op(lw.fu_p, 2 * B, D), op(lw.w_v, D, D), 2 * B, D, D, eg);
  };
  bool fu0_done = false;
  {
    GemmEpilogue ep;
    ep.ldc = D; ep.ld_planes = D;
    cudaStream_t ms = side.fork();
    ep.C = ws.x0; ep.bias = p.b_mt; ep.out_hi = ws.xm0_p.hi; ep.out_lo = ws.xm0_p.lo;
    DRIN_TRY(gemm_tcgen05(ms, GEMM_NT, op(ws.span, B, D), op(ws.w_mt, D, D), B, D, D, ep));
    ep.C = ws.x0 + B * D; ep.bias = p.b_mi; ep.out_hi = ws.xm0_p.hi + B * D;
    ep.out_lo = ws.xm0_p.lo ? ws.xm0_p.lo + B * D : nullptr;
    DRIN_TRY(gemm_tcgen05(ms, GEMM_NT, op(ws.mim, B, R), op(ws.w_mi, D, R), B, D, R, ep));
    if (side.on() && !ws.vec && ws.layer[0].dyn) {
      DRIN_TRY(fu_chain(ms, 0));
      fu0_done = true;
    }
    ep.out_hi = ep.out_lo = nullptr;
    auto planes_at = [&](long long row) {          // vector edges: the candidate rows feed the W_v GEMM too
      if (!ws.vec) return;
      ep.out_hi = ws.x0_p.hi + row * D;
      ep.out_lo = ws.x0_p.lo ? ws.x0_p.lo + row * D : nullptr;
    };
    ep.C = ws.x0 + 2 * B * D; ep.bias = p.b_et;
    planes_at(2 * B);
    DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(ws.epool, BC, D), op(ws.w_et, D, D), BC, D, D, ep));
    ep.C = ws.x0 + (2 * B + BC) * D; ep.bias = p.b_ei;
    planes_at(2 * B + BC);
    DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(ws.eimg, BC, R), op(ws.w_ei, D, R), BC, D, R, ep));
    DRIN_TRY(side.join());
  }
  if (ws.vec) return forward_vector_layers(c, p, ws, scores, stream);

  // ---- GCN layers (model.py:205-206) ----
  for (int l = 0; l < L; ++l) {
    LayerWs& lw = ws.layer[l];
    const drin_layer_params& lp = p.layer[l];
    LayerFwdArgs la{};
    la.B = c.batch; la.C = c.candidates; la.D = D; la.full = lw.full;
    for (int k = 0; k < 4; ++k) la.en[k] = c.edge_enabled[k];
    if (l == 0) {
      la.x_et = ws.x0 + 2 * B * D;
      la.x_ei = ws.x0 + (2 * B + BC) * D;
      la.edges_in = ws.edges0;
    } else {
      const LayerWs& pw = ws.layer[l - 1];
      const drin_layer_params& pp = p.layer[l - 1];
      DRIN_TRY(mention_ln(stream, D, pw.h, 2 * B, pp.ln_w, pp.ln_b, lw.xm, lw.dyn ? lw.xm_p.hi : nullptr,
                          lw.dyn ? lw.xm_p.lo : nullptr));
      la.x_et = pw.h + 2 * B * D;
      la.x_ei = pw.h + (2 * B + BC) * D;
      la.ln_gamma = pp.ln_w;
      la.ln_beta = pp.ln_b;
      if (pw.dyn) {
        la.edges_in = pw.edges_out;
      } else {
        // static edges (model.py:135-136): layer l sees the input edges masked l + 1 times (model.py:122)
        la.edges_in = ws.edges0;
        for (int k = 0; k < 4; ++k) la.en[k] = powf(c.edge_enabled[k], (float)(l + 1));
      }
    }
    la.xm = lw.xm;
    if (lw.dyn) {
      if (!(l == 0 && fu0_done)) DRIN_TRY(fu_chain(stream, l));
      la.g = lw.g;
      la.beta_u = lw.beta_u;
      la.edges_out = lw.edges_out;
    }
    la.z_hi = lw.z.hi;
    la.z_lo = lw.z.lo;
    la.slices = ws.slices;
    la.acc_part = ws.acc_part;
    DRIN_TRY(gcn_layer_fwd(stream, la));
    GemmEpilogue eh;
    eh.ldc = D; eh.C = lw.h; eh.bias = lp.b_h;
    DRIN_TRY(gemm_tcgen05(stream, GEMM_NT, op(lw.z, lw.rows, D), op(lw.w_h, D, D), lw.rows, D, D, eh));
  }

  // ---- candidate scoring (model.py:207-209) ----
  const LayerWs& last = ws.layer[L - 1];
  DRIN_TRY(score_fwd(stream, D, last.h, last.h + B * D, p.layer[L - 1].ln_w, p.layer[L - 1].ln_b, c.batch,
                     c.candidates, scores));
  return DRIN_OK;
}

}  // namespace drin
