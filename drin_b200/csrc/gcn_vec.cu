// GCN layer with VECTOR edges (gcn_edge_feature == "vector", reference common/args.py:33, drin/model.py:97-153).
//
// Every edge of the per-candidate relation graph is a D-vector, so a layer is column-wise work only:
//   messages      a_mt = mean_c(E0*et) + mean_c(E1*ei)      a_et = E0*mt + E2*mi           (model.py:139-146)
//                 a_mi = mean_c(E2*et) + mean_c(E3*ei)      a_ei = E1*mt + E3*mi
//   edge update   E_k' = sigmoid(W_m(cat[W_u u_k, W_v v_k] + E_k) + b_m)                   (model.py:133,148-152)
// with (u_k, v_k) = (mt,et), (mt,ei), (mi,et), (mi,ei) and `*` elementwise.  A thread owns FOUR columns of one
// mention and walks its candidates: no cross-lane reduction anywhere, every access is a coalesced 16-byte one, the
// sums over candidates (messages to the mention vertices, W_u gradients) and the kernel-long column sums (bias
// gradients) live in registers.  The three contractions of the layer (W_h, W_u / W_v, W_m) are tcgen05 GEMMs fed
// by the bf16 planes these kernels emit (engine.cu).  Edges travel between layers PRE-sigmoid (`q`), the consumer
// applies sigmoid and the enable mask (model.py:122) on the fly; the first layer reads the scalar input edges.
// The [4BC, D] edge matrices (m, q, dm, dq) are CANDIDATE-major, row r * 4 + k: the four edge vectors of a candidate
// are one contiguous 12 KB block, so a CTA walks 8 HBM streams instead of 17 (the GEMMs do not care about row order).
//
// FIRST-layer shortcut ("affine" edges): the first layer's own edges are scalars e_k broadcast over D, so
//   W_m(cat[fu, fv] + e_k 1) + b_m = (fu W_m[:, :H]^T + b_m) + fv W_m[:, H:]^T + e_k (W_m 1)  =  A_u + Bv_v + e_k w1 :
// the [4BC, D] x [D, D] GEMM over the four edge types shrinks to a [2BC, D/2] x [D/2, D] one, the m / q matrices of that
// layer are never built, and the consumer layer evaluates sigmoid(A_u[b] + Bv_v[r] + e_k[r] w1) on the fly from two
// candidate rows instead of four.  Its backward hands back dA, dBv (planes) and sum e_k dq_k (the w1 gradient).
#include "kernels.cuh"
#include "rows.cuh"

namespace drin {

static constexpr int VEC_GRID = 148 * 8;      // upper bound of the persistent grid (rows of the partial-sum buffers)
static constexpr int VR_NW = 4;               // warps per CTA of vec_rows_bwd
static constexpr int VR_CTAS = 148 * 3;

int vec_layer_ctas() { return VEC_GRID; }
int vec_rows_ctas() { return VR_CTAS; }

enum EdgeMode : int { EDGE_VECTOR = 0, EDGE_SCALAR = 1, EDGE_AFFINE = 2 };

namespace {

// W consecutive columns of one row (W = 4: 16-byte accesses, D/4 threads per mention; W = 2: 8-byte accesses, D/2
// threads per mention and half the registers per thread -- for the variants that are register-limited)
template <int W>
struct VF {
  float v[W];
};
template <int W>
__device__ __forceinline__ VF<W> ldv(const float* p) {
  VF<W> r;
  if (W == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[W - 2] = t.z; r.v[W - 1] = t.w;
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}
template <int W>
__device__ __forceinline__ void stv(float* p, const VF<W>& x) {
  if (W == 4) *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[W - 2], x.v[W - 1]);
  else *reinterpret_cast<float2*>(p) = make_float2(x.v[0], x.v[1]);
}
template <int W>
__device__ __forceinline__ VF<W> splat(float s) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = s;
  return r;
}
template <int W>
__device__ __forceinline__ VF<W> operator+(const VF<W>& a, const VF<W>& b) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = a.v[i] + b.v[i];
  return r;
}
template <int W>
__device__ __forceinline__ VF<W> operator*(const VF<W>& a, const VF<W>& b) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = a.v[i] * b.v[i];
  return r;
}
template <int W>
__device__ __forceinline__ VF<W> operator*(const VF<W>& a, float s) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = a.v[i] * s;
  return r;
}
template <int W>
__device__ __forceinline__ VF<W> fmav(const VF<W>& a, const VF<W>& b, const VF<W>& c) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = fmaf(a.v[i], b.v[i], c.v[i]);
  return r;
}
template <int W>
__device__ __forceinline__ VF<W> fmas(const VF<W>& a, float s, const VF<W>& c) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = fmaf(a.v[i], s, c.v[i]);
  return r;
}
// sigmoid from the two MUFU approximations (ex2, rcp; ~2 ulp each): 4 instructions instead of an IEEE division
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.0f + exp2_ftz(x * -1.4426950408889634f)); }
template <int W>
__device__ __forceinline__ VF<W> sigmoidv(const VF<W>& q) {
  VF<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.v[i] = sigmoid_f(q.v[i]);
  return r;
}
// W consecutive columns -> split-bf16 planes (lo may be null: plain bf16 rounding)
template <int W>
__device__ __forceinline__ void st_planes(bf16* hi, bf16* lo, long long idx, const VF<W>& x) {
  uint32_t h0, l0;
  split_bf16x2(x.v[0], x.v[1], h0, l0);
  if (W == 4) {
    uint32_t h1, l1;
    split_bf16x2(x.v[W - 2], x.v[W - 1], h1, l1);
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h0, h1);
    if (lo) *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l0, l1);
  } else {
    *reinterpret_cast<uint32_t*>(hi + idx) = h0;
    if (lo) *reinterpret_cast<uint32_t*>(lo + idx) = l0;
  }
}

// Unmasked edge vector S_k of candidate row r (the caller applies the enable mask: E_k = en_k S_k).
//   SCALAR  S = e_k[r]                                         (first layer: input edges, model.py:201-204)
//   VECTOR  S = sigmoid(q[4r + k])                             (q written by the W_m GEMM of the layer below)
//   AFFINE  S = sigmoid(A_u[b] + Bv_v[r] + (en_k e_k[r]) w1)   (layer below was a first layer: see the file header)
template <int W, int EDGE>
__device__ __forceinline__ VF<W> edge_value(const VecLayerArgs& a, long long BC, long long r, int col, int k,
                                            const VF<W>& a_u, const VF<W>& bv_v, const VF<W>& w1, float& e_masked) {
  if (EDGE == EDGE_SCALAR) return splat<W>(a.e_scalar[k * BC + r]);
  if (EDGE == EDGE_VECTOR) return sigmoidv<W>(ldv<W>(a.q_in + (r * 4 + k) * a.D + col));
  e_masked = a.e_scalar[k * BC + r] * a.en[k];
  return sigmoidv<W>(fmas<W>(w1, e_masked, a_u + bv_v));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// forward: z (A operand of the W_h GEMM) and, in a layer with a general edge update, m_k = cat[fu, fv] + E_k
// ---------------------------------------------------------------------------------------------
template <int D, int W, bool FULL, bool DYN, int EDGE>
__global__ void __launch_bounds__(D / W) vec_layer_fwd_kernel(const VecLayerArgs a) {
  constexpr int H = D / 2;
  typedef VF<W> V;
  const int col = threadIdx.x * W;
  const long long B = a.B, C = a.C, BC = B * C;
  const long long row_et = FULL ? 2 * B : B;          // first et row of this layer's z / h layout
  const float c_den = (float)a.C;
  const V zero = splat<W>(0.f);
  const V w1 = EDGE == EDGE_AFFINE ? ldv<W>(a.edge_w1 + col) : zero;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const V mt = ldv<W>(a.xa + b * D + col);
    const V mi = ldv<W>(a.xa + (B + b) * D + col);
    V fu_mt = zero, fu_mi = zero, a_mt = zero, a_mi = zero;
    if (DYN && col < H) {
      fu_mt = ldv<W>(a.fu + b * H + col);
      fu_mi = ldv<W>(a.fu + (B + b) * H + col);
    }
    if (EDGE == EDGE_AFFINE) {
      a_mt = ldv<W>(a.edge_a + b * D + col);
      a_mi = ldv<W>(a.edge_a + (B + b) * D + col);
    }
    V amt = zero, ami = zero;
    for (long long c = 0; c < C; ++c) {
      const long long r = b * C + c;
      const V et = ldv<W>(a.xa + (2 * B + r) * D + col);
      const V ei = ldv<W>(a.xa + (2 * B + BC + r) * D + col);
      V bv_et = zero, bv_ei = zero;
      if (EDGE == EDGE_AFFINE) {
        bv_et = ldv<W>(a.edge_bv + r * D + col);
        bv_ei = ldv<W>(a.edge_bv + (BC + r) * D + col);
      }
      V v_et = zero, v_ei = zero;
      if (DYN && col >= H) {
        v_et = ldv<W>(a.fv + r * H + (col - H));
        v_ei = ldv<W>(a.fv + (BC + r) * H + (col - H));
      }
      V zet = et, zei = ei;
      // one edge type at a time (k = 2 u + v: u = mt | mi, v = et | ei):  a_u += E_k x_v / C,  z_v += E_k m_u
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float eh = 0.f;
        const V E = edge_value<W, EDGE>(a, BC, r, col, k, k < 2 ? a_mt : a_mi, (k & 1) ? bv_ei : bv_et, w1, eh) * a.en[k];
        const V& x_v = (k & 1) ? ei : et;
        const V& m_u = k < 2 ? mt : mi;
        if (k < 2) amt = fmav<W>(E, x_v, amt);
        else if (FULL) ami = fmav<W>(E, x_v, ami);
        if (k & 1) { if (FULL) zei = fmav<W>(E, m_u, zei); }
        else zet = fmav<W>(E, m_u, zet);
        if (DYN) {
          const V base = col < H ? (k < 2 ? fu_mt : fu_mi) : ((k & 1) ? v_ei : v_et);
          st_planes<W>(a.m_hi, a.m_lo, (r * 4 + k) * D + col, base + E);
        }
      }
      st_planes<W>(a.z_hi, a.z_lo, (row_et + r) * D + col, zet);
      if (FULL) st_planes<W>(a.z_hi, a.z_lo, (2 * B + BC + r) * D + col, zei);
    }
    V zm;
#pragma unroll
    for (int i = 0; i < W; ++i) zm.v[i] = amt.v[i] / c_den + mt.v[i];       // mean over ALL C slots (model.py:144)
    st_planes<W>(a.z_hi, a.z_lo, b * D + col, zm);
    if (FULL) {
#pragma unroll
      for (int i = 0; i < W; ++i) zm.v[i] = ami.v[i] / c_den + mi.v[i];
      st_planes<W>(a.z_hi, a.z_lo, (B + b) * D + col, zm);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward of the kernel above.  In: dz (gradient w.r.t. z) and, with a general edge update in this layer, dm (gradient
// w.r.t. m, = dq W_m).  Out: dxa (gradient w.r.t. the activated vertices, without the W_u / W_v paths, which are
// GEMMs), dfu / dfv planes, the gradient w.r.t. the edge outputs of the layer BELOW (VECTOR: dq planes; AFFINE: dA, dBv
// planes), column partials [cta][3][D]: sum dq (b_m of the layer below) | [b_u, b_v] of this layer | sum e_k dq_k (w1).
// ---------------------------------------------------------------------------------------------
template <int D, int W, bool FULL, bool DYN, int EDGE>
__global__ void __launch_bounds__(D / W, W == 4 ? 3 : 2) vec_layer_bwd_kernel(const VecLayerArgs a) {
  constexpr int H = D / 2;
  typedef VF<W> V;
  const int col = threadIdx.x * W;
  const long long B = a.B, C = a.C, BC = B * C;
  const long long row_et = FULL ? 2 * B : B;
  const float inv_c = 1.0f / (float)a.C;
  const V zero = splat<W>(0.f);
  const V w1 = EDGE == EDGE_AFFINE ? ldv<W>(a.edge_w1 + col) : zero;
  V p_bm = zero, p_uv = zero, p_w1 = zero;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const V mt = ldv<W>(a.xa + b * D + col);
    const V mi = ldv<W>(a.xa + (B + b) * D + col);
    const V dzmt = ldv<W>(a.dz + b * D + col);
    const V dzmi = FULL ? ldv<W>(a.dz + (B + b) * D + col) : zero;
    const V smt = dzmt * inv_c, smi = dzmi * inv_c;
    V dmt = dzmt, dmi = dzmi;
    V dfu_mt = zero, dfu_mi = zero, a_mt = zero, a_mi = zero, da_mt = zero, da_mi = zero;
    if (EDGE == EDGE_AFFINE) {
      a_mt = ldv<W>(a.edge_a + b * D + col);
      a_mi = ldv<W>(a.edge_a + (B + b) * D + col);
    }
    for (long long c = 0; c < C; ++c) {
      const long long r = b * C + c;
      const V et = ldv<W>(a.xa + (2 * B + r) * D + col);
      const V ei = ldv<W>(a.xa + (2 * B + BC + r) * D + col);
      const V dzet = ldv<W>(a.dz + (row_et + r) * D + col);
      const V dzei = FULL ? ldv<W>(a.dz + (2 * B + BC + r) * D + col) : zero;
      V bv_et = zero, bv_ei = zero;
      if (EDGE == EDGE_AFFINE) {
        bv_et = ldv<W>(a.edge_bv + r * D + col);
        bv_ei = ldv<W>(a.edge_bv + (BC + r) * D + col);
      }
      V det = dzet, dei = dzei;
      V dfv_et = zero, dfv_ei = zero, dbv_et = zero, dbv_ei = zero;
      // one edge type at a time (k = 2 u + v: u = mt | mi, v = et | ei) keeps a single edge vector live:
      //   forward   a_u += E_k x_v / C,  a_v += E_k m_u
      //   backward  dx_v += (dz_u / C) E_k,  dm_u += dz_v E_k,  dE_k = (dz_u / C) x_v + dz_v m_u
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const V& s_u = k < 2 ? smt : smi;
        const V& m_u = k < 2 ? mt : mi;
        const V& x_v = (k & 1) ? ei : et;
        const V& dz_v = (k & 1) ? dzei : dzet;
        float eh = 0.f;
        const V S = edge_value<W, EDGE>(a, BC, r, col, k, k < 2 ? a_mt : a_mi, (k & 1) ? bv_ei : bv_et, w1, eh);
        const V E = S * a.en[k];
        if (k & 1) dei = fmav<W>(s_u, E, dei); else det = fmav<W>(s_u, E, det);
        if (k < 2) dmt = fmav<W>(dz_v, E, dmt); else dmi = fmav<W>(dz_v, E, dmi);
        if (DYN || EDGE != EDGE_SCALAR) {
          V dE = fmav<W>(s_u, x_v, dz_v * m_u);
          if (DYN) {
            const V dmk = ldv<W>(a.dm + (r * 4 + k) * D + col);
            dE = dE + dmk;
            if (col < H) {
              if (k < 2) dfu_mt = dfu_mt + dmk; else dfu_mi = dfu_mi + dmk;
            } else {
              if (k & 1) dfv_ei = dfv_ei + dmk; else dfv_et = dfv_et + dmk;
            }
          }
          if (EDGE != EDGE_SCALAR) {
            // E_k = en_k * S_k, S_k = sigmoid(q_k):  dq_k = dE_k * en_k * S_k * (1 - S_k)
            V dq;
#pragma unroll
            for (int i = 0; i < W; ++i) dq.v[i] = dE.v[i] * (S.v[i] * (1.f - S.v[i])) * a.en[k];
            p_bm = p_bm + dq;
            if (EDGE == EDGE_VECTOR) {
              st_planes<W>(a.dq_hi, a.dq_lo, (r * 4 + k) * D + col, dq);
            } else {     // q_k = A_u + Bv_v + e_k w1
              if (k < 2) da_mt = da_mt + dq; else da_mi = da_mi + dq;
              if (k & 1) dbv_ei = dbv_ei + dq; else dbv_et = dbv_et + dq;
              p_w1 = fmas<W>(dq, eh, p_w1);
            }
          }
        }
      }
      stv<W>(a.dxa + (2 * B + r) * D + col, det);
      stv<W>(a.dxa + (2 * B + BC + r) * D + col, dei);
      if (DYN && col >= H) {
        st_planes<W>(a.dfv_hi, a.dfv_lo, r * H + (col - H), dfv_et);
        st_planes<W>(a.dfv_hi, a.dfv_lo, (BC + r) * H + (col - H), dfv_ei);
        p_uv = p_uv + dfv_et + dfv_ei;
      }
      if (EDGE == EDGE_AFFINE) {
        st_planes<W>(a.dbv_hi, a.dbv_lo, r * D + col, dbv_et);
        st_planes<W>(a.dbv_hi, a.dbv_lo, (BC + r) * D + col, dbv_ei);
      }
    }
    stv<W>(a.dxa + b * D + col, dmt);
    stv<W>(a.dxa + (B + b) * D + col, dmi);
    if (DYN && col < H) {
      st_planes<W>(a.dfu_hi, a.dfu_lo, b * H + col, dfu_mt);
      st_planes<W>(a.dfu_hi, a.dfu_lo, (B + b) * H + col, dfu_mi);
      p_uv = p_uv + dfu_mt + dfu_mi;
    }
    if (EDGE == EDGE_AFFINE) {
      st_planes<W>(a.da_hi, a.da_lo, b * D + col, da_mt);
      st_planes<W>(a.da_hi, a.da_lo, (B + b) * D + col, da_mi);
    }
  }
  stv<W>(a.partials + ((long long)blockIdx.x * 3 + 0) * D + col, p_bm);
  stv<W>(a.partials + ((long long)blockIdx.x * 3 + 1) * D + col, p_uv);
  stv<W>(a.partials + ((long long)blockIdx.x * 3 + 2) * D + col, p_w1);
}

static int g_vec_bwd_width = 0;          // 0 auto, 2 or 4 columns per thread (test / A-B hook)
void debug_set_vec_bwd_width(int v) { g_vec_bwd_width = (v == 2 || v == 4) ? v : 0; }
static int g_vec_ctas_per_sm = 0;        // 0 auto (as many as are resident), else an upper bound (A-B hook)
void debug_set_vec_ctas_per_sm(int v) { g_vec_ctas_per_sm = v > 0 ? v : 0; }

// Persistent grid = (resident CTAs per SM) x 148, so that every CTA of the grid-stride loop over the mentions is on an
// SM from the start: a grid larger than what is resident would run its tail at a fraction of the occupancy.
template <typename K>
static int launch_vec_kernel(K kernel, int threads, cudaStream_t stream, const VecLayerArgs& a, int* grid_out) {
  int resident = 0;
  DRIN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, threads, 0));
  if (resident < 1) resident = 1;
  if (resident > VEC_GRID / 148) resident = VEC_GRID / 148;
  if (g_vec_ctas_per_sm && resident > g_vec_ctas_per_sm) resident = g_vec_ctas_per_sm;
  const int grid = a.B < 148 * resident ? a.B : 148 * resident;
  kernel<<<grid, threads, 0, stream>>>(a);
  if (grid_out) *grid_out = grid;
  return DRIN_OK;
}

template <int D, int W, bool BWD, bool FULL, bool DYN>
static int launch_vec(cudaStream_t stream, const VecLayerArgs& a, int edge, int* grid) {
  if (edge == EDGE_SCALAR) {
    if (BWD) return launch_vec_kernel(vec_layer_bwd_kernel<D, W, FULL, DYN, EDGE_SCALAR>, D / W, stream, a, grid);
    return launch_vec_kernel(vec_layer_fwd_kernel<D, W, FULL, DYN, EDGE_SCALAR>, D / W, stream, a, grid);
  } else if (edge == EDGE_VECTOR) {
    if (BWD) return launch_vec_kernel(vec_layer_bwd_kernel<D, W, FULL, DYN, EDGE_VECTOR>, D / W, stream, a, grid);
    return launch_vec_kernel(vec_layer_fwd_kernel<D, W, FULL, DYN, EDGE_VECTOR>, D / W, stream, a, grid);
  }
  if (BWD) return launch_vec_kernel(vec_layer_bwd_kernel<D, W, FULL, DYN, EDGE_AFFINE>, D / W, stream, a, grid);
  return launch_vec_kernel(vec_layer_fwd_kernel<D, W, FULL, DYN, EDGE_AFFINE>, D / W, stream, a, grid);
}

template <int W, bool BWD>
static int launch_vec_shape(cudaStream_t stream, const VecLayerArgs& a, int edge, int* grid) {
  if (a.full) {
    if (a.dyn) return launch_vec<768, W, BWD, true, true>(stream, a, edge, grid);
    return launch_vec<768, W, BWD, true, false>(stream, a, edge, grid);
  }
  return launch_vec<768, W, BWD, false, false>(stream, a, edge, grid);
}

template <bool BWD>
static int vec_layer_dispatch(cudaStream_t stream, const VecLayerArgs& a, int* grid) {
  if (a.D != 768) return fail(DRIN_ERR_ARG, "vector-edge GCN layer: gcn_embed_dim %d not built (768 only)", a.D);
  const int edge = a.edge_a ? EDGE_AFFINE : (a.e_scalar ? EDGE_SCALAR : EDGE_VECTOR);
  if (!a.xa || (edge == EDGE_VECTOR && !a.q_in) || (edge == EDGE_AFFINE && (!a.edge_bv || !a.edge_w1 || !a.e_scalar)))
    return fail(DRIN_ERR_ARG, "vector-edge GCN layer: missing inputs");
  if (a.dyn && (!a.full || edge == EDGE_SCALAR))
    return fail(DRIN_ERR_ARG, "vector-edge GCN layer: a general edge update needs a full layer with vector edges in");
  if (!BWD) {
    if (!a.z_hi || (a.dyn && (!a.fu || !a.fv || !a.m_hi)))
      return fail(DRIN_ERR_ARG, "vector-edge GCN layer forward: missing buffers");
  } else {
    if (!a.dz || !a.dxa || !a.partials || (a.dyn && (!a.dm || !a.dfu_hi || !a.dfv_hi)) ||
        (edge == EDGE_VECTOR && !a.dq_hi) || (edge == EDGE_AFFINE && (!a.da_hi || !a.dbv_hi)))
      return fail(DRIN_ERR_ARG, "vector-edge GCN layer backward: missing buffers");
  }
  // 2 columns per thread (D/2 threads per mention) is kept as an A/B option for the backward kernels: it needs 70-80
  // registers against 84-96 at 4 columns, i.e. no more resident warps and half the bytes in flight per thread
  const int width = (BWD && g_vec_bwd_width == 2) ? 2 : 4;
  if (width == 2) DRIN_TRY((launch_vec_shape<2, BWD>(stream, a, edge, grid)));
  else DRIN_TRY((launch_vec_shape<4, BWD>(stream, a, edge, grid)));
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

int vec_layer_fwd(cudaStream_t stream, const VecLayerArgs& a) {
  prof::Scope prof_scope(stream, prof::GCN_FWD);
  return vec_layer_dispatch<false>(stream, a, nullptr);
}

// partials: [*partial_rows][3][D] with *partial_rows = the launched grid <= vec_layer_ctas()
int vec_layer_bwd(cudaStream_t stream, const VecLayerArgs& a, int* partial_rows) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  return vec_layer_dispatch<true>(stream, a, partial_rows);
}

// ---------------------------------------------------------------------------------------------
// first-layer shortcut helpers
// ---------------------------------------------------------------------------------------------
// w1[o] = sum_i W[o, i]  (W_m 1), one warp per row
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ w, int rows, int cols, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float t = 0.f;
  for (int i = lane; i < cols; i += 32) t += w[(long long)r * cols + i];
  t = warp_sum(t);
  if (lane == 0) out[r] = t;
}

int rowsum(cudaStream_t stream, const float* w, int rows, int cols, float* out) {
  prof::Scope prof_scope(stream, prof::GCN_FWD);
  if (!w || !out) return fail(DRIN_ERR_ARG, "rowsum: null argument");
  rowsum_kernel<<<(rows * 32 + 255) / 256, 256, 0, stream>>>(w, rows, cols, out);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// Gradients of the first layer's edge-update parameters that the shortcut leaves in pieces:
//   dW_m[o, i] = (i < H ? dWa[o, i] : dWb[o, i - H]) + dw1[o]          (W_m[:, :H] | W_m[:, H:] | the W_m 1 term)
//   [db_u | db_v][j] = sum_o db_m[o] W_m[o, j]                         (b_u, b_v enter only through fu, fv -> A, Bv)
__global__ void __launch_bounds__(256) wm_assemble_kernel(const float* __restrict__ dwa, const float* __restrict__ dwb,
                                                          const float* __restrict__ dw1, int D,
                                                          float* __restrict__ dw_m) {
  const int H = D / 2;
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index into dW_m
  if (i4 >= (long long)D * D / 4) return;
  const int o = (int)(i4 * 4 / D), j = (int)(i4 * 4 - (long long)o * D);
  const float4 t = j < H ? *reinterpret_cast<const float4*>(dwa + (long long)o * H + j)
                         : *reinterpret_cast<const float4*>(dwb + (long long)o * H + (j - H));
  const float w = dw1[o];
  *reinterpret_cast<float4*>(dw_m + i4 * 4) = make_float4(t.x + w, t.y + w, t.z + w, t.w + w);
}

// 32 columns x 32 row-slices per block, fixed-order combine (bit-reproducible)
__global__ void __launch_bounds__(1024) bias_matvec_kernel(const float* __restrict__ db_m, const float* __restrict__ w_m,
                                                           int D, float* __restrict__ db_u, float* __restrict__ db_v) {
  __shared__ float red[32][33];
  const int H = D / 2;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float t = 0.f;
  if (j < D)
    for (int o = slice; o < D; o += 32) t = fmaf(db_m[o], w_m[(long long)o * D + j], t);
  red[slice][lane] = t;
  __syncthreads();
  if (slice == 0 && j < D) {
    float r = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 32; ++s2) r += red[s2][lane];
    if (j < H) db_u[j] = r; else db_v[j - H] = r;
  }
}

int wm_fixup(cudaStream_t stream, const float* dwa, const float* dwb, const float* dw1, const float* db_m,
             const float* w_m, int D, float* dw_m, float* db_u, float* db_v) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  if (!dwa || !dwb || !dw1 || !db_m || !w_m || !dw_m || !db_u || !db_v) return fail(DRIN_ERR_ARG, "wm_fixup: null argument");
  if (D % 8) return fail(DRIN_ERR_ARG, "wm_fixup: D %% 8 != 0");
  const long long n4 = (long long)D * D / 4;
  wm_assemble_kernel<<<(int)((n4 + 255) / 256), 256, 0, stream>>>(dwa, dwb, dw1, D, dw_m);
  DRIN_LAUNCH_CHECK();
  bias_matvec_kernel<<<(D + 31) / 32, 1024, 0, stream>>>(db_m, w_m, D, db_u, db_v);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// vec_rows_bwd: gradient w.r.t. the activated vertices of ALL rows (mt, mi, et, ei) -> gradient planes of the
// previous layer's W_h output (through GELU + LayerNorm) or of the projection outputs (first layer).
// ---------------------------------------------------------------------------------------------
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) vec_rows_bwd_kernel(const VecRowsBwdArgs a) {
  extern __shared__ __align__(16) float s_part[];          // [4][NW][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 4 * NW * D; i += NW * 32) s_part[i] = 0.f;
  __syncthreads();
  const bool ln = a.ln_gamma != nullptr;
  const long long rows = 2 * a.B + 2 * a.BC;
  for (long long r = (long long)blockIdx.x * NW + warp; r < rows; r += (long long)gridDim.x * NW) {
    RowT<D> d;
    row_load<D>(d, a.d0 + r * D, lane);
    if (a.d1) {
      RowT<D> u;
      row_load<D>(u, a.d1 + r * D, lane);
#pragma unroll
      for (int i = 0; i < RowT<D>::NV * 4; ++i) d.v[i] += u.v[i];
    }
    if (ln) {
      RowT<D> h;
      row_load<D>(h, a.h_prev + r * D, lane);
      row_ln_gelu_bwd<D>(h, d, a.ln_gamma, a.ln_beta, s_part + (0 * NW + warp) * D, s_part + (1 * NW + warp) * D, lane);
      row_accum_smem<D>(d, s_part + (2 * NW + warp) * D, lane);
    } else {
      const int seg = r < a.B ? 0 : (r < 2 * a.B ? 1 : (r < 2 * a.B + a.BC ? 2 : 3));   // b_mt, b_mi, b_et, b_ei
      row_accum_smem<D>(d, s_part + (seg * NW + warp) * D, lane);
    }
    row_store_planes<D>(d, a.out_hi + r * D, a.out_lo ? a.out_lo + r * D : nullptr, lane);
  }
  __syncthreads();
  for (int i = tid; i < 4 * D; i += NW * 32) {
    const int v = i / D, c = i - v * D;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_part[(v * NW + w) * D + c];
    a.partials[((long long)blockIdx.x * 4 + v) * D + c] = t;
  }
}

// partials: [vec_rows_ctas()][4][D]
int vec_rows_bwd(cudaStream_t stream, const VecRowsBwdArgs& a) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  if (a.D != 768) return fail(DRIN_ERR_ARG, "vec_rows_bwd: gcn_embed_dim %d not built (768 only)", a.D);
  if (!a.d0 || !a.out_hi || !a.partials || (a.ln_gamma && (!a.ln_beta || !a.h_prev)))
    return fail(DRIN_ERR_ARG, "vec_rows_bwd: missing buffers");
  constexpr int D = 768;
  const size_t smem = (size_t)4 * VR_NW * D * sizeof(float);
  DRIN_CUDA(cudaFuncSetAttribute(vec_rows_bwd_kernel<D, VR_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vec_rows_bwd_kernel<D, VR_NW><<<VR_CTAS, VR_NW * 32, smem, stream>>>(a);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// strided_colsum_multi: out[i] = sum_{c < count} src[c * stride + i], i < n, for every queued job in ONE launch
// (blockIdx.y = job).  32 columns x 32 CTA-slices per block, fixed summation order: bit-reproducible.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) strided_colsum_multi_kernel(const StridedColsumJobs jobs) {
  __shared__ float red[32][33];
  const StridedColsumJob j = jobs.job[blockIdx.y];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float t = 0.f;
  if (i < j.n) {
#pragma unroll 4
    for (int c = slice; c < j.count; c += 32) t += j.src[(long long)c * j.stride + i];
  }
  red[slice][lane] = t;
  __syncthreads();
  if (slice == 0 && i < j.n) {
    float r = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 32; ++s2) r += red[s2][lane];
    j.out[i] = r;
  }
}

int strided_colsum_multi(cudaStream_t stream, const StridedColsumJobs& jobs) {
  if (jobs.count <= 0) return DRIN_OK;
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  int nmax = 0;
  for (int k = 0; k < jobs.count; ++k) {
    const StridedColsumJob& j = jobs.job[k];
    if (!j.src || !j.out || j.n <= 0 || j.count <= 0) return fail(DRIN_ERR_ARG, "strided_colsum_multi: bad job %d", k);
    nmax = j.n > nmax ? j.n : nmax;
  }
  strided_colsum_multi_kernel<<<dim3((nmax + 31) / 32, jobs.count), 1024, 0, stream>>>(jobs);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
