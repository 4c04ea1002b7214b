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
#include "kernels.cuh"
#include "rows.cuh"

namespace drin {

static constexpr int VEC_GRID = 148 * 4;      // persistent CTAs of D/4 threads (rows of the partial-sum buffers)
static constexpr int VR_NW = 4;               // warps per CTA of vec_rows_bwd
static constexpr int VR_CTAS = 148 * 3;

int vec_layer_ctas() { return VEC_GRID; }
int vec_rows_ctas() { return VR_CTAS; }

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float s) { return make_float4(s, s, s, s); }
__device__ __forceinline__ float4 operator+(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 operator*(const float4& a, const float4& b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 operator*(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 fma4(const float4& a, const float4& b, const float4& c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
// sigmoid from the two MUFU approximations (ex2, rcp; ~2 ulp each): 4 instructions instead of an IEEE division
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.0f + exp2_ftz(x * -1.4426950408889634f)); }
__device__ __forceinline__ float4 sigmoid4(const float4& q) {
  return make_float4(sigmoid_f(q.x), sigmoid_f(q.y), sigmoid_f(q.z), sigmoid_f(q.w));
}
// four consecutive columns -> split-bf16 planes (lo may be null: plain bf16 rounding)
__device__ __forceinline__ void st_planes4(bf16* hi, bf16* lo, long long idx, const float4& v) {
  uint32_t h0, l0, h1, l1;
  split_bf16x2(v.x, v.y, h0, l0);
  split_bf16x2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h0, h1);
  if (lo) *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l0, l1);
}

// masked edge vectors E_k (and, for the backward pass, the raw sigmoid values S_k) of candidate row r
template <bool SCALAR>
__device__ __forceinline__ void load_edges(const VecLayerArgs& a, long long BC, long long r, int col, float4 (&E)[4],
                                           float4 (&S)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (SCALAR) {
      S[k] = f4(a.e_scalar[k * BC + r]);
    } else {
      S[k] = sigmoid4(ld4(a.q_in + (r * 4 + k) * a.D + col));
    }
    E[k] = S[k] * a.en[k];
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// forward: z (A operand of the W_h GEMM) and, in a layer with an edge update, m_k = cat[fu, fv] + E_k
// ---------------------------------------------------------------------------------------------
template <int D, bool FULL, bool DYN, bool SCALAR>
__global__ void __launch_bounds__(D / 4) vec_layer_fwd_kernel(const VecLayerArgs a) {
  constexpr int H = D / 2;
  const int col = threadIdx.x * 4;
  const long long B = a.B, C = a.C, BC = B * C;
  const long long row_et = FULL ? 2 * B : B;          // first et row of this layer's z / h layout
  const float inv_c_den = (float)a.C;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const float4 mt = ld4(a.xa + b * D + col);
    const float4 mi = ld4(a.xa + (B + b) * D + col);
    float4 fu_mt = f4(0.f), fu_mi = f4(0.f);
    if (DYN && col < H) {
      fu_mt = ld4(a.fu + b * H + col);
      fu_mi = ld4(a.fu + (B + b) * H + col);
    }
    float4 amt = f4(0.f), ami = f4(0.f);
    for (long long c = 0; c < C; ++c) {
      const long long r = b * C + c;
      const float4 et = ld4(a.xa + (2 * B + r) * D + col);
      const float4 ei = ld4(a.xa + (2 * B + BC + r) * D + col);
      float4 E[4], S[4];
      load_edges<SCALAR>(a, BC, r, col, E, S);
      amt = fma4(E[0], et, fma4(E[1], ei, amt));
      if (FULL) ami = fma4(E[2], et, fma4(E[3], ei, ami));
      st_planes4(a.z_hi, a.z_lo, (row_et + r) * D + col, fma4(E[0], mt, fma4(E[2], mi, et)));
      if (FULL) st_planes4(a.z_hi, a.z_lo, (2 * B + BC + r) * D + col, fma4(E[1], mt, fma4(E[3], mi, ei)));
      if (DYN) {
        float4 v_et = f4(0.f), v_ei = f4(0.f);
        if (col >= H) {
          v_et = ld4(a.fv + r * H + (col - H));
          v_ei = ld4(a.fv + (BC + r) * H + (col - H));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 base = col < H ? (k < 2 ? fu_mt : fu_mi) : ((k & 1) ? v_ei : v_et);
          st_planes4(a.m_hi, a.m_lo, (r * 4 + k) * D + col, base + E[k]);
        }
      }
    }
    const float4 zmt = make_float4(amt.x / inv_c_den + mt.x, amt.y / inv_c_den + mt.y, amt.z / inv_c_den + mt.z,
                                   amt.w / inv_c_den + mt.w);
    st_planes4(a.z_hi, a.z_lo, b * D + col, zmt);
    if (FULL) {
      const float4 zmi = make_float4(ami.x / inv_c_den + mi.x, ami.y / inv_c_den + mi.y, ami.z / inv_c_den + mi.z,
                                     ami.w / inv_c_den + mi.w);
      st_planes4(a.z_hi, a.z_lo, (B + b) * D + col, zmi);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward of the kernel above.  In: dz (gradient w.r.t. z) and, with an edge update in this layer, dm (gradient
// w.r.t. m, = dq W_m).  Out: dxa (gradient w.r.t. the activated vertices, without the W_u / W_v paths, which are
// GEMMs), dfu / dfv planes, dq of the PREVIOUS layer's edge outputs (vector edges in), column partials.
// ---------------------------------------------------------------------------------------------
template <int D, bool FULL, bool DYN, bool SCALAR>
__global__ void __launch_bounds__(D / 4, (DYN && SCALAR) ? 4 : 3) vec_layer_bwd_kernel(const VecLayerArgs a) {
  constexpr int H = D / 2;
  const int col = threadIdx.x * 4;
  const long long B = a.B, C = a.C, BC = B * C;
  const long long row_et = FULL ? 2 * B : B;
  const float inv_c = 1.0f / (float)a.C;
  float4 p_bm = f4(0.f);        // sum of dq over this CTA's rows: b_m gradient of the previous layer
  float4 p_uv = f4(0.f);        // columns < H: b_u gradient, columns >= H: b_v gradient
  // (measured: __restrict__ copies of the pointers let the compiler hoist the next candidate's loads, which costs
  // 24 registers and two resident CTAs per SM -- slower; occupancy, not load hoisting, feeds HBM here)
  const float* xa = a.xa;
  const float* dz = a.dz;
  const float* q_in = a.q_in;
  const float* e_scalar = a.e_scalar;
  const float* dm = a.dm;
  float* dxa = a.dxa;
  bf16* dq_hi = a.dq_hi;
  bf16* dq_lo = a.dq_lo;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const float4 mt = ld4(xa + b * D + col);
    const float4 mi = ld4(xa + (B + b) * D + col);
    const float4 dzmt = ld4(dz + b * D + col);
    const float4 dzmi = FULL ? ld4(dz + (B + b) * D + col) : f4(0.f);
    const float4 smt = dzmt * inv_c, smi = dzmi * inv_c;
    float4 dmt = dzmt, dmi = dzmi;
    float4 dfu_mt = f4(0.f), dfu_mi = f4(0.f);
    for (long long c = 0; c < C; ++c) {
      const long long r = b * C + c;
      const float4 et = ld4(xa + (2 * B + r) * D + col);
      const float4 ei = ld4(xa + (2 * B + BC + r) * D + col);
      const float4 dzet = ld4(dz + (row_et + r) * D + col);
      const float4 dzei = FULL ? ld4(dz + (2 * B + BC + r) * D + col) : f4(0.f);
      float4 det = dzet, dei = dzei;
      float4 dfv_et = f4(0.f), dfv_ei = f4(0.f);
      // one edge type at a time (k = 2 u + v: u = mt | mi, v = et | ei) keeps a single edge vector live:
      //   forward   a_u += E_k x_v / C,  a_v += E_k m_u
      //   backward  dx_v += (dz_u / C) E_k,  dm_u += dz_v E_k,  dE_k = (dz_u / C) x_v + dz_v m_u
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4& s_u = k < 2 ? smt : smi;
        const float4& m_u = k < 2 ? mt : mi;
        const float4& x_v = (k & 1) ? ei : et;
        const float4& dz_v = (k & 1) ? dzei : dzet;
        float4 S;
        if (SCALAR) S = f4(e_scalar[k * BC + r]);
        else S = sigmoid4(ld4(q_in + (r * 4 + k) * D + col));
        const float4 E = S * a.en[k];
        if (k & 1) dei = fma4(s_u, E, dei); else det = fma4(s_u, E, det);
        if (k < 2) dmt = fma4(dz_v, E, dmt); else dmi = fma4(dz_v, E, dmi);
        if (DYN || !SCALAR) {
          float4 dE = fma4(s_u, x_v, dz_v * m_u);
          if (DYN) {
            const float4 dmk = ld4(dm + (r * 4 + k) * D + col);
            dE = dE + dmk;
            if (col < H) {
              if (k < 2) dfu_mt = dfu_mt + dmk; else dfu_mi = dfu_mi + dmk;
            } else {
              if (k & 1) dfv_ei = dfv_ei + dmk; else dfv_et = dfv_et + dmk;
            }
          }
          if (!SCALAR) {
            // E_k = en_k * S_k, S_k = sigmoid(q_k):  dq_k = dE_k * en_k * S_k * (1 - S_k)
            const float4 ds = make_float4(S.x * (1.f - S.x), S.y * (1.f - S.y), S.z * (1.f - S.z), S.w * (1.f - S.w));
            const float4 dq = dE * ds * a.en[k];
            st_planes4(dq_hi, dq_lo, (r * 4 + k) * D + col, dq);
            p_bm = p_bm + dq;
          }
        }
      }
      st4(dxa + (2 * B + r) * D + col, det);
      st4(dxa + (2 * B + BC + r) * D + col, dei);
      if (DYN && col >= H) {
        st_planes4(a.dfv_hi, a.dfv_lo, r * H + (col - H), dfv_et);
        st_planes4(a.dfv_hi, a.dfv_lo, (BC + r) * H + (col - H), dfv_ei);
        p_uv = p_uv + dfv_et + dfv_ei;
      }
    }
    st4(dxa + b * D + col, dmt);
    st4(dxa + (B + b) * D + col, dmi);
    if (DYN && col < H) {
      st_planes4(a.dfu_hi, a.dfu_lo, b * H + col, dfu_mt);
      st_planes4(a.dfu_hi, a.dfu_lo, (B + b) * H + col, dfu_mi);
      p_uv = p_uv + dfu_mt + dfu_mi;
    }
  }
  st4(a.partials + ((long long)blockIdx.x * 2 + 0) * D + col, p_bm);
  st4(a.partials + ((long long)blockIdx.x * 2 + 1) * D + col, p_uv);
}

template <int D, bool BWD, bool FULL, bool DYN>
static void launch_vec(cudaStream_t stream, const VecLayerArgs& a, int grid) {
  if (a.e_scalar) {
    if (BWD) vec_layer_bwd_kernel<D, FULL, DYN, true><<<grid, D / 4, 0, stream>>>(a);
    else vec_layer_fwd_kernel<D, FULL, DYN, true><<<grid, D / 4, 0, stream>>>(a);
  } else {
    if (BWD) vec_layer_bwd_kernel<D, FULL, DYN, false><<<grid, D / 4, 0, stream>>>(a);
    else vec_layer_fwd_kernel<D, FULL, DYN, false><<<grid, D / 4, 0, stream>>>(a);
  }
}

template <bool BWD>
static int vec_layer_dispatch(cudaStream_t stream, const VecLayerArgs& a) {
  if (a.D != 768) return fail(DRIN_ERR_ARG, "vector-edge GCN layer: gcn_embed_dim %d not built (768 only)", a.D);
  if (!a.xa || (!a.e_scalar && !a.q_in)) return fail(DRIN_ERR_ARG, "vector-edge GCN layer: missing inputs");
  if (a.dyn && !a.full) return fail(DRIN_ERR_ARG, "vector-edge GCN layer: an edge update needs a full layer");
  if (!BWD) {
    if (!a.z_hi || (a.dyn && (!a.fu || !a.fv || !a.m_hi)))
      return fail(DRIN_ERR_ARG, "vector-edge GCN layer forward: missing buffers");
  } else {
    if (!a.dz || !a.dxa || !a.partials || (a.dyn && (!a.dm || !a.dfu_hi || !a.dfv_hi)) || (!a.e_scalar && !a.dq_hi))
      return fail(DRIN_ERR_ARG, "vector-edge GCN layer backward: missing buffers");
  }
  const int grid = a.B < VEC_GRID ? a.B : VEC_GRID;
  if (a.full) {
    if (a.dyn) launch_vec<768, BWD, true, true>(stream, a, grid);
    else launch_vec<768, BWD, true, false>(stream, a, grid);
  } else {
    launch_vec<768, BWD, false, false>(stream, a, grid);
  }
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

int vec_layer_fwd(cudaStream_t stream, const VecLayerArgs& a) {
  prof::Scope prof_scope(stream, prof::GCN_FWD);
  return vec_layer_dispatch<false>(stream, a);
}

// partials: [min(B, vec_layer_ctas())][2][D]
int vec_layer_bwd(cudaStream_t stream, const VecLayerArgs& a, int* partial_rows) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  if (partial_rows) *partial_rows = a.B < VEC_GRID ? a.B : VEC_GRID;
  return vec_layer_dispatch<true>(stream, a);
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
