// Backward of the fused GCN stages (mirrors gcn_fwd.cu; reference semantics: autograd through
// drin/model.py:121-153,207-209).  Same warp-per-row layout; activated vertices are recomputed from
// the saved pre-LayerNorm rows instead of being stored.
//
//   score_bwd            d(cosine score) -> dL/dh of the last layer (mention + candidate rows)
//   gcn_layer_bwd        per mention: gradients of the message aggregation, the enable mask and the
//                        dynamic edge update w.r.t. candidate vertices (then through LayerNorm+GELU of
//                        the previous layer), mention vertices (partial), edges, g = fu W_v and beta
//   mention_bwd_finish   mention rows: add the W_u path, LayerNorm+GELU backward
//   dfu_finish           dfu = dg W_v^T + dbeta * b_v ; bias gradients of w_u / w_v
//   colsum_reduce        deterministic reduction of per-CTA partial column sums (bias / LayerNorm grads)
//
// Partial sums live in per-warp shared-memory slices and are reduced in a fixed order, so gradients are
// bit-reproducible run to run.
#include <stdlib.h>

#include <type_traits>

#include "kernels.cuh"
#include "pipe.cuh"
#include "rows.cuh"

namespace drin {

static constexpr int BW_NW = 4;           // warps per CTA in the backward kernels
static constexpr int BW_CTAS = 148 * 3;   // persistent grid (must match Workspace::colsum_ctas)

template <int D>
__device__ __forceinline__ void row_zero_smem(float* s, int lane) {
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) *reinterpret_cast<float4*>(s + (j * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// write the CTA's partial column sums: out[(cta * nvec + v) * D + col] = sum_w part[v][w][col]
template <int D, int NW>
__device__ __forceinline__ void flush_partials(const float* s_part, int nvec, float* out, int tid) {
  for (int i = tid; i < nvec * D; i += NW * 32) {
    const int v = i / D, col = i - v * D;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_part[(v * NW + w) * D + col];
    out[((long long)blockIdx.x * nvec + v) * D + col] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// score_bwd
// ---------------------------------------------------------------------------------------------
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) score_bwd_kernel(const ScoreBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_gamma = sm;
  float* s_beta = s_gamma + D;
  float* s_m = s_beta + D;                 // activated mention row
  float* s_dam = s_m + D;                  // [NW][D] per-warp dL/da_m (vector part)
  float* s_part = s_dam + NW * D;          // [3][NW][D]: dgamma, dbeta, db_h
  __shared__ float s_mnorm;
  __shared__ float s_coef[NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += NW * 32) {
    s_gamma[i] = a.gamma[i];
    s_beta[i] = a.beta[i];
  }
  for (int i = tid; i < 3 * NW * D; i += NW * 32) s_part[i] = 0.f;
  float* pg = s_part + (0 * NW + warp) * D;
  float* pb = s_part + (1 * NW + warp) * D;
  float* ph = s_part + (2 * NW + warp) * D;
  const long long B = a.B;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();
    if (warp == 0) {
      RowT<D> m;
      row_load<D>(m, a.h_mt + (long long)b * D, lane);
      row_ln_gelu<D>(m, s_gamma, s_beta, lane);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < RowT<D>::NV * 4; ++i) q += m.v[i] * m.v[i];
      q = warp_sum(q);
      row_store<D>(m, s_m, lane);
      if (lane == 0) s_mnorm = fmaxf(sqrtf(q), 1e-8f);
    }
    row_zero_smem<D>(s_dam + warp * D, lane);
    float coef = 0.f;
    __syncthreads();
    const float nm = s_mnorm;
    for (int c = warp; c < a.C; c += NW) {
      const long long r = (long long)b * a.C + c;
      RowT<D> h, e, de;
      row_load<D>(h, a.h_et + r * D, lane);
      const float rstd = row_ln_gelu_recompute<D>(h, e, de, s_gamma, s_beta, lane);   // h := xhat, e := activated
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < RowT<D>::NV * 4; ++i) q += e.v[i] * e.v[i];
      q = warp_sum(q);
      const float ne = fmaxf(sqrtf(q), 1e-8f);
      const float cs = warp_sum(row_dot<D>(e, s_m, lane)) / (nm * ne);
      const float ds = a.dscores[r];
      const float w1 = ds / (nm * ne), w2 = ds * cs / (ne * ne);
      coef += ds * cs / (nm * nm);
      // dL/da_e = w1 * a_m - w2 * a_e ; dL/da_m += w1 * a_e
#pragma unroll
      for (int j = 0; j < RowT<D>::NV; ++j) {
        const int off = (j * 32 + lane) * 4;
        const float4 mv = *reinterpret_cast<const float4*>(s_m + off);
        float4 acc = *reinterpret_cast<float4*>(s_dam + warp * D + off);
        acc.x += w1 * e.v[4 * j]; acc.y += w1 * e.v[4 * j + 1]; acc.z += w1 * e.v[4 * j + 2]; acc.w += w1 * e.v[4 * j + 3];
        *reinterpret_cast<float4*>(s_dam + warp * D + off) = acc;
        e.v[4 * j] = w1 * mv.x - w2 * e.v[4 * j];
        e.v[4 * j + 1] = w1 * mv.y - w2 * e.v[4 * j + 1];
        e.v[4 * j + 2] = w1 * mv.z - w2 * e.v[4 * j + 2];
        e.v[4 * j + 3] = w1 * mv.w - w2 * e.v[4 * j + 3];
      }
      row_ln_gelu_bwd_from<D>(h, de, rstd, e, s_gamma, pg, pb, lane);
      row_accum_smem<D>(e, ph, lane);
      const long long zr = B + r;
      row_store_planes<D>(e, a.dh_hi + zr * D, a.dh_lo ? a.dh_lo + zr * D : nullptr, lane);
    }
    if (lane == 0) s_coef[warp] = coef;
    __syncthreads();
    if (warp == 0) {
      float ctot = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) ctot += s_coef[w];
      RowT<D> d, h;
#pragma unroll
      for (int j = 0; j < RowT<D>::NV; ++j) {
        const int off = (j * 32 + lane) * 4;
        const float4 mv = *reinterpret_cast<const float4*>(s_m + off);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const float4 u = *reinterpret_cast<const float4*>(s_dam + w * D + off);
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        d.v[4 * j] = t.x - ctot * mv.x;
        d.v[4 * j + 1] = t.y - ctot * mv.y;
        d.v[4 * j + 2] = t.z - ctot * mv.z;
        d.v[4 * j + 3] = t.w - ctot * mv.w;
      }
      row_load<D>(h, a.h_mt + (long long)b * D, lane);
      row_ln_gelu_bwd<D>(h, d, s_gamma, s_beta, pg, pb, lane);
      row_accum_smem<D>(d, ph, lane);
      row_store_planes<D>(d, a.dh_hi + (long long)b * D, a.dh_lo ? a.dh_lo + (long long)b * D : nullptr, lane);
    }
  }
  __syncthreads();
  flush_partials<D, NW>(s_part, 3, a.partials, tid);
}


// ---------------------------------------------------------------------------------------------
// score_bwd, warp-autonomous version (default when the batch has enough mentions to fill the machine)
// ---------------------------------------------------------------------------------------------
// One warp owns one mention at a time: its mention row, then its C candidate rows in order.  Everything that is
// summed over rows lives in REGISTERS of that warp -- the mention's dL/da_m over its candidates, and the kernel-long
// partial column sums (dgamma, dbeta, db_h) -- so the row loop has no CTA barrier, no shared-memory read-modify-write
// and no atomics; the next candidate row is prefetched into registers while the current one is processed.  Partial
// sums are combined once at the end of the kernel in a fixed order (bit-reproducible).
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32, 1) score_bwd_warp_kernel(const ScoreBwdArgs a, int partial_rows) {
  constexpr int NV = RowT<D>::NV, NE = NV * 4;
  extern __shared__ __align__(16) float sm[];
  float* s_gamma = sm;
  float* s_beta = s_gamma + D;
  float* s_am = s_beta + D;                // [NW][D] activated mention row of each warp's current mention
  float* s_red = s_am + NW * D;            // [NW][3][D] end-of-kernel partials; during the row loop the first two
                                           // vectors of a warp's slice park xhat / gelu' of its mention row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += NW * 32) {
    s_gamma[i] = a.gamma[i];
    s_beta[i] = a.beta[i];
  }
  __syncthreads();
  float* am = s_am + warp * D;
  float* park = s_red + warp * 3 * D;
  const long long B = a.B;
  float pg[NE], pb[NE], ph[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) pg[i] = pb[i] = ph[i] = 0.f;

  // dL/dh of one row from dL/da (in d), given xhat / dact / rstd; accumulates the column partials in registers
  auto ln_bwd = [&](const RowT<D>& xhat, const RowT<D>& dact, float rstd, RowT<D>& d) {
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 g = *reinterpret_cast<const float4*>(s_gamma + (j * 32 + lane) * 4);
      const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = 4 * j + k;
        const float dy = d.v[i] * dact.v[i];
        pg[i] = fmaf(dy, xhat.v[i], pg[i]);
        pb[i] += dy;
        const float dxh = dy * gg[k];
        d.v[i] = dxh;
        m1 += dxh;
        m2 = fmaf(dxh, xhat.v[i], m2);
      }
    }
    m1 = warp_sum(m1) * (1.0f / D);
    m2 = warp_sum(m2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      d.v[i] = rstd * (d.v[i] - m1 - xhat.v[i] * m2);
      ph[i] += d.v[i];
    }
  };

  const long long gwarp = (long long)blockIdx.x * NW + warp, nwarps = (long long)gridDim.x * NW;
  const int S = a.slices, cps = (a.C + S - 1) / S;        // work unit = (mention, candidate slice)
  const long long units = B * S;
  for (long long u = gwarp; u < units; u += nwarps) {
    const long long b = u / S;
    const int c0 = (int)(u - b * S) * cps, c1 = min(a.C, c0 + cps);
    if (c0 >= a.C) {                                      // empty trailing slice
      float* part = a.slice_part + u * (D + 32);
      for (int i = lane; i < D + 32; i += 32) part[i] = 0.f;
      continue;
    }
    // ---- mention row: activated vertex -> smem (read back as float4 by the candidate rows), norm
    RowT<D> xm, actm, dactm;
    row_load<D>(xm, a.h_mt + b * D, lane);
    RowT<D> hn;                                           // prefetched candidate row
    row_load<D>(hn, a.h_et + (b * a.C + c0) * D, lane);
    float ds_n = a.dscores[b * a.C + c0];
    const float rstd_m = row_ln_gelu_recompute<D>(xm, actm, dactm, s_gamma, s_beta, lane);   // xm := xhat
    float qm = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) qm = fmaf(actm.v[i], actm.v[i], qm);
    qm = warp_sum(qm);
    const float nm = fmaxf(sqrtf(qm), 1e-8f);
    row_store<D>(actm, am, lane);
    row_store<D>(xm, park, lane);                          // keep the mention row out of registers during the loop
    row_store<D>(dactm, park + D, lane);
    __syncwarp();
    RowT<D> dam;
#pragma unroll
    for (int i = 0; i < NE; ++i) dam.v[i] = 0.f;
    float coef = 0.f;
    for (int c = c0; c < c1; ++c) {
      const long long r = b * a.C + c;
      RowT<D> h = hn, e, de;
      const float ds = ds_n;
      if (c + 1 < c1) {
        row_load<D>(hn, a.h_et + (r + 1) * D, lane);
        ds_n = a.dscores[r + 1];
      }
      if (lane == 0) {        // DRAM latency runs ahead of the register prefetch (bulk L2 prefetch)
        long long pr = -1;
        if (c + 3 < c1) {
          pr = r + 3;
        } else if (u + nwarps < units) {
          const long long nb = (u + nwarps) / S;
          pr = nb * a.C + min((int)(u + nwarps - nb * S) * cps + (c + 3 - c1), a.C - 1);
        }
        if (pr >= 0) l2_prefetch(a.h_et + pr * D, D * 4);
      }
      asm volatile("" ::: "memory");                         // keeps the prefetch and the row's math apart (registers)
      const float rstd = row_ln_gelu_recompute<D>(h, e, de, s_gamma, s_beta, lane);          // h := xhat, e := activated
      float q = 0.f, dot = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 mv = *reinterpret_cast<const float4*>(am + (j * 32 + lane) * 4);
        q = fmaf(e.v[4 * j], e.v[4 * j], q); q = fmaf(e.v[4 * j + 1], e.v[4 * j + 1], q);
        q = fmaf(e.v[4 * j + 2], e.v[4 * j + 2], q); q = fmaf(e.v[4 * j + 3], e.v[4 * j + 3], q);
        dot = fmaf(e.v[4 * j], mv.x, dot); dot = fmaf(e.v[4 * j + 1], mv.y, dot);
        dot = fmaf(e.v[4 * j + 2], mv.z, dot); dot = fmaf(e.v[4 * j + 3], mv.w, dot);
      }
      q = warp_sum(q);
      dot = warp_sum(dot);
      const float ne = fmaxf(sqrtf(q), 1e-8f);
      const float cs = dot / (nm * ne);
      const float w1 = ds / (nm * ne), w2 = ds * cs / (ne * ne);
      coef += ds * cs / (nm * nm);
      asm volatile("" ::: "memory");
      // dL/da_e = w1 * a_m - w2 * a_e ; dL/da_m += w1 * a_e
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 mv = *reinterpret_cast<const float4*>(am + (j * 32 + lane) * 4);
        const float mm[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 4 * j + k;
          dam.v[i] = fmaf(w1, e.v[i], dam.v[i]);
          e.v[i] = w1 * mm[k] - w2 * e.v[i];
        }
      }
      asm volatile("" ::: "memory");
      ln_bwd(h, de, rstd, e);
      asm volatile("" ::: "memory");
      const long long zr = B + r;
      row_store_planes<D>(e, a.dh_hi + zr * D, a.dh_lo ? a.dh_lo + zr * D : nullptr, lane);
    }
    if (S > 1) {
      // sliced mention: leave the slice's share of dL/da_m to score_bwd_mention_finish (fixed slice order)
      float* part = a.slice_part + u * (D + 32);
      row_store<D>(dam, part, lane);
      if (lane == 0) part[D] = coef;
      __syncwarp();
      continue;
    }
    // ---- mention row gradient: dL/da_m = sum_c w1_c a_e_c - (sum_c ds_c cs_c / nm^2) a_m
    {
      RowT<D> t;
      row_load<D>(t, am, lane);
#pragma unroll
      for (int i = 0; i < NE; ++i) dam.v[i] = fmaf(-coef, t.v[i], dam.v[i]);
    }
    {
      RowT<D> xh, da;
      row_load<D>(xh, park, lane);
      row_load<D>(da, park + D, lane);
      ln_bwd(xh, da, rstd_m, dam);
    }
    row_store_planes<D>(dam, a.dh_hi + b * D, a.dh_lo ? a.dh_lo + b * D : nullptr, lane);
    __syncwarp();                                          // am is rewritten by the next mention
  }

  // ---- fixed-order combination of the per-warp partials; unused rows of the partial buffer are zeroed
  __syncwarp();
  float* mine = s_red + warp * 3 * D;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int off = (j * 32 + lane) * 4;
    *reinterpret_cast<float4*>(mine + off) = make_float4(pg[4 * j], pg[4 * j + 1], pg[4 * j + 2], pg[4 * j + 3]);
    *reinterpret_cast<float4*>(mine + D + off) = make_float4(pb[4 * j], pb[4 * j + 1], pb[4 * j + 2], pb[4 * j + 3]);
    *reinterpret_cast<float4*>(mine + 2 * D + off) = make_float4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
  }
  __syncthreads();
  for (int i = tid; i < 3 * D; i += NW * 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_red[w * 3 * D + i];
    a.partials[(long long)blockIdx.x * 3 * D + i] = t;
    for (int extra = blockIdx.x + gridDim.x; extra < partial_rows; extra += gridDim.x)
      a.partials[(long long)extra * 3 * D + i] = 0.f;
  }
}

// sliced mentions: dL/da_m = sum over slices of the partial sums - (sum of coefficients) * a_m, then through
// LayerNorm + GELU of the mention row; column partials go to their own buffer (second source of the reduction)
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) score_bwd_mention_finish_kernel(const ScoreBwdArgs a) {
  constexpr int NE = RowT<D>::NV * 4;
  extern __shared__ __align__(16) float sm[];
  float* s_gamma = sm;
  float* s_beta = s_gamma + D;
  float* s_part = s_beta + D;              // [3][NW][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += NW * 32) {
    s_gamma[i] = a.gamma[i];
    s_beta[i] = a.beta[i];
  }
  for (int i = tid; i < 3 * NW * D; i += NW * 32) s_part[i] = 0.f;
  __syncthreads();
  float* pg = s_part + (0 * NW + warp) * D;
  float* pb = s_part + (1 * NW + warp) * D;
  float* ph = s_part + (2 * NW + warp) * D;
  for (long long b = (long long)blockIdx.x * NW + warp; b < a.B; b += (long long)gridDim.x * NW) {
    RowT<D> h, act, dact, d, t;
    row_load<D>(h, a.h_mt + b * D, lane);
    RowT<D> xhat = h;
    const float rstd = row_ln_gelu_recompute<D>(xhat, act, dact, s_gamma, s_beta, lane);
#pragma unroll
    for (int i = 0; i < NE; ++i) d.v[i] = 0.f;
    float coef = 0.f;
    for (int sl = 0; sl < a.slices; ++sl) {
      const float* part = a.slice_part + (b * a.slices + sl) * (D + 32);
      row_load<D>(t, part, lane);
#pragma unroll
      for (int i = 0; i < NE; ++i) d.v[i] += t.v[i];
      coef += part[D];
    }
#pragma unroll
    for (int i = 0; i < NE; ++i) d.v[i] = fmaf(-coef, act.v[i], d.v[i]);
    row_ln_gelu_bwd_from<D>(xhat, dact, rstd, d, s_gamma, pg, pb, lane);
    row_accum_smem<D>(d, ph, lane);
    row_store_planes<D>(d, a.dh_hi + b * D, a.dh_lo ? a.dh_lo + b * D : nullptr, lane);
  }
  __syncthreads();
  flush_partials<D, NW>(s_part, 3, a.partials2, tid);
}

static int g_score_bwd_variant = -1;     // -1 auto, 0 CTA-per-mention kernel, 1 warp-autonomous kernel (test hook)
void debug_set_score_bwd_variant(int v) { g_score_bwd_variant = v; }

int score_bwd(cudaStream_t stream, const ScoreBwdArgs& a_in, bool* used_partials2) {
  prof::Scope prof_scope(stream, prof::SCORE);
  ScoreBwdArgs a = a_in;
  if (a.slices < 1 || !a.slice_part || !a.partials2) a.slices = 1;
  if (used_partials2) *used_partials2 = false;
  if (a.D != 768) return fail(DRIN_ERR_ARG, "score_bwd: gcn_embed_dim %d not built (768 only)", a.D);
  constexpr int D = 768;
  constexpr int WNW = 8, WGRID = 148;
  const bool warp_kernel = g_score_bwd_variant < 0 ? (long long)a.B * a.slices >= WGRID * WNW : g_score_bwd_variant >= 1;
  if (warp_kernel) {            // enough (mention, slice) units for one per warp: barrier-free register-accumulating kernel
    const size_t smem = (size_t)(2 + WNW + 3 * WNW) * D * sizeof(float);
    DRIN_CUDA(cudaFuncSetAttribute(score_bwd_warp_kernel<D, WNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_bwd_warp_kernel<D, WNW><<<WGRID, WNW * 32, smem, stream>>>(a, BW_CTAS);
    DRIN_LAUNCH_CHECK();
    if (a.slices > 1) {
      const size_t fsmem = (size_t)(2 + 3 * BW_NW) * D * sizeof(float);
      score_bwd_mention_finish_kernel<D, BW_NW><<<BW_CTAS, BW_NW * 32, fsmem, stream>>>(a);
      DRIN_LAUNCH_CHECK();
      if (used_partials2) *used_partials2 = true;
    }
    return DRIN_OK;
  }
  a.slices = 1;
  const size_t smem = (size_t)(3 + BW_NW + 3 * BW_NW) * D * sizeof(float);
  DRIN_CUDA(cudaFuncSetAttribute(score_bwd_kernel<D, BW_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  score_bwd_kernel<D, BW_NW><<<BW_CTAS, BW_NW * 32, smem, stream>>>(a);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// gcn_layer_bwd (shared-memory staged)
// ---------------------------------------------------------------------------------------------
// A producer warp streams chunks of BS_CH candidates (vertex rows + their dz rows + the mention-side
// vectors) into a 2-stage ring with bulk async copies; BS_NW consumer warps alternate between
//   column phases : thread t owns columns t and t + 384; all sums over candidates (messages to the mention
//                   vertices, g gradients, bias / LayerNorm parameter gradients) are accumulated in REGISTERS in
//                   candidate order -> no per-warp shared-memory slices, deterministic, and
//   a row phase   : one warp per vertex row (LayerNorm/GELU recompute, edge-gradient dot products, dL/dx,
//                   LayerNorm backward, gradient planes), which leaves xhat / dy (or dL/dx) in the ring for the
//                   second column phase.
// HBM latency is hidden by the ring, and the small per-warp state allows 12 consumer warps per SM.
static constexpr int BS_NW = 12;
static constexpr int BS_CH = 6;
static constexpr int BS_CONS = BS_NW * 32;            // 384 consumer threads = D / 2
static constexpr int BS_THREADS = BS_CONS + 32;
static constexpr int BS_STAGES = 2;
static constexpr int BS_GRID = 148;                   // one persistent CTA per SM

template <int D, bool FULL>
__global__ void __launch_bounds__(BS_THREADS, 1) gcn_layer_bwd_stream_kernel(const LayerBwdArgs a) {
  static_assert(D == 2 * BS_CONS, "column ownership assumes D = 2 * consumer threads");
  extern __shared__ __align__(128) float sm[];
  constexpr int STAGE_FLOATS = (4 * BS_CH + 6) * D;
  float* s_gamma = sm + BS_STAGES * STAGE_FLOATS;
  float* s_beta = s_gamma + D;
  float* s_sc = s_beta + D;                            // [2][CH][8]  e0..e3, ds0..ds3
  float* s_dots = s_sc + 2 * BS_CH * 8;                // [CH][2][4]  pd_mt, pd_mi, q_mt, q_mi per row
  float* s_stat = s_dots + BS_CH * 8;                  // [CH][2][4]  rstd, m1, m2
  __shared__ __align__(8) unsigned long long bars[2 * BS_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long B = a.B, BC = (long long)a.B * a.C;
  const bool ln = a.ln_gamma != nullptr;
  const bool dyn = FULL && a.g != nullptr;               // dynamic edge update in this layer (false: static edges)
  const int nchunks = (a.C + BS_CH - 1) / BS_CH;
  auto full_bar = [&](int st) { return smem_u32(&bars[st]); };
  auto empty_bar = [&](int st) { return smem_u32(&bars[BS_STAGES + st]); };
  if (tid == 0) {
    for (int st = 0; st < BS_STAGES; ++st) {
      mbar_init(full_bar(st), 1);
      mbar_init(empty_bar(st), BS_NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (ln) {
    for (int i = tid; i < D; i += BS_THREADS) {
      s_gamma[i] = a.ln_gamma[i];
      s_beta[i] = a.ln_beta[i];
    }
  }
  __syncthreads();
  const float* dz_mt = a.dz;
  const float* dz_mi = FULL ? a.dz + B * D : nullptr;
  const float* dz_et = a.dz + (FULL ? 2 * B : B) * D;
  const float* dz_ei = FULL ? a.dz + (2 * B + BC) * D : nullptr;

  if (warp == BS_NW) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        for (int k = 0; k < nchunks; ++k) {
          const int n = min(BS_CH, a.C - k * BS_CH);
          const long long r0 = (long long)b * a.C + k * BS_CH;
          mbar_wait(empty_bar(stage), phase ^ 1u, nullptr, 0);
          const uint32_t rb = (uint32_t)(n * D * sizeof(float)), vb = D * (uint32_t)sizeof(float);
          mbar_arrive_expect_tx(full_bar(stage), (FULL ? 4u : 3u) * rb + (dyn ? 6u : (FULL ? 4u : 3u)) * vb);
          const uint32_t base = smem_u32(sm + stage * STAGE_FLOATS);
          const uint32_t arr = BS_CH * D * 4;
          bulk_copy_g2s(base, a.x_et + r0 * D, rb, full_bar(stage));
          bulk_copy_g2s(base + arr, a.x_ei + r0 * D, rb, full_bar(stage));
          bulk_copy_g2s(base + 2 * arr, dz_et + r0 * D, rb, full_bar(stage));
          if (FULL) bulk_copy_g2s(base + 3 * arr, dz_ei + r0 * D, rb, full_bar(stage));
          const uint32_t v0 = base + 4 * arr;
          bulk_copy_g2s(v0, a.xm + (long long)b * D, vb, full_bar(stage));
          bulk_copy_g2s(v0 + vb, a.xm + (B + b) * D, vb, full_bar(stage));
          bulk_copy_g2s(v0 + 2 * vb, dz_mt + (long long)b * D, vb, full_bar(stage));
          if (FULL) bulk_copy_g2s(v0 + 3 * vb, dz_mi + (long long)b * D, vb, full_bar(stage));
          if (dyn) {
            bulk_copy_g2s(v0 + 4 * vb, a.g + (long long)b * D, vb, full_bar(stage));
            bulk_copy_g2s(v0 + 5 * vb, a.g + (B + b) * D, vb, full_bar(stage));
          }
          if (++stage == BS_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  const float invC = 1.0f / (float)a.C, invD = 1.0f / (float)D;
  const int col0 = tid, col1 = tid + BS_CONS;
  float part[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};   // per-thread column partial sums for the whole kernel
  int stage = 0;
  uint32_t phase = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float A_mt[2] = {0.f, 0.f}, A_mi[2] = {0.f, 0.f}, G_mt[2] = {0.f, 0.f}, G_mi[2] = {0.f, 0.f};
    float dbeta_mt = 0.f, dbeta_mi = 0.f;
    for (int k = 0; k < nchunks; ++k) {
      const int n = min(BS_CH, a.C - k * BS_CH);
      const long long r0 = (long long)b * a.C + k * BS_CH;
      float* st = sm + stage * STAGE_FLOATS;
      float* Xet = st;
      float* Xei = st + BS_CH * D;
      float* DZet = st + 2 * BS_CH * D;
      float* DZei = st + 3 * BS_CH * D;                  // full: dz of the entity-image rows; else scratch
      const float* s_xmt = st + 4 * BS_CH * D;
      const float* s_xmi = s_xmt + D;
      const float* s_dzmt = s_xmi + D;
      const float* s_dzmi = s_dzmt + D;
      const float* s_gmt = s_dzmi + D;
      const float* s_gmi = s_gmt + D;
      float* sc = s_sc + stage * BS_CH * 8;
      // 1. per-candidate scalars (enable mask model.py:122, sigmoid backward of the edge update)
      if (tid < n) {
        const long long r = r0 + tid;
        sc[tid * 8 + 0] = a.edges_in[r] * a.en[0];
        sc[tid * 8 + 1] = a.edges_in[BC + r] * a.en[1];
        sc[tid * 8 + 2] = a.edges_in[2 * BC + r] * a.en[2];
        sc[tid * 8 + 3] = a.edges_in[3 * BC + r] * a.en[3];
        float ds[4] = {0.f, 0.f, 0.f, 0.f};
        if (dyn) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float o = a.edges_out[q * BC + r];
            ds[q] = a.dedges_out[q * BC + r] * o * (1.f - o);
          }
          dbeta_mt += (ds[0] + ds[1]) * invD;
          dbeta_mi += (ds[2] + ds[3]) * invD;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sc[tid * 8 + 4 + q] = ds[q];
      }
      mbar_wait(full_bar(stage), phase, nullptr, 0);
      named_bar_sync(1, BS_CONS);
      // 2. column phase 1: sums over candidates that need the ORIGINAL dz / x rows
      for (int c = 0; c < n; ++c) {
        const float e0 = sc[c * 8], e1 = sc[c * 8 + 1], e2 = sc[c * 8 + 2], e3 = sc[c * 8 + 3];
        const float dze0 = DZet[c * D + col0], dze1 = DZet[c * D + col1];
        A_mt[0] += e0 * dze0; A_mt[1] += e0 * dze1;
        A_mi[0] += e2 * dze0; A_mi[1] += e2 * dze1;
        if (FULL) {
          const float dzi0 = DZei[c * D + col0], dzi1 = DZei[c * D + col1];
          A_mt[0] += e1 * dzi0; A_mt[1] += e1 * dzi1;
          A_mi[0] += e3 * dzi0; A_mi[1] += e3 * dzi1;
          const float d0 = sc[c * 8 + 4] * invD, d1 = sc[c * 8 + 5] * invD, d2 = sc[c * 8 + 6] * invD, d3 = sc[c * 8 + 7] * invD;
          if (!ln && dyn) {                                                            // layer-0 rows are already activated;
            const float xe0 = Xet[c * D + col0], xe1 = Xet[c * D + col1];       // with a LayerNorm in front the
            const float xi0 = Xei[c * D + col0], xi1 = Xei[c * D + col1];       // g sums wait for column phase 2
            G_mt[0] += d0 * xe0 + d1 * xi0; G_mt[1] += d0 * xe1 + d1 * xi1;
            G_mi[0] += d2 * xe0 + d3 * xi0; G_mi[1] += d2 * xe1 + d3 * xi1;
          }
        }
      }
      named_bar_sync(1, BS_CONS);
      // 3. row phase: one warp per vertex row (et rows first, then ei rows)
      for (int q = warp; q < 2 * n; q += BS_NW) {
        const int kind = q >= n, c = kind ? q - n : q;
        float* xrow = (kind ? Xei : Xet) + c * D;
        float* drow = (kind ? DZei : DZet) + c * D;
        const bool has_dz = FULL || !kind;
        const float e_mt = sc[c * 8 + (kind ? 1 : 0)], e_mi = sc[c * 8 + (kind ? 3 : 2)];
        const float ds_mt = sc[c * 8 + 4 + (kind ? 1 : 0)], ds_mi = sc[c * 8 + 4 + (kind ? 3 : 2)];
        RowT<D> x, d, xhat, dact;
        row_load<D>(x, xrow, lane);
        if (has_dz) {
          row_load<D>(d, drow, lane);
        } else {
#pragma unroll
          for (int i = 0; i < RowT<D>::NV * 4; ++i) d.v[i] = 0.f;
        }
        float rstd = 0.f;
        if (ln) {
          xhat = x;
          rstd = row_ln_gelu_recompute<D>(xhat, x, dact, s_gamma, s_beta, lane);     // x := activated row
        }
        float pd_mt = row_dot<D>(x, s_dzmt, lane);
        float pd_mi = FULL ? row_dot<D>(x, s_dzmi, lane) : 0.f;
        float q_mt = has_dz ? row_dot<D>(d, s_xmt, lane) : 0.f;
        float q_mi = has_dz ? row_dot<D>(d, s_xmi, lane) : 0.f;
        pd_mt = warp_sum(pd_mt); q_mt = warp_sum(q_mt); q_mi = warp_sum(q_mi);
        if (FULL) pd_mi = warp_sum(pd_mi);
        if (lane == 0) {
          float* dd = s_dots + (c * 2 + kind) * 4;
          dd[0] = pd_mt; dd[1] = pd_mi; dd[2] = q_mt; dd[3] = q_mi;
        }
        const float ec_mt = e_mt * invC, ec_mi = e_mi * invC, dd_mt = ds_mt * invD, dd_mi = ds_mi * invD;
#pragma unroll
        for (int j = 0; j < RowT<D>::NV; ++j) {
          const int off = (j * 32 + lane) * 4;
          const float4 zmt = *reinterpret_cast<const float4*>(s_dzmt + off);
          d.v[4 * j] += ec_mt * zmt.x; d.v[4 * j + 1] += ec_mt * zmt.y; d.v[4 * j + 2] += ec_mt * zmt.z; d.v[4 * j + 3] += ec_mt * zmt.w;
          if (FULL) {
            const float4 zmi = *reinterpret_cast<const float4*>(s_dzmi + off);
            d.v[4 * j] += ec_mi * zmi.x; d.v[4 * j + 1] += ec_mi * zmi.y; d.v[4 * j + 2] += ec_mi * zmi.z; d.v[4 * j + 3] += ec_mi * zmi.w;
            if (dyn) {
              const float4 gmt = *reinterpret_cast<const float4*>(s_gmt + off);
              const float4 gmi = *reinterpret_cast<const float4*>(s_gmi + off);
              d.v[4 * j] += dd_mt * gmt.x + dd_mi * gmi.x;
              d.v[4 * j + 1] += dd_mt * gmt.y + dd_mi * gmi.y;
              d.v[4 * j + 2] += dd_mt * gmt.z + dd_mi * gmi.z;
              d.v[4 * j + 3] += dd_mt * gmt.w + dd_mi * gmi.w;
            }
          }
        }
        const long long orow = (kind ? 2 * B + BC : 2 * B) + r0 + c;      // row in the [mt; mi; et; ei] layout
        if (ln) {
          // dy = dL/da * gelu'(y); keep xhat and dy in the ring for the parameter-gradient column phase
          float m1 = 0.f, m2 = 0.f;
#pragma unroll
          for (int j = 0; j < RowT<D>::NV; ++j) {
            const int off = (j * 32 + lane) * 4;
            const float4 g = *reinterpret_cast<const float4*>(s_gamma + off);
            const float gg[4] = {g.x, g.y, g.z, g.w};
            float4 dy4;
            float* dyp = reinterpret_cast<float*>(&dy4);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float dy = d.v[4 * j + t] * dact.v[4 * j + t];
              dyp[t] = dy;
              const float dxh = dy * gg[t];
              d.v[4 * j + t] = dxh;
              m1 += dxh;
              m2 += dxh * xhat.v[4 * j + t];
            }
            *reinterpret_cast<float4*>(drow + off) = dy4;
            *reinterpret_cast<float4*>(xrow + off) =
                make_float4(xhat.v[4 * j], xhat.v[4 * j + 1], xhat.v[4 * j + 2], xhat.v[4 * j + 3]);
          }
          m1 = warp_sum(m1) * (1.0f / D);
          m2 = warp_sum(m2) * (1.0f / D);
#pragma unroll
          for (int i = 0; i < RowT<D>::NV * 4; ++i) d.v[i] = rstd * (d.v[i] - m1 - xhat.v[i] * m2);
          if (lane == 0) {
            float* ss = s_stat + (c * 2 + kind) * 4;
            ss[0] = rstd; ss[1] = m1; ss[2] = m2;
          }
        } else {
          row_store<D>(d, drow, lane);                    // dL/dx stays in the ring for the bias-gradient sums
        }
        row_store_planes<D>(d, a.dcand_hi + orow * D, a.dcand_lo ? a.dcand_lo + orow * D : nullptr, lane);
      }
      named_bar_sync(1, BS_CONS);
      // 4. edge gradients (the enable mask is applied again on the way back) and column phase 2
      if (a.dedges_in && tid < n) {
        const float* de = s_dots + (tid * 2 + 0) * 4;     // et row
        const float* di = s_dots + (tid * 2 + 1) * 4;     // ei row
        const long long r = r0 + tid;
        a.dedges_in[r] = (de[0] * invC + de[2] + sc[tid * 8 + 4]) * a.en[0];
        a.dedges_in[BC + r] = (di[0] * invC + di[2] + sc[tid * 8 + 5]) * a.en[1];
        a.dedges_in[2 * BC + r] = (de[1] * invC + de[3] + sc[tid * 8 + 6]) * a.en[2];
        a.dedges_in[3 * BC + r] = (di[1] * invC + di[3] + sc[tid * 8 + 7]) * a.en[3];
      }
      if (ln) {
        const float g0 = s_gamma[col0], g1 = s_gamma[col1];
        for (int c = 0; c < n; ++c) {
#pragma unroll
          for (int kind = 0; kind < 2; ++kind) {
            const float* xr = (kind ? Xei : Xet) + c * D;
            const float* dr = (kind ? DZei : DZet) + c * D;
            const float* ss = s_stat + (c * 2 + kind) * 4;
            const float rstd = ss[0], m1 = ss[1], m2 = ss[2];
            const float dy0 = dr[col0], dy1 = dr[col1], xh0 = xr[col0], xh1 = xr[col1];
            part[0][0] += dy0 * xh0; part[0][1] += dy1 * xh1;                       // dgamma
            part[1][0] += dy0; part[1][1] += dy1;                                   // dbeta (LayerNorm)
            part[2][0] += rstd * (g0 * dy0 - m1 - xh0 * m2);                        // db_h = sum dL/dh
            part[2][1] += rstd * (g1 * dy1 - m1 - xh1 * m2);
            if (dyn) {    // middle layers (L >= 3): g sums need the activated vertex, recomputed from xhat
              const float a0 = gelu_f(fmaf(xh0, g0, s_beta[col0])), a1 = gelu_f(fmaf(xh1, g1, s_beta[col1]));
              const float dm = sc[c * 8 + 4 + kind] * invD, di = sc[c * 8 + 6 + kind] * invD;   // edges (mt,row), (mi,row)
              G_mt[0] += dm * a0; G_mt[1] += dm * a1;
              G_mi[0] += di * a0; G_mi[1] += di * a1;
            }
          }
        }
      } else {
        for (int c = 0; c < n; ++c) {
          part[0][0] += DZet[c * D + col0]; part[0][1] += DZet[c * D + col1];       // db_et
          part[1][0] += DZei[c * D + col0]; part[1][1] += DZei[c * D + col1];       // db_ei
        }
      }
      if (k == nchunks - 1) {
        // 5. mention-side results: dxm = dz_m + sum_c(...) ; dg ; dbeta
        a.dxm[(long long)b * D + col0] = s_dzmt[col0] + A_mt[0];
        a.dxm[(long long)b * D + col1] = s_dzmt[col1] + A_mt[1];
        a.dxm[(B + b) * D + col0] = (FULL ? s_dzmi[col0] : 0.f) + A_mi[0];
        a.dxm[(B + b) * D + col1] = (FULL ? s_dzmi[col1] : 0.f) + A_mi[1];
        if (dyn) {
          bf16 h, l;
          split_bf16(G_mt[0], h, l);
          a.dg_hi[(long long)b * D + col0] = h; if (a.dg_lo) a.dg_lo[(long long)b * D + col0] = l;
          split_bf16(G_mt[1], h, l);
          a.dg_hi[(long long)b * D + col1] = h; if (a.dg_lo) a.dg_lo[(long long)b * D + col1] = l;
          split_bf16(G_mi[0], h, l);
          a.dg_hi[(B + b) * D + col0] = h; if (a.dg_lo) a.dg_lo[(B + b) * D + col0] = l;
          split_bf16(G_mi[1], h, l);
          a.dg_hi[(B + b) * D + col1] = h; if (a.dg_lo) a.dg_lo[(B + b) * D + col1] = l;
          if (warp == 0) {
            const float t0 = warp_sum(dbeta_mt), t1 = warp_sum(dbeta_mi);
            if (lane == 0) {
              a.dbeta[b] = t0;
              a.dbeta[B + b] = t1;
            }
          }
        }
      }
      // release the stage: generic-proxy accesses are ordered before the producer's next bulk copy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(stage));
      if (++stage == BS_STAGES) { stage = 0; phase ^= 1u; }
    }
  }
  // 6. per-thread column partials -> partials[cta][v][col]  (reduced by colsum_reduce in fixed order)
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    a.partials[((long long)blockIdx.x * 3 + v) * D + col0] = part[v][0];
    a.partials[((long long)blockIdx.x * 3 + v) * D + col1] = part[v][1];
  }
}


// ---------------------------------------------------------------------------------------------
// gcn_layer_bwd, warp-autonomous version (default when there is at least one mention per warp)
// ---------------------------------------------------------------------------------------------
// One warp owns one mention at a time and walks its 2C candidate rows (et_0, ei_0, et_1, ...), prefetching the next
// row pair (vertex row + its dz row) into registers.  The mention-side vectors and the per-mention sums over
// candidates (messages to the mention vertices, g gradients) live in a private shared-memory slice of the warp; the
// kernel-long column sums (bias / LayerNorm parameter gradients) live in registers.  No CTA barrier in the row loop;
// all sums run in candidate order and the per-warp partials are combined once, in a fixed order, at the end.
template <int D, int NW, bool FULL, bool LN, bool SLICED>
__global__ void __launch_bounds__(NW * 32, 1) gcn_layer_bwd_warp_kernel(const LayerBwdArgs a, int partial_rows) {
  constexpr int NV = RowT<D>::NV, NE = NV * 4;
  constexpr int NVEC = FULL ? 6 : 3;             // xm_t, xm_i, dz_mt (, dz_mi, g_mt, g_mi)
  constexpr int NACC = FULL ? 4 : 2;             // A_mt, A_mi (, G_mt, G_mi)
  constexpr int SLICE = (NVEC + NACC) * D;
  extern __shared__ __align__(16) float sm[];
  float* s_gamma = sm;
  float* s_beta = s_gamma + D;
  float* s_slices = s_beta + D;                  // [NW][SLICE]; reused for the end-of-kernel partials ([NW][3][D])
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long B = a.B, BC = (long long)a.B * a.C;
  constexpr bool ln = LN;                       // LayerNorm + GELU in front of the candidate rows (layers > 0)
  constexpr bool LATE_DZ = LN && FULL;           // middle layers (L >= 3): too many live rows to prefetch dz as well
  const bool dyn = FULL && a.g != nullptr;
  const bool want_de = a.dedges_in != nullptr;
  if (ln) {
    for (int i = tid; i < D; i += NW * 32) {
      s_gamma[i] = a.ln_gamma[i];
      s_beta[i] = a.ln_beta[i];
    }
  }
  __syncthreads();
  float* v_xmt = s_slices + warp * SLICE;
  float* v_xmi = v_xmt + D;
  float* v_dzmt = v_xmi + D;
  float* v_dzmi = v_dzmt + D;                    // FULL only (pointer arithmetic below stays inside the slice)
  float* v_gmt = v_dzmi + D;
  float* v_gmi = v_gmt + D;
  float* A_mt = v_xmt + NVEC * D;
  float* A_mi = A_mt + D;
  float* G_mt = A_mi + D;                        // FULL only
  float* G_mi = G_mt + D;
  const float invC = 1.0f / (float)a.C, invD = 1.0f / (float)D;
  const float* dz_mt = a.dz;
  const float* dz_mi = FULL ? a.dz + B * D : nullptr;
  const float* dz_et = a.dz + (FULL ? 2 * B : B) * D;
  const float* dz_ei = FULL ? a.dz + (2 * B + BC) * D : nullptr;
  float p0[NE], p1[NE], p2[NE];                  // ln: dgamma, dbeta, db_h ; else: db_et, db_ei, (unused)
#pragma unroll
  for (int i = 0; i < NE; ++i) p0[i] = p1[i] = p2[i] = 0.f;

  const long long gwarp = (long long)blockIdx.x * NW + warp, nwarps = (long long)gridDim.x * NW;
  constexpr uint32_t ROW_BYTES = D * sizeof(float);
  // L2 prefetch of everything one candidate needs (lane 0): runs the DRAM latency two candidates ahead of the
  // register loads, which then hit L2
  auto prefetch_candidate = [&](long long r) {
    l2_prefetch(a.x_et + r * D, ROW_BYTES);
    l2_prefetch(a.x_ei + r * D, ROW_BYTES);
    l2_prefetch(dz_et + r * D, ROW_BYTES);
    if (FULL) l2_prefetch(dz_ei + r * D, ROW_BYTES);
  };
  auto prefetch_mention = [&](long long m) {
    l2_prefetch(a.xm + m * D, ROW_BYTES);
    l2_prefetch(a.xm + (B + m) * D, ROW_BYTES);
    l2_prefetch(dz_mt + m * D, ROW_BYTES);
    if (FULL) l2_prefetch(dz_mi + m * D, ROW_BYTES);
    if (dyn) {
      l2_prefetch(a.g + m * D, ROW_BYTES);
      l2_prefetch(a.g + (B + m) * D, ROW_BYTES);
    }
  };
  // work unit = (mention, candidate slice); the unsliced instantiation (WikiDiverse-sized lists) folds S to 1
  const int S = SLICED ? a.slices : 1, cps = (a.C + S - 1) / S;
  const long long units = B * S;
  auto unit_first_row = [&](long long u) {
    const long long ub = u / S;
    return ub * a.C + min((int)(u - ub * S) * cps, a.C - 1);
  };
  if (lane == 0 && gwarp < units) {
    const long long fr = unit_first_row(gwarp);
    prefetch_candidate(fr);
    if (fr + 1 < BC) prefetch_candidate(fr + 1);
  }

  for (long long u = gwarp; u < units; u += nwarps) {
    const long long b = u / S;
    const int c0 = (int)(u - b * S) * cps, nC = min(a.C, c0 + cps) - c0;     // candidates c0 .. c0 + nC - 1
    if (SLICED && nC <= 0) {                     // empty trailing slice
      float* part = a.slice_part + u * 4 * D;
      for (int i = lane; i < 4 * D; i += 32) part[i] = 0.f;
      if (lane < 2) a.slice_dbeta[u * 2 + lane] = 0.f;
      continue;
    }
    const long long r0 = b * a.C + c0;
    RowT<D> ax, ad, bx, bd;                      // register sets of the et row pair and the ei row pair
    row_load<D>(ax, a.x_et + r0 * D, lane);
    if (!LATE_DZ) row_load<D>(ad, dz_et + r0 * D, lane);
    {
      RowT<D> t;
      row_load<D>(t, a.xm + b * D, lane); row_store<D>(t, v_xmt, lane);
      row_load<D>(t, a.xm + (B + b) * D, lane); row_store<D>(t, v_xmi, lane);
      row_load<D>(t, dz_mt + b * D, lane); row_store<D>(t, v_dzmt, lane);
      if (FULL) { row_load<D>(t, dz_mi + b * D, lane); row_store<D>(t, v_dzmi, lane); }
      if (dyn) {
        row_load<D>(t, a.g + b * D, lane); row_store<D>(t, v_gmt, lane);
        row_load<D>(t, a.g + (B + b) * D, lane); row_store<D>(t, v_gmi, lane);
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(A_mt + (j * 32 + lane) * 4) = z;
        *reinterpret_cast<float4*>(A_mi + (j * 32 + lane) * 4) = z;
        if (dyn) {
          *reinterpret_cast<float4*>(G_mt + (j * 32 + lane) * 4) = z;
          *reinterpret_cast<float4*>(G_mi + (j * 32 + lane) * 4) = z;
        }
      }
    }
    __syncwarp();
    float dbeta_mt = 0.f, dbeta_mi = 0.f;
    float e[4], ds[4] = {0.f, 0.f, 0.f, 0.f};
    float ne[4], no[4] = {0.f, 0.f, 0.f, 0.f}, ndo[4] = {0.f, 0.f, 0.f, 0.f};   // scalars of the NEXT candidate
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ne[k] = a.edges_in[k * BC + r0];
      if (dyn) {
        no[k] = a.edges_out[k * BC + r0];
        ndo[k] = a.dedges_out[k * BC + r0];
      }
    }

    // one vertex row (kind 0: entity text, 1: entity image) of candidate row r, in place: x -> unused, d -> dL/dx
    // LATE_DZ: the dz row is loaded here, from L2 (bulk-prefetched two candidates ahead); its latency hides behind
    // the LayerNorm/GELU recompute and the registers of a second prefetched row are saved.
    auto process_row = [&](auto kind_c, long long r, RowT<D>& x, RowT<D>& d, const float* dzrow, float e_mt, float e_mi,
                           float ds_mt, float ds_mi) {
      constexpr int kind = decltype(kind_c)::value;
      constexpr bool has_dz = FULL || kind == 0;
      if (!has_dz) {
#pragma unroll
        for (int i = 0; i < NE; ++i) d.v[i] = 0.f;
      } else if (LATE_DZ) {
        row_load<D>(d, dzrow, lane);
      }
      RowT<D> xhat, dact;
      float rstd = 0.f;
      if (ln) {
        xhat = x;
        rstd = row_ln_gelu_recompute<D>(xhat, x, dact, s_gamma, s_beta, lane);     // x := activated row
      }
      const float ec_mt = e_mt * invC, ec_mi = e_mi * invC, dd_mt = ds_mt * invD, dd_mi = ds_mi * invD;
      float pd_mt = 0.f, pd_mi = 0.f, q_mt = 0.f, q_mi = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int off = (j * 32 + lane) * 4;
        const float4 zmt4 = *reinterpret_cast<const float4*>(v_dzmt + off);
        const float zmt[4] = {zmt4.x, zmt4.y, zmt4.z, zmt4.w};
        float zmi[4] = {0.f, 0.f, 0.f, 0.f}, gmt[4] = {0.f, 0.f, 0.f, 0.f}, gmi[4] = {0.f, 0.f, 0.f, 0.f};
        if (FULL) {
          const float4 t = *reinterpret_cast<const float4*>(v_dzmi + off);
          zmi[0] = t.x; zmi[1] = t.y; zmi[2] = t.z; zmi[3] = t.w;
        }
        if (dyn) {
          const float4 t = *reinterpret_cast<const float4*>(v_gmt + off);
          gmt[0] = t.x; gmt[1] = t.y; gmt[2] = t.z; gmt[3] = t.w;
          const float4 u = *reinterpret_cast<const float4*>(v_gmi + off);
          gmi[0] = u.x; gmi[1] = u.y; gmi[2] = u.z; gmi[3] = u.w;
        }
        if (want_de) {
          const float4 xt4 = *reinterpret_cast<const float4*>(v_xmt + off);
          const float4 xi4 = *reinterpret_cast<const float4*>(v_xmi + off);
          const float xt[4] = {xt4.x, xt4.y, xt4.z, xt4.w}, xi[4] = {xi4.x, xi4.y, xi4.z, xi4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * j + k;
            pd_mt = fmaf(x.v[i], zmt[k], pd_mt);
            if (FULL) pd_mi = fmaf(x.v[i], zmi[k], pd_mi);
            if (has_dz) {
              q_mt = fmaf(d.v[i], xt[k], q_mt);
              q_mi = fmaf(d.v[i], xi[k], q_mi);
            }
          }
        }
        if (has_dz) {      // messages to the mention vertices: sums of the ORIGINAL dz rows
          float4 amt = *reinterpret_cast<float4*>(A_mt + off);
          float4 ami = *reinterpret_cast<float4*>(A_mi + off);
          amt.x = fmaf(e_mt, d.v[4 * j], amt.x); amt.y = fmaf(e_mt, d.v[4 * j + 1], amt.y);
          amt.z = fmaf(e_mt, d.v[4 * j + 2], amt.z); amt.w = fmaf(e_mt, d.v[4 * j + 3], amt.w);
          ami.x = fmaf(e_mi, d.v[4 * j], ami.x); ami.y = fmaf(e_mi, d.v[4 * j + 1], ami.y);
          ami.z = fmaf(e_mi, d.v[4 * j + 2], ami.z); ami.w = fmaf(e_mi, d.v[4 * j + 3], ami.w);
          *reinterpret_cast<float4*>(A_mt + off) = amt;
          *reinterpret_cast<float4*>(A_mi + off) = ami;
        }
        if (dyn) {         // gradient of g = fu W_v: sums of the activated vertex rows
          float4 gt = *reinterpret_cast<float4*>(G_mt + off);
          float4 gi = *reinterpret_cast<float4*>(G_mi + off);
          gt.x = fmaf(dd_mt, x.v[4 * j], gt.x); gt.y = fmaf(dd_mt, x.v[4 * j + 1], gt.y);
          gt.z = fmaf(dd_mt, x.v[4 * j + 2], gt.z); gt.w = fmaf(dd_mt, x.v[4 * j + 3], gt.w);
          gi.x = fmaf(dd_mi, x.v[4 * j], gi.x); gi.y = fmaf(dd_mi, x.v[4 * j + 1], gi.y);
          gi.z = fmaf(dd_mi, x.v[4 * j + 2], gi.z); gi.w = fmaf(dd_mi, x.v[4 * j + 3], gi.w);
          *reinterpret_cast<float4*>(G_mt + off) = gt;
          *reinterpret_cast<float4*>(G_mi + off) = gi;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 4 * j + k;
          d.v[i] = fmaf(ec_mt, zmt[k], d.v[i]);
          if (FULL) d.v[i] = fmaf(ec_mi, zmi[k], d.v[i]);
          if (dyn) d.v[i] += dd_mt * gmt[k] + dd_mi * gmi[k];
        }
      }
      if (want_de) {
        pd_mt = warp_sum(pd_mt); q_mt = warp_sum(q_mt); q_mi = warp_sum(q_mi);
        if (FULL) pd_mi = warp_sum(pd_mi);
        // edges (mt,row) and (mi,row): tt / it for the et row, ti / ii for the ei row; the mask is applied again
        if (lane == 0) {
          constexpr int k_mt = kind ? 1 : 0, k_mi = kind ? 3 : 2;
          a.dedges_in[k_mt * BC + r] = (pd_mt * invC + q_mt + ds_mt) * a.en[k_mt];
          a.dedges_in[k_mi * BC + r] = (pd_mi * invC + q_mi + ds_mi) * a.en[k_mi];
        }
      }
      if (ln) {
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const float4 g = *reinterpret_cast<const float4*>(s_gamma + (j * 32 + lane) * 4);
          const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * j + k;
            const float dy = d.v[i] * dact.v[i];
            p0[i] = fmaf(dy, xhat.v[i], p0[i]);
            p1[i] += dy;
            const float dxh = dy * gg[k];
            d.v[i] = dxh;
            m1 += dxh;
            m2 = fmaf(dxh, xhat.v[i], m2);
          }
        }
        m1 = warp_sum(m1) * (1.0f / D);
        m2 = warp_sum(m2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          d.v[i] = rstd * (d.v[i] - m1 - xhat.v[i] * m2);
          p2[i] += d.v[i];
        }
      } else if (kind == 0) {
#pragma unroll
        for (int i = 0; i < NE; ++i) p0[i] += d.v[i];          // db_et (bias of the entity-text projection)
      } else {
#pragma unroll
        for (int i = 0; i < NE; ++i) p1[i] += d.v[i];          // db_ei
      }
      const long long orow = (kind ? 2 * B + BC : 2 * B) + r;  // row in the [mt; mi; et; ei] layout
      row_store_planes<D>(d, a.dcand_hi + orow * D, a.dcand_lo ? a.dcand_lo + orow * D : nullptr, lane);
    };

    for (int c = 0; c < nC; ++c) {
      const long long r = r0 + c;
      // per-candidate scalars (enable mask model.py:122; sigmoid backward of the dynamic edge update), loaded one
      // candidate ahead so their latency is off the critical path
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e[k] = ne[k] * a.en[k];
        if (dyn) ds[k] = ndo[k] * no[k] * (1.f - no[k]);
      }
      if (dyn) {
        dbeta_mt += (ds[0] + ds[1]) * invD;
        dbeta_mi += (ds[2] + ds[3]) * invD;
      }
      if (c + 1 < nC) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ne[k] = a.edges_in[k * BC + r + 1];
          if (dyn) {
            no[k] = a.edges_out[k * BC + r + 1];
            ndo[k] = a.dedges_out[k * BC + r + 1];
          }
        }
      }
      if (lane == 0) {        // L2 prefetch two candidates ahead (crossing into the warp's next unit)
        if (c + 2 < nC) {
          prefetch_candidate(r + 2);
        } else if (u + nwarps < units) {
          const long long fr = unit_first_row(u + nwarps) + (c + 2 - nC);
          prefetch_candidate(fr < BC ? fr : BC - 1);
          if (c + 2 == nC) prefetch_mention((u + nwarps) / S);
        }
      }
      row_load<D>(bx, a.x_ei + r * D, lane);                    // ei rows arrive while the et row is processed
      if (FULL && !LATE_DZ) row_load<D>(bd, dz_ei + r * D, lane);
      process_row(std::integral_constant<int, 0>{}, r, ax, ad, dz_et + r * D, e[0], e[2], ds[0], ds[2]);
      asm volatile("" ::: "memory");                            // keep the two rows' instruction streams apart (registers)
      if (c + 1 < nC) {                                         // next candidate's et rows arrive during the ei row
        row_load<D>(ax, a.x_et + (r + 1) * D, lane);
        if (!LATE_DZ) row_load<D>(ad, dz_et + (r + 1) * D, lane);
      }
      process_row(std::integral_constant<int, 1>{}, r, bx, bd, FULL ? dz_ei + r * D : nullptr, e[1], e[3], ds[1], ds[3]);
    }
    __syncwarp();
    if (SLICED) {
      // sliced mention: the slice's sums go to layer_bwd_slice_finish (fixed slice order -> deterministic)
      float* part = a.slice_part + u * 4 * D;
      RowT<D> t;
      row_load<D>(t, A_mt, lane); row_store<D>(t, part, lane);
      row_load<D>(t, A_mi, lane); row_store<D>(t, part + D, lane);
      if (dyn) {
        row_load<D>(t, G_mt, lane); row_store<D>(t, part + 2 * D, lane);
        row_load<D>(t, G_mi, lane); row_store<D>(t, part + 3 * D, lane);
        if (lane == 0) {
          a.slice_dbeta[u * 2] = dbeta_mt;
          a.slice_dbeta[u * 2 + 1] = dbeta_mi;
        }
      }
      __syncwarp();
      continue;
    }
    // ---- mention-side results: dxm = dz_m + sum_c(...) ; dg ; dbeta
    {
      RowT<D> t, z;
      row_load<D>(t, A_mt, lane);
      row_load<D>(z, v_dzmt, lane);
#pragma unroll
      for (int i = 0; i < NE; ++i) t.v[i] += z.v[i];
      row_store<D>(t, a.dxm + b * D, lane);
      row_load<D>(t, A_mi, lane);
      if (FULL) {
        row_load<D>(z, v_dzmi, lane);
#pragma unroll
        for (int i = 0; i < NE; ++i) t.v[i] += z.v[i];
      }
      row_store<D>(t, a.dxm + (B + b) * D, lane);
      if (dyn) {
        row_load<D>(t, G_mt, lane);
        row_store_planes<D>(t, a.dg_hi + b * D, a.dg_lo ? a.dg_lo + b * D : nullptr, lane);
        row_load<D>(t, G_mi, lane);
        row_store_planes<D>(t, a.dg_hi + (B + b) * D, a.dg_lo ? a.dg_lo + (B + b) * D : nullptr, lane);
        if (lane == 0) {
          a.dbeta[b] = dbeta_mt;
          a.dbeta[B + b] = dbeta_mi;
        }
      }
    }
    __syncwarp();                                // the slice is rewritten for the next mention
  }

  // ---- fixed-order combination of the per-warp column partials; unused rows of the partial buffer are zeroed
  __syncthreads();
  float* mine = s_slices + warp * 3 * D;         // SLICE >= 5 D >= 3 D
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int off = (j * 32 + lane) * 4;
    *reinterpret_cast<float4*>(mine + off) = make_float4(p0[4 * j], p0[4 * j + 1], p0[4 * j + 2], p0[4 * j + 3]);
    *reinterpret_cast<float4*>(mine + D + off) = make_float4(p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
    *reinterpret_cast<float4*>(mine + 2 * D + off) = make_float4(p2[4 * j], p2[4 * j + 1], p2[4 * j + 2], p2[4 * j + 3]);
  }
  __syncthreads();
  for (int i = tid; i < 3 * D; i += NW * 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_slices[w * 3 * D + i];
    a.partials[(long long)blockIdx.x * 3 * D + i] = t;
    for (int extra = blockIdx.x + gridDim.x; extra < partial_rows; extra += gridDim.x)
      a.partials[(long long)extra * 3 * D + i] = 0.f;
  }
}

// sliced mentions: dxm = dz_m + sum over slices of A ; dg = sum of G ; dbeta = sum -- one warp per mention row
template <int D>
__global__ void __launch_bounds__(256) layer_bwd_slice_finish_kernel(const LayerBwdArgs a) {
  constexpr int NE = RowT<D>::NV * 4;
  const int lane = threadIdx.x & 31;
  const long long B = a.B;
  const bool dyn = a.full && a.g != nullptr;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < 2 * B; r += nwarps) {
    const long long b = r < B ? r : r - B;
    const int which = r < B ? 0 : 1;                     // mention text / mention image vertex
    RowT<D> acc, t;
    if (which == 0 || a.full) {
      row_load<D>(acc, a.dz + r * D, lane);              // dz rows of this layer: mt rows, then (full) mi rows
    } else {
#pragma unroll
      for (int i = 0; i < NE; ++i) acc.v[i] = 0.f;
    }
    for (int sl = 0; sl < a.slices; ++sl) {
      row_load<D>(t, a.slice_part + ((b * a.slices + sl) * 4 + which) * D, lane);
#pragma unroll
      for (int i = 0; i < NE; ++i) acc.v[i] += t.v[i];
    }
    row_store<D>(acc, a.dxm + r * D, lane);
    if (dyn) {
#pragma unroll
      for (int i = 0; i < NE; ++i) acc.v[i] = 0.f;
      float db = 0.f;
      for (int sl = 0; sl < a.slices; ++sl) {
        row_load<D>(t, a.slice_part + ((b * a.slices + sl) * 4 + 2 + which) * D, lane);
#pragma unroll
        for (int i = 0; i < NE; ++i) acc.v[i] += t.v[i];
        db += a.slice_dbeta[(b * a.slices + sl) * 2 + which];
      }
      row_store_planes<D>(acc, a.dg_hi + r * D, a.dg_lo ? a.dg_lo + r * D : nullptr, lane);
      if (lane == 0) a.dbeta[r] = db;
    }
  }
}

static int g_layer_bwd_variant = -1;     // -1 auto, 0 staged CTA-per-SM kernel, 1 warp-per-mention kernel (test hook)
void debug_set_layer_bwd_variant(int v) { g_layer_bwd_variant = v; }

template <int D, int NW, bool FULL, bool LN>
static int launch_layer_bwd_warp(cudaStream_t stream, const LayerBwdArgs& a) {
  const size_t smem = (size_t)(2 + NW * (FULL ? 10 : 5)) * D * sizeof(float);
  if (a.slices > 1) {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_bwd_warp_kernel<D, NW, FULL, LN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gcn_layer_bwd_warp_kernel<D, NW, FULL, LN, true><<<BS_GRID, NW * 32, smem, stream>>>(a, BS_GRID);
  } else {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_bwd_warp_kernel<D, NW, FULL, LN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gcn_layer_bwd_warp_kernel<D, NW, FULL, LN, false><<<BS_GRID, NW * 32, smem, stream>>>(a, BS_GRID);
  }
  DRIN_LAUNCH_CHECK();
  if (a.slices > 1) {
    const long long rows = 2LL * a.B;
    const int grid = (int)((rows + 7) / 8 < 148 * 4 ? (rows + 7) / 8 : 148 * 4);
    layer_bwd_slice_finish_kernel<D><<<grid, 256, 0, stream>>>(a);
    DRIN_LAUNCH_CHECK();
  }
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// gcn_layer_bwd, column-wise version for the FIRST layer.  Its candidate rows are projection outputs (no LayerNorm in
// front) and its edges are inputs (no edge gradient), so nothing in its backward is a dot product along D: a thread
// owns four columns of one mention and walks the candidates -- no shuffles, no shared memory, 16-byte coalesced
// accesses, sums over candidates in registers (candidate order -> deterministic).  Same outputs as the warp kernel.
// ---------------------------------------------------------------------------------------------
static constexpr int COL_CTAS_MAX = 148 * 3;     // rows of the partial buffer this kernel may use (= BW_CTAS rows per region)

__device__ __forceinline__ float4 col_ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 col_fma(float s, const float4& a, const float4& c) {
  return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ void col_acc(float4& c, const float4& a) { c.x += a.x; c.y += a.y; c.z += a.z; c.w += a.w; }
__device__ __forceinline__ void col_st_planes(bf16* hi, bf16* lo, long long idx, const float4& v) {
  uint32_t h0, l0, h1, l1;
  split_bf16x2(v.x, v.y, h0, l0);
  split_bf16x2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h0, h1);
  if (lo) *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l0, l1);
}

template <int D, bool FULL, bool DYN>
__global__ void __launch_bounds__(D / 4, 3) gcn_layer0_bwd_col_kernel(const LayerBwdArgs a) {
  const int col = threadIdx.x * 4;
  const long long B = a.B, C = a.C, BC = B * C;
  const float invC = 1.0f / (float)a.C, invD = 1.0f / (float)D;
  const float* dz_et = a.dz + (FULL ? 2 * B : B) * D;
  const float* dz_ei = FULL ? a.dz + (2 * B + BC) * D : nullptr;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 p_et = zero, p_ei = zero;               // db_et, db_ei (bias gradients of the entity projections)
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const float4 dzmt = col_ld(a.dz + b * D + col);
    const float4 dzmi = FULL ? col_ld(a.dz + (B + b) * D + col) : zero;
    float4 gmt = zero, gmi = zero;
    if (DYN) {
      gmt = col_ld(a.g + b * D + col);
      gmi = col_ld(a.g + (B + b) * D + col);
    }
    float4 A_mt = dzmt, A_mi = dzmi, G_mt = zero, G_mi = zero;     // dxm = dz_m + messages from the candidate rows
    float dbeta_mt = 0.f, dbeta_mi = 0.f;
    for (long long c = 0; c < C; ++c) {
      const long long r = b * C + c;
      const float4 dzet = col_ld(dz_et + r * D + col);
      const float4 dzei = FULL ? col_ld(dz_ei + r * D + col) : zero;
      float e[4], dd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e[k] = a.edges_in[k * BC + r] * a.en[k];                  // enable mask (model.py:122)
        if (DYN) {
          const float o = a.edges_out[k * BC + r];
          dd[k] = a.dedges_out[k * BC + r] * o * (1.f - o) * invD;    // sigmoid backward of the edge update, / D of the mean
        }
      }
      // z_et = et + e0 mt + e2 mi, z_ei = ei + e1 mt + e3 mi, z_m += mean_c(e . cand);  edge update: v . g_u
      float4 det = col_fma(e[0] * invC, dzmt, dzet);
      float4 dei = col_fma(e[1] * invC, dzmt, dzei);
      if (FULL) {
        det = col_fma(e[2] * invC, dzmi, det);
        dei = col_fma(e[3] * invC, dzmi, dei);
      }
      A_mt = col_fma(e[0], dzet, A_mt);
      A_mi = col_fma(e[2], dzet, A_mi);
      if (FULL) {
        A_mt = col_fma(e[1], dzei, A_mt);
        A_mi = col_fma(e[3], dzei, A_mi);
      }
      if (DYN) {
        const float4 xet = col_ld(a.x_et + r * D + col);
        const float4 xei = col_ld(a.x_ei + r * D + col);
        det = col_fma(dd[0], gmt, col_fma(dd[2], gmi, det));
        dei = col_fma(dd[1], gmt, col_fma(dd[3], gmi, dei));
        G_mt = col_fma(dd[0], xet, col_fma(dd[1], xei, G_mt));
        G_mi = col_fma(dd[2], xet, col_fma(dd[3], xei, G_mi));
        dbeta_mt += dd[0] + dd[1];
        dbeta_mi += dd[2] + dd[3];
      }
      col_acc(p_et, det);
      col_acc(p_ei, dei);
      col_st_planes(a.dcand_hi, a.dcand_lo, (2 * B + r) * D + col, det);
      col_st_planes(a.dcand_hi, a.dcand_lo, (2 * B + BC + r) * D + col, dei);
    }
    *reinterpret_cast<float4*>(a.dxm + b * D + col) = A_mt;
    *reinterpret_cast<float4*>(a.dxm + (B + b) * D + col) = A_mi;
    if (DYN) {
      col_st_planes(a.dg_hi, a.dg_lo, b * D + col, G_mt);
      col_st_planes(a.dg_hi, a.dg_lo, (B + b) * D + col, G_mi);
      if (threadIdx.x == 0) {
        a.dbeta[b] = dbeta_mt;
        a.dbeta[B + b] = dbeta_mi;
      }
    }
  }
  float* part = a.partials + (long long)blockIdx.x * 3 * D + col;
  *reinterpret_cast<float4*>(part) = p_et;
  *reinterpret_cast<float4*>(part + D) = p_ei;
  *reinterpret_cast<float4*>(part + 2 * D) = zero;
}

template <int D, bool FULL, bool DYN>
static int launch_layer0_bwd_col(cudaStream_t stream, const LayerBwdArgs& a) {
  int resident = 0;
  DRIN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, gcn_layer0_bwd_col_kernel<D, FULL, DYN>, D / 4, 0));
  if (resident < 1) resident = 1;
  int grid = 148 * resident;                     // every CTA of the grid-stride loop resident from the start
  if (grid > COL_CTAS_MAX) grid = COL_CTAS_MAX;
  if (grid > a.B) grid = a.B;
  gcn_layer0_bwd_col_kernel<D, FULL, DYN><<<grid, D / 4, 0, stream>>>(a);
  DRIN_LAUNCH_CHECK();
  if (a.partial_rows) *a.partial_rows = grid;
  return DRIN_OK;
}

int gcn_layer_bwd(cudaStream_t stream, const LayerBwdArgs& a_in) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  if (a_in.D != 768) return fail(DRIN_ERR_ARG, "gcn_layer_bwd: gcn_embed_dim %d not built (768 only)", a_in.D);
  constexpr int D = 768;
  LayerBwdArgs a = a_in;
  if (a.slices < 1 || !a.slice_part || !a.slice_dbeta) a.slices = 1;
  if (a.partial_rows) *a.partial_rows = BS_GRID;
  const bool warp_kernel = g_layer_bwd_variant < 0 ? (long long)a.B * a.slices >= BS_GRID * 7 : g_layer_bwd_variant >= 1;
  // first layer (no LayerNorm in front, edges are inputs): nothing is a dot product along D -> column-wise kernel.
  // Variant 2 forces it (tests), variant 0 / 1 force the other two families.
  const bool first_layer = !a.ln_gamma && !a.dedges_in && a.partial_rows;
  if (first_layer && (g_layer_bwd_variant == 2 || (g_layer_bwd_variant < 0 && warp_kernel && a.slices == 1))) {
    const bool dyn = a.full && a.g != nullptr;
    if (a.full) return dyn ? launch_layer0_bwd_col<D, true, true>(stream, a) : launch_layer0_bwd_col<D, true, false>(stream, a);
    return launch_layer0_bwd_col<D, false, false>(stream, a);
  }
  if (warp_kernel) {
    // 7 warps x 30 KB (full layers) or 8 warps x 15 KB (last layer) of private shared memory per SM
    const bool ln = a.ln_gamma != nullptr;
    if (a.full) return ln ? launch_layer_bwd_warp<D, 7, true, true>(stream, a) : launch_layer_bwd_warp<D, 7, true, false>(stream, a);
    return ln ? launch_layer_bwd_warp<D, 8, false, true>(stream, a) : launch_layer_bwd_warp<D, 8, false, false>(stream, a);
  }

  // persistent: every CTA of the fixed partial-sum grid writes its partials
  const size_t smem = (size_t)(BS_STAGES * (4 * BS_CH + 6) * D + 2 * D + 2 * BS_CH * 8 + 2 * BS_CH * 8) * sizeof(float);
  if (a.full) {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_bwd_stream_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gcn_layer_bwd_stream_kernel<D, true><<<BS_GRID, BS_THREADS, smem, stream>>>(a);
  } else {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_bwd_stream_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gcn_layer_bwd_stream_kernel<D, false><<<BS_GRID, BS_THREADS, smem, stream>>>(a);
  }
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// mention_bwd_finish: rows of the 2B mention vertices
// ---------------------------------------------------------------------------------------------
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) mention_bwd_finish_kernel(const MentionBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_gamma = sm;
  float* s_beta = s_gamma + D;
  float* s_part = s_beta + D;              // [3][NW][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool ln = a.ln_gamma != nullptr;
  if (ln) {
    for (int i = tid; i < D; i += NW * 32) {
      s_gamma[i] = a.ln_gamma[i];
      s_beta[i] = a.ln_beta[i];
    }
  }
  for (int i = tid; i < 3 * NW * D; i += NW * 32) s_part[i] = 0.f;
  __syncthreads();
  float* p0 = s_part + (0 * NW + warp) * D;
  float* p1 = s_part + (1 * NW + warp) * D;
  float* p2 = s_part + (2 * NW + warp) * D;
  const long long rows = 2LL * a.B;
  for (long long r = (long long)blockIdx.x * NW + warp; r < rows; r += (long long)gridDim.x * NW) {
    RowT<D> d;
    row_load<D>(d, a.dxm + r * D, lane);
    if (a.dxu) {
      RowT<D> u;
      row_load<D>(u, a.dxu + r * D, lane);
#pragma unroll
      for (int i = 0; i < RowT<D>::NV * 4; ++i) d.v[i] += u.v[i];
    }
    if (ln) {
      RowT<D> h;
      row_load<D>(h, a.h_prev + r * D, lane);
      row_ln_gelu_bwd<D>(h, d, s_gamma, s_beta, p0, p1, lane);
      row_accum_smem<D>(d, p2, lane);
    } else {
      row_accum_smem<D>(d, r < a.B ? p0 : p1, lane);      // db_mt / db_mi
    }
    row_store_planes<D>(d, a.out_hi + r * D, a.out_lo ? a.out_lo + r * D : nullptr, lane);
  }
  __syncthreads();
  flush_partials<D, NW>(s_part, 3, a.partials, tid);
}

int mention_bwd_finish(cudaStream_t stream, const MentionBwdArgs& a) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  if (a.D != 768) return fail(DRIN_ERR_ARG, "mention_bwd_finish: gcn_embed_dim %d not built (768 only)", a.D);
  constexpr int D = 768;
  const size_t smem = (size_t)(2 + 3 * BW_NW) * D * sizeof(float);
  mention_bwd_finish_kernel<D, BW_NW><<<BW_CTAS, BW_NW * 32, smem, stream>>>(a);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// dfu_finish: dfu = dfu_raw + dbeta * b_v (in place + planes); partials: db_u = sum dfu, db_v = sum dbeta * fu
// ---------------------------------------------------------------------------------------------
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) dfu_finish_kernel(float* __restrict__ dfu, const float* __restrict__ dbeta,
                                                              const float* __restrict__ b_v,
                                                              const float* __restrict__ fu, long long rows,
                                                              bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                                                              float* __restrict__ partials) {
  extern __shared__ __align__(16) float sm[];
  float* s_bv = sm;
  float* s_part = s_bv + D;                // [2][NW][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += NW * 32) s_bv[i] = b_v[i];
  for (int i = tid; i < 2 * NW * D; i += NW * 32) s_part[i] = 0.f;
  __syncthreads();
  float* p0 = s_part + (0 * NW + warp) * D;
  float* p1 = s_part + (1 * NW + warp) * D;
  for (long long r = (long long)blockIdx.x * NW + warp; r < rows; r += (long long)gridDim.x * NW) {
    RowT<D> d, f;
    row_load<D>(d, dfu + r * D, lane);
    row_load<D>(f, fu + r * D, lane);
    const float db = dbeta[r];
#pragma unroll
    for (int j = 0; j < RowT<D>::NV; ++j) {
      const float4 bv = *reinterpret_cast<const float4*>(s_bv + (j * 32 + lane) * 4);
      d.v[4 * j] += db * bv.x; d.v[4 * j + 1] += db * bv.y; d.v[4 * j + 2] += db * bv.z; d.v[4 * j + 3] += db * bv.w;
      f.v[4 * j] *= db; f.v[4 * j + 1] *= db; f.v[4 * j + 2] *= db; f.v[4 * j + 3] *= db;
    }
    row_accum_smem<D>(d, p0, lane);
    row_accum_smem<D>(f, p1, lane);
    row_store<D>(d, dfu + r * D, lane);
    row_store_planes<D>(d, out_hi + r * D, out_lo ? out_lo + r * D : nullptr, lane);
  }
  __syncthreads();
  flush_partials<D, NW>(s_part, 2, partials, tid);
}

int dfu_finish(cudaStream_t stream, int D, float* dfu, const float* dbeta, const float* b_v, const float* fu,
               long long rows, bf16* out_hi, bf16* out_lo, float* partials) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  if (D != 768) return fail(DRIN_ERR_ARG, "dfu_finish: gcn_embed_dim %d not built (768 only)", D);
  const size_t smem = (size_t)(1 + 2 * BW_NW) * 768 * sizeof(float);
  dfu_finish_kernel<768, BW_NW><<<BW_CTAS, BW_NW * 32, smem, stream>>>(dfu, dbeta, b_v, fu, rows, out_hi, out_lo,
                                                                       partials);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// colsum_reduce: out_v[col] = sum over sources and CTAs of partials[(cta * nvec + v) * D + col]
// ---------------------------------------------------------------------------------------------
// block = 32 columns x 8 CTA-slices; slice s sums partial rows s, s+8, ... then a fixed-order smem reduce
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ src0, int ctas0,
                                                            const float* __restrict__ src1, int ctas1, int nvec, int D,
                                                            float* out0, float* out1, float* out2) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;                  // flat (v, col)
  float t = 0.f;
  if (i < nvec * D) {
    const int v = i / D, col = i - v * D;
    for (int c = slice; c < ctas0; c += 8) t += src0[((long long)c * nvec + v) * D + col];
    if (src1)
      for (int c = slice; c < ctas1; c += 8) t += src1[((long long)c * nvec + v) * D + col];
  }
  red[slice][lane] = t;
  __syncthreads();
  if (slice == 0 && i < nvec * D) {
    float r = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 8; ++s2) r += red[s2][lane];
    const int v = i / D, col = i - v * D;
    float* out = v == 0 ? out0 : (v == 1 ? out1 : out2);
    if (out) out[col] = r;
  }
}

int colsum_reduce(cudaStream_t stream, const float* src0, int ctas0, const float* src1, int ctas1, int nvec, int D,
                  float* out0, float* out1, float* out2) {
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  const int n = nvec * D;
  colsum_reduce_kernel<<<(n + 31) / 32, 256, 0, stream>>>(src0, ctas0, src1, ctas1, nvec, D, out0, out1, out2);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// every deferred column-sum reduction of a backward pass in one launch: blockIdx.y selects the job
__global__ void __launch_bounds__(256) colsum_reduce_multi_kernel(const ColsumJobs jobs, int D) {
  __shared__ float red[8][33];
  const ColsumJob j = jobs.job[blockIdx.y];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;                  // flat (v, col)
  float t = 0.f;
  if (i < j.nvec * D) {
    const int v = i / D, col = i - v * D;
    for (int c = slice; c < j.ctas0; c += 8) t += j.src0[((long long)c * j.nvec + v) * D + col];
    if (j.src1)
      for (int c = slice; c < j.ctas1; c += 8) t += j.src1[((long long)c * j.nvec + v) * D + col];
  }
  red[slice][lane] = t;
  __syncthreads();
  if (slice == 0 && i < j.nvec * D) {
    float r = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 8; ++s2) r += red[s2][lane];
    const int v = i / D, col = i - v * D;
    float* out = v == 0 ? j.out0 : (v == 1 ? j.out1 : j.out2);
    if (out) out[col] = r;
  }
}

int colsum_reduce_multi(cudaStream_t stream, const ColsumJobs& jobs, int D) {
  if (jobs.count <= 0) return DRIN_OK;
  prof::Scope prof_scope(stream, prof::GCN_BWD);
  colsum_reduce_multi_kernel<<<dim3((3 * D + 31) / 32, jobs.count), 256, 0, stream>>>(jobs, D);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

int backward_ctas() { return BW_CTAS; }
int layer_bwd_ctas() { return BS_GRID; }

}  // namespace drin
