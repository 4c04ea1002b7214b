// Warp-per-row register tiles for the 768-wide vertex rows (shared by the forward and backward GCN kernels).
// Lane l of a warp holds elements (j*32 + l)*4 .. +3 for j < D/128: every global access is a coalesced
// 512-B warp transaction and every shared-memory access a conflict-free 128-bit one.
#pragma once
#include "common.cuh"

namespace drin {

template <int D>
struct RowT {
  static constexpr int NV = D / 128;      // float4 per lane
  float v[NV * 4];
};

template <int D>
__device__ __forceinline__ void row_load(RowT<D>& r, const float* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(p + (j * 32 + lane) * 4);
    r.v[4 * j] = t.x; r.v[4 * j + 1] = t.y; r.v[4 * j + 2] = t.z; r.v[4 * j + 3] = t.w;
  }
}
template <int D>
__device__ __forceinline__ void row_store(const RowT<D>& r, float* __restrict__ p, int lane) {
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j)
    *reinterpret_cast<float4*>(p + (j * 32 + lane) * 4) =
        make_float4(r.v[4 * j], r.v[4 * j + 1], r.v[4 * j + 2], r.v[4 * j + 3]);
}
// (hi, lo) planes of two floats with packed converts: hi2 = cvt.rn.bf16x2(b, a); lo2 = cvt.rn.bf16x2(b - hi_b, a - hi_a)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi2 << 16);
  const float rb = b - __uint_as_float(hi2 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(rb), "f"(ra));
}
template <int D>
__device__ __forceinline__ void row_store_planes(const RowT<D>& r, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                                 int lane) {
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    uint32_t h0, l0, h1, l1;
    split_bf16x2(r.v[4 * j], r.v[4 * j + 1], h0, l0);
    split_bf16x2(r.v[4 * j + 2], r.v[4 * j + 3], h1, l1);
    *reinterpret_cast<uint2*>(hi + (j * 32 + lane) * 4) = make_uint2(h0, h1);
    if (lo) *reinterpret_cast<uint2*>(lo + (j * 32 + lane) * 4) = make_uint2(l0, l1);
  }
}
template <int D>
__device__ __forceinline__ float row_dot(const RowT<D>& a, const float* __restrict__ s, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(s + (j * 32 + lane) * 4);
    acc += a.v[4 * j] * t.x + a.v[4 * j + 1] * t.y + a.v[4 * j + 2] * t.z + a.v[4 * j + 3] * t.w;
  }
  return acc;
}

// y = gelu(LayerNorm(h)) in place (eps 1e-5, biased variance, like nn.LayerNorm); gamma/beta in smem.
// Packed fp32 (two elements per instruction, see gelu_both2); the row sums run as two interleaved partial sums
// (even / odd elements) that are added before the warp reduction.
template <int D>
__device__ __forceinline__ void row_ln_gelu(RowT<D>& r, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            int lane) {
  float2 s2 = f2(0.f);
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; i += 2) s2 = __fadd2_rn(s2, f2(r.v[i], r.v[i + 1]));
  const float mean = warp_sum(s2.x + s2.y) * (1.0f / D);
  float2 q2 = f2(0.f);
  const float2 nmean = f2(-mean);
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; i += 2) {
    const float2 d = __fadd2_rn(f2(r.v[i], r.v[i + 1]), nmean);
    q2 = __ffma2_rn(d, d, q2);
  }
  const float rstd = rsqrtf(warp_sum(q2.x + q2.y) * (1.0f / D) + 1e-5f);
  const float2 rs = f2(rstd), sh = f2(-mean * rstd);       // xhat = h * rstd + shift: one FMA per element
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    const float4 g = *reinterpret_cast<const float4*>(gamma + (j * 32 + lane) * 4);
    const float4 b = *reinterpret_cast<const float4*>(beta + (j * 32 + lane) * 4);
    const float2 y0 = gelu_f2(__ffma2_rn(__ffma2_rn(f2(r.v[4 * j], r.v[4 * j + 1]), rs, sh), f2(g.x, g.y), f2(b.x, b.y)));
    const float2 y1 = gelu_f2(__ffma2_rn(__ffma2_rn(f2(r.v[4 * j + 2], r.v[4 * j + 3]), rs, sh), f2(g.z, g.w), f2(b.z, b.w)));
    r.v[4 * j] = y0.x; r.v[4 * j + 1] = y0.y; r.v[4 * j + 2] = y1.x; r.v[4 * j + 3] = y1.y;
  }
}


// acc[slice] += r  (per-warp shared-memory accumulator slice, no conflicts, no atomics)
template <int D>
__device__ __forceinline__ void row_accum_smem(const RowT<D>& r, float* __restrict__ s, int lane) {
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    float4* p = reinterpret_cast<float4*>(s + (j * 32 + lane) * 4);
    float4 t = *p;
    t.x += r.v[4 * j]; t.y += r.v[4 * j + 1]; t.z += r.v[4 * j + 2]; t.w += r.v[4 * j + 3];
    *p = t;
  }
}

// Backward of a = gelu(LayerNorm(h)): on entry `d` holds dL/da, on exit dL/dh.
// Accumulates dL/dgamma, dL/dbeta into the per-warp slices pg, pb.
template <int D>
__device__ __forceinline__ void row_ln_gelu_bwd(const RowT<D>& h, RowT<D>& d, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float* __restrict__ pg,
                                                float* __restrict__ pb, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; ++i) s += h.v[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; ++i) {
    const float t = h.v[i] - mean;
    q += t * t;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    const int off = (j * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + off);
    const float4 b = *reinterpret_cast<const float4*>(beta + off);
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
    float4 ag = *reinterpret_cast<float4*>(pg + off);
    float4 ab = *reinterpret_cast<float4*>(pb + off);
    float* agp = reinterpret_cast<float*>(&ag);
    float* abp = reinterpret_cast<float*>(&ab);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (h.v[4 * j + k] - mean) * rstd;
      const float dy = d.v[4 * j + k] * gelu_grad_f(xh * gg[k] + bb[k]);
      agp[k] += dy * xh;
      abp[k] += dy;
      const float dxh = dy * gg[k];
      d.v[4 * j + k] = dxh;
      m1 += dxh;
      m2 += dxh * xh;
    }
    *reinterpret_cast<float4*>(pg + off) = ag;
    *reinterpret_cast<float4*>(pb + off) = ab;
  }
  m1 = warp_sum(m1) * (1.0f / D);
  m2 = warp_sum(m2) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; ++i) {
    const float xh = (h.v[i] - mean) * rstd;
    d.v[i] = rstd * (d.v[i] - m1 - xh * m2);
  }
}

// Forward recompute for the backward kernels, done ONCE per row: h -> xhat (in place), a = gelu(y),
// da = gelu'(y) with y = gamma * xhat + beta; returns rstd.  Packed fp32 like row_ln_gelu (same per-element arithmetic).
template <int D>
__device__ __forceinline__ float row_ln_gelu_recompute(RowT<D>& h, RowT<D>& act, RowT<D>& dact,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int lane) {
  float2 s2 = f2(0.f);
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; i += 2) s2 = __fadd2_rn(s2, f2(h.v[i], h.v[i + 1]));
  const float mean = warp_sum(s2.x + s2.y) * (1.0f / D);
  float2 q2 = f2(0.f);
  const float2 nmean = f2(-mean);
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; i += 2) {
    const float2 t = __fadd2_rn(f2(h.v[i], h.v[i + 1]), nmean);
    q2 = __ffma2_rn(t, t, q2);
  }
  const float rstd = rsqrtf(warp_sum(q2.x + q2.y) * (1.0f / D) + 1e-5f);
  const float2 rs = f2(rstd), sh = f2(-mean * rstd);
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    const int off = (j * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + off);
    const float4 b = *reinterpret_cast<const float4*>(beta + off);
    const float2 gg[2] = {f2(g.x, g.y), f2(g.z, g.w)}, bb[2] = {f2(b.x, b.y), f2(b.z, b.w)};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = 4 * j + 2 * k;
      const float2 xh = __ffma2_rn(f2(h.v[i], h.v[i + 1]), rs, sh);
      h.v[i] = xh.x; h.v[i + 1] = xh.y;
      float2 a, da;
      gelu_both2(__ffma2_rn(xh, gg[k], bb[k]), a, da);
      act.v[i] = a.x; act.v[i + 1] = a.y;
      dact.v[i] = da.x; dact.v[i + 1] = da.y;
    }
  }
  return rstd;
}

// LayerNorm+GELU backward from the recomputed pieces: on entry d = dL/da, on exit dL/dh.
template <int D>
__device__ __forceinline__ void row_ln_gelu_bwd_from(const RowT<D>& xhat, const RowT<D>& dact, float rstd, RowT<D>& d,
                                                     const float* __restrict__ gamma, float* __restrict__ pg,
                                                     float* __restrict__ pb, int lane) {
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int j = 0; j < RowT<D>::NV; ++j) {
    const int off = (j * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + off);
    const float gg[4] = {g.x, g.y, g.z, g.w};
    float4 ag = *reinterpret_cast<float4*>(pg + off);
    float4 ab = *reinterpret_cast<float4*>(pb + off);
    float* agp = reinterpret_cast<float*>(&ag);
    float* abp = reinterpret_cast<float*>(&ab);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = xhat.v[4 * j + k];
      const float dy = d.v[4 * j + k] * dact.v[4 * j + k];
      agp[k] += dy * xh;
      abp[k] += dy;
      const float dxh = dy * gg[k];
      d.v[4 * j + k] = dxh;
      m1 += dxh;
      m2 += dxh * xh;
    }
    *reinterpret_cast<float4*>(pg + off) = ag;
    *reinterpret_cast<float4*>(pb + off) = ab;
  }
  m1 = warp_sum(m1) * (1.0f / D);
  m2 = warp_sum(m2) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < RowT<D>::NV * 4; ++i) d.v[i] = rstd * (d.v[i] - m1 - xhat.v[i] * m2);
}

}  // namespace drin
