// Shared device/host helpers for the drin_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace drin {

// ---------------------------------------------------------------------------------------------
// host: error reporting (C-ABI functions return an int status; text via drin_last_error())
// ---------------------------------------------------------------------------------------------
enum Status : int {
  DRIN_OK = 0,
  DRIN_ERR_ARG = 1,       // bad argument / unsupported shape
  DRIN_ERR_CUDA = 2,      // CUDA runtime or driver error
  DRIN_ERR_WORKSPACE = 3, // workspace too small
};

int fail(int code, const char* fmt, ...);   // records the message, returns `code`

#define DRIN_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return ::drin::fail(::drin::DRIN_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, \
                          cudaGetErrorString(e__));                                        \
  } while (0)

void count_launch();          // prof.cu: every kernel launch of this library is counted
long long launch_count();
#define DRIN_LAUNCH_CHECK()          \
  do {                               \
    ::drin::count_launch();          \
    DRIN_CUDA(cudaGetLastError());   \
  } while (0)

#define DRIN_TRY(expr)                 \
  do {                                 \
    int s__ = (expr);                  \
    if (s__ != ::drin::DRIN_OK) return s__; \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// device: warp reductions and vector access
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load (read-once inputs: keep them out of L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// split-bf16 representation of an fp32 value: x ~= hi + lo with |x - hi - lo| <= 2^-18 |x|
__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(bf16 a, bf16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// erf-based GELU (F.gelu default) and its derivative from ONE exp + ONE reciprocal:
// erf(u) = 1 - (a1 t + .. + a5 t^5) exp(-u^2), t = 1 / (1 + p u)  (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7),
// and exp(-u^2) = exp(-x^2 / 2) is also the Gaussian density needed by the derivative.
// MUFU approximations with flush-to-zero: no denormal pre/post-scaling code around the special-function unit
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float exp2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
  const float u = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_ftz(fmaf(0.3275911f, u, 1.0f));                  // argument >= 1
  const float ex = exp2_ftz((x * x) * -0.72134752044448170f);          // exp(-x^2 / 2) = exp(-u^2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, ex, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  g = x * cdf;
  dg = fmaf(x * 0.3989422804014327f, ex, cdf);
}
// Packed-fp32 (FFMA2 / FMUL2, sm_100 `fma.rn.f32x2`) form of gelu_both for two elements at once.  The FMA pipe is not
// faster per flop (scripts/ubench/fp32_pipe.cu: 70.8 vs 73.2 TFLOP/s) but a packed instruction takes ONE issue slot for
// two results, and the row kernels are issue-bound.  Every lane-element goes through exactly the same
// round-to-nearest operations as the scalar version: bit-identical results.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ void gelu_both2(float2 x, float2& g, float2& dg) {
  const float2 u = __fmul2_rn(f2(fabsf(x.x), fabsf(x.y)), f2(0.70710678118654752f));
  const float2 den = __ffma2_rn(f2(0.3275911f), u, f2(1.0f));
  const float2 t = f2(rcp_ftz(den.x), rcp_ftz(den.y));
  const float2 e = __fmul2_rn(__fmul2_rn(x, x), f2(-0.72134752044448170f));
  const float2 ex = f2(exp2_ftz(e.x), exp2_ftz(e.y));
  float2 poly = __ffma2_rn(t, f2(1.061405429f), f2(-1.453152027f));
  poly = __ffma2_rn(t, poly, f2(1.421413741f));
  poly = __ffma2_rn(t, poly, f2(-0.284496736f));
  poly = __ffma2_rn(t, poly, f2(0.254829592f));
  const float2 npt = __fmul2_rn(f2(-poly.x, -poly.y), t);               // (-poly) * t, as the scalar code
  const float2 erf_abs = __ffma2_rn(npt, ex, f2(1.0f));
  const float2 cdf = __ffma2_rn(f2(0.5f), f2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y)), f2(0.5f));
  g = __fmul2_rn(x, cdf);
  dg = __ffma2_rn(__fmul2_rn(x, f2(0.3989422804014327f)), ex, cdf);
}
__device__ __forceinline__ float2 gelu_f2(float2 x) {
  float2 g, dg;
  gelu_both2(x, g, dg);
  return g;
}
__device__ __forceinline__ float gelu_f(float x) {
  float g, dg;
  gelu_both(x, g, dg);
  return g;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float g, dg;
  gelu_both(x, g, dg);
  return dg;
}
#endif  // __CUDACC__

}  // namespace drin
