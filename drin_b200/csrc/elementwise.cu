// Small memory-bound helpers: fp32 -> split-bf16 planes, weight preparation.
#include "kernels.cuh"

namespace drin {

// x[n] (fp32) -> hi[n], lo[n] (bf16) with x ~= hi + lo.  n4 = n / 4 (callers guarantee n % 4 == 0).
__global__ void split_planes_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                    long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(x + i);
    bf16 h0, l0, h1, l1, h2, l2, h3, l3;
    split_bf16(v.x, h0, l0);
    split_bf16(v.y, h1, l1);
    split_bf16(v.z, h2, l2);
    split_bf16(v.w, h3, l3);
    hi[i] = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    if (lo) lo[i] = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
  }
}

int split_planes(cudaStream_t stream, const float* x, bf16* hi, bf16* lo, long long n) {
  prof::Scope prof_scope(stream, prof::PREP);
  if (n % 4) return fail(DRIN_ERR_ARG, "split_planes: n must be a multiple of 4 (got %lld)", n);
  if (n == 0) return DRIN_OK;
  const long long n4 = n / 4;
  const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  split_planes_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(hi),
                                                reinterpret_cast<uint2*>(lo), n4);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
