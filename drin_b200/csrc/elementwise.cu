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

// All weight matrices of a step in ONE launch: blockIdx.y selects the segment.
__global__ void split_planes_multi_kernel(const SplitJobs jobs) {
  const SplitJob j = jobs.job[blockIdx.y];
  const float4* x = reinterpret_cast<const float4*>(j.x);
  uint2* hi = reinterpret_cast<uint2*>(j.hi);
  uint2* lo = reinterpret_cast<uint2*>(j.lo);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < j.n4; i += (long long)gridDim.x * blockDim.x) {
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

int split_planes_multi(cudaStream_t stream, const SplitJobs& jobs) {
  prof::Scope prof_scope(stream, prof::PREP);
  if (jobs.count <= 0) return DRIN_OK;
  if (jobs.count > SplitJobs::MAX) return fail(DRIN_ERR_ARG, "split_planes_multi: too many segments");
  for (int i = 0; i < jobs.count; ++i)
    if (!jobs.job[i].x || !jobs.job[i].hi) return fail(DRIN_ERR_ARG, "null parameter pointer");
  split_planes_multi_kernel<<<dim3(148, jobs.count), 256, 0, stream>>>(jobs);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
