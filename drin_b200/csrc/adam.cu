// torch.optim.Adam (reference train.py:55-56: lr 1e-3, betas (0.9, 0.999), eps 1e-8, no weight decay) as one
// vectorised pass over a flat fp32 parameter buffer: 4 streams read, 3 written, 16-byte accesses.
#include "kernels.cuh"

namespace drin {

__global__ void adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                            float4* __restrict__ v, const unsigned char* __restrict__ skip, long long n4, float lr_bc1,
                            float inv_sqrt_bc2, float b1, float b2, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (skip && skip[i * 4]) continue;                  // parameters are multiples of 4 elements
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    float* pa = reinterpret_cast<float*>(&pp);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
    const float* ga = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
      va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= lr_bc1 * (ma[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

// Same update with the 1-based step count read from DEVICE memory, so that a captured CUDA graph of the train step
// can be replayed (the bias corrections change every step).
__global__ void adam_dev_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                     float4* __restrict__ v, const unsigned char* __restrict__ skip, long long n4,
                                     const int* __restrict__ step_dev, float lr, float b1, float b2, float eps) {
  const int step = *step_dev;
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  const float lr_bc1 = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (skip && skip[i * 4]) continue;
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = g[i];
    float* pa = reinterpret_cast<float*>(&pp);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
    const float* ga = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
      va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= lr_bc1 * (ma[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

int adam_step_dev(cudaStream_t stream, float* p, const float* g, float* m, float* v, const unsigned char* skip,
                  long long n, const int* step_dev, float lr, float b1, float b2, float eps) {
  prof::Scope prof_scope(stream, prof::ADAM);
  if (n % 4) return fail(DRIN_ERR_ARG, "adam_step: n must be a multiple of 4");
  if (!step_dev) return fail(DRIN_ERR_ARG, "adam_step_dev: step pointer is null");
  const long long n4 = n / 4;
  const int grid = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  adam_dev_step_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                                 reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), skip, n4,
                                                 step_dev, lr, b1, b2, eps);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

int adam_step(cudaStream_t stream, float* p, const float* g, float* m, float* v, const unsigned char* skip, long long n,
              int step, float lr, float b1, float b2, float eps) {
  prof::Scope prof_scope(stream, prof::ADAM);
  if (n % 4) return fail(DRIN_ERR_ARG, "adam_step: n must be a multiple of 4");
  if (step < 1) return fail(DRIN_ERR_ARG, "adam_step: step is 1-based");
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  const long long n4 = n / 4;
  const int grid = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  adam_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                        reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), skip, n4,
                                        (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), b1, b2, eps);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
