// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fp32_pipe scripts/ubench/fp32_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a), "f"(b));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
  unsigned long long acc[ILP];
  float2 av = make_float2(a, a), bv = make_float2(b, b);
  unsigned long long A = *reinterpret_cast<unsigned long long*>(&av), B = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float2 t = make_float2(threadIdx.x * 0.001f + i, i);
    acc[i] = *reinterpret_cast<unsigned long long*>(&t);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = ffma2(A, B, acc[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float2 t = *reinterpret_cast<float2*>(&acc[i]);
    s += t.x + t.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 4096;
  constexpr int ILP = 8;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = 256, blocks = sms * warps * 32 / threads;
    float m1 = time_ms([&] { k_ffma<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    float m2 = time_ms([&] { k_ffma2<ILP><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    const double n = (double)blocks * threads * iters * ILP;
    printf("warps/SM %2d: FFMA %.1f TFLOP/s (%.3f ms)   FFMA2 %.1f TFLOP/s (%.3f ms)\n", warps, 2 * n / m1 / 1e9, m1,
           4 * n / m2 / 1e9, m2);
  }
  return 0;
}
