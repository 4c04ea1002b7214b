// Host interface of the tcgen05 GEMM family (gemm_tcgen05.cu).
#pragma once
#include "common.cuh"

namespace drin {

// One GEMM operand: a row-major bf16 matrix [rows, cols] with leading dimension ld (elements), given
// as one plane (bf16 mode) or two planes hi/lo (split-bf16, fp32-parity mode: value = hi + lo).
struct Operand {
  const bf16* hi = nullptr;
  const bf16* lo = nullptr;   // nullptr in single-plane mode
  long long rows = 0;
  int cols = 0;
  int ld = 0;
};

enum GemmLayout : int {
  // C[M,N] = A[M,K] * B[N,K]^T        forward  (x W^T):   A rows=M cols=K,  B rows=N cols=K
  GEMM_NT = 0,
  // C[M,N] = A[M,K] * B[K,N]          data grad (dY W):   A rows=M cols=K,  B rows=K cols=N
  GEMM_NN = 1,
  // C[M,N] = A[K,M]^T * B[K,N]        weight grad (dY^T X): A rows=K cols=M, B rows=K cols=N
  GEMM_TN = 2,
};

struct GemmEpilogue {
  float* C = nullptr;          // fp32 output [M, ldc] (may be nullptr when only planes are wanted)
  int ldc = 0;
  const float* bias = nullptr; // per output column, or nullptr
  bf16* out_hi = nullptr;      // optional split-bf16 copy of the output (A operand of a following GEMM)
  bf16* out_lo = nullptr;
  int ld_planes = 0;
};

// A split-K reduction that has not been run yet (see gemm_tcgen05 `defer`): C[M, ldc] = bias + sum_s partial[s]
struct SplitKJob {
  const float* partial;
  long long split_stride;   // elements between slices
  float* C;
  const float* bias;
  long long M;
  int N, ldc, ksplit;
};
struct SplitKJobs {
  static constexpr int MAX = 40;
  SplitKJob job[MAX];
  int count = 0;
};
// one launch that finishes every deferred split-K GEMM of a backward pass
int splitk_reduce_multi(cudaStream_t stream, const SplitKJobs& jobs);

// ksplit > 1: the contraction is cut into ksplit slices, slice s writes partial[s][M][ldc] (fp32, no
// bias) and a second kernel reduces the slices deterministically into C (adding bias if given).
int gemm_tcgen05(cudaStream_t stream, GemmLayout layout, const Operand& A, const Operand& B, long long M, int N,
                 long long K, const GemmEpilogue& ep, int ksplit = 1, float* partial = nullptr,
                 SplitKJobs* defer = nullptr);   // defer: append the reduction to the list instead of launching it

// fp32 CUDA-core reference of the same contract (bring-up / unit tests only; never on the product path)
int gemm_reference_simt(cudaStream_t stream, GemmLayout layout, const Operand& A, const Operand& B, long long M,
                        int N, long long K, const GemmEpilogue& ep);

// MN-major descriptor strides are runtime values so a bring-up test can probe them; 0 = defaults.
void gemm_debug_set_mn_desc(int lbo_bytes, int sbo_bytes);

}  // namespace drin
