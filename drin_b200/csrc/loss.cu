// TripletLoss forward + backward and the threshold top-k metric on the device.
//
// Reference: common/utils.py:35-43 (TripletLoss), :60-66 (TopkAccuracy.update).  The loss couples the
// whole batch -- line 42 subtracts the ENTIRE [B, C-1] score matrix from every mention's positive:
//     loss = k * sum_i sum_{b,c} max(s[b,c] - p_i + margin, 0),   k = 1 / (B * B * (C-1)),
//     p_i  = sum_c s[i,c] * y[i,c]   (0 when the gold entity is not among the candidates)
//     dL/ds[b,c] = k * ( A[b,c] - y[b,c] * N_b ),
//        A[b,c] = #{ i : s[b,c] - p_i + margin > 0 },  N_b = #{ (b',c') : s[b',c'] - p_b + margin > 0 }.
// Under data parallelism every rank holds the gathered global score matrix and evaluates its own rows.
// Counts come from binary searches in the sorted positives (one small CUB radix sort) plus an integer histogram of the
// global scores; the hinge sum is reduced in a fixed order in double precision, so the result is bit-reproducible.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace drin {

static constexpr int TL_THREADS = 256;

// O(N log N) evaluation with ONE small sort: with the positives p sorted ascending (prefix sums P),
//   A[b,c]      = lower_bound(p_sorted, s + m)                       (# p_i < s + m)
//   hinge[b,c]  = A * (s + m) - P[A]                                 (sum_i max(s - p_i + m, 0), in double)
//   N_b         = # global scores s' with p_b < s' + m = sum_{j > pos_b} hist[j],
//                 hist[j] = # global scores whose A equals j, pos_b = lower_bound(p_sorted, p_b)
// so every rank sorts only the B_glob positives, bins all global scores with integer atomics (exact, order
// independent) and pays O((B_glob C) log B_glob) -- flat when the global batch grows with the number of GPUs.
struct TripletScratch {
  float* p;            // [B]
  float* p_sorted;     // [B]
  double* prefix;      // [B + 1]
  int* hist;           // [B + 1]  -> inclusive cumulative counts after triplet_cum_kernel
  double* partial;     // [blocks]
  void* cub_temp;
  size_t cub_bytes;
};

static TripletScratch carve_triplet(void* scratch, int B, int C, size_t* bytes) {
  TripletScratch t;
  size_t off = 0;
  char* base = static_cast<char*>(scratch);
  auto take = [&](size_t nbytes) {
    char* ptr = base ? base + off : nullptr;
    off = align_up(off + nbytes, 256);
    return ptr;
  };
  t.p = reinterpret_cast<float*>(take(sizeof(float) * B));
  t.p_sorted = reinterpret_cast<float*>(take(sizeof(float) * B));
  t.prefix = reinterpret_cast<double*>(take(sizeof(double) * (B + 1)));
  t.hist = reinterpret_cast<int*>(take(sizeof(int) * (B + 1)));
  t.partial = reinterpret_cast<double*>(take(sizeof(double) * 65536));
  t.cub_bytes = 2 * (size_t)B * sizeof(float) + (4u << 20);   // upper bound of the radix-sort temp storage (checked at run time)
  t.cub_temp = take(t.cub_bytes);
  (void)C;
  if (bytes) *bytes = off;
  return t;
}

size_t triplet_scratch_bytes(int B, int C) {
  size_t bytes = 0;
  carve_triplet(nullptr, B, C, &bytes);
  return bytes;
}

// p[i] = sum_c s[i,c] y[i,c] over the real candidates (the gold slot is sliced off, utils.py:36-37)
__global__ void triplet_pos_kernel(const float* __restrict__ s, const unsigned char* __restrict__ y, int B, int C,
                                   float* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float acc = 0.f;
  for (int c = 0; c < C - 1; ++c) acc += s[(long long)i * C + c] * (float)y[(long long)i * (C - 1) + c];
  p[i] = acc;
}

// single-block exclusive prefix sums in double: prefix[j] = sum_{i<j} p_sorted[i], j = 0..B; also clears hist
__global__ void __launch_bounds__(1024) triplet_prefix_kernel(const float* __restrict__ ps, int B, double* __restrict__ prefix,
                                                               int* __restrict__ hist) {
  typedef cub::BlockScan<double, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int t = threadIdx.x;
  for (int i = t; i <= B; i += 1024) hist[i] = 0;
  const int per = (B + 1023) / 1024;
  const int i0 = min(B, t * per), i1 = min(B, i0 + per);
  double acc = 0.0;
  for (int i = i0; i < i1; ++i) acc += (double)ps[i];
  double run;
  Scan(tmp).ExclusiveSum(acc, run);          // fixed tree order: deterministic
  for (int i = i0; i < i1; ++i) {
    prefix[i] = run;
    run += (double)ps[i];
  }
  if (i0 < B && i1 == B) prefix[B] = run;      // the thread that owns the last element
}

__device__ __forceinline__ int lower_bound_f(const float* __restrict__ a, int n, float v) {   // # a[i] < v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// one thread per GLOBAL real-candidate score: bin it (hist[A] += 1, warp-aggregated integer atomics); the threads of
// the local rows also emit k * A into dscores and their hinge term
__global__ void __launch_bounds__(TL_THREADS) triplet_count_kernel(const float* __restrict__ s,
                                                                   const float* __restrict__ ps,
                                                                   const double* __restrict__ prefix, int B, int C,
                                                                   int row0, int rows, float margin, float k,
                                                                   float* __restrict__ dscores, int* __restrict__ hist,
                                                                   double* __restrict__ partial) {
  __shared__ double sred[TL_THREADS / 32];
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // over B * (C-1)
  const long long total = (long long)B * (C - 1);
  double hinge = 0.0;
  int A = -1;
  if (e < total) {
    const int b = (int)(e / (C - 1));
    const int c = (int)(e - (long long)b * (C - 1));
    const float sm = s[(long long)b * C + c] + margin;
    A = lower_bound_f(ps, B, sm);                      // s - p_i + margin > 0  <=>  p_i < s + margin
    if (b >= row0 && b < row0 + rows) {
      hinge = (double)A * (double)sm - prefix[A];
      if (sm != sm) hinge = (double)sm;                // NaN scores poison the loss like upstream
      dscores[(long long)(b - row0) * C + c] = k * (float)A;
    }
  }
  // lanes with the same bin combine into one atomic (at init every score lands in the same bin)
  const unsigned peers = __match_any_sync(0xffffffffu, A);
  if (A >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + A, __popc(peers));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hinge += __shfl_xor_sync(0xffffffffu, hinge, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = hinge;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < TL_THREADS / 32; ++w) t += sred[w];
    partial[blockIdx.x] = t;
  }
}

// single block: hist[j] -> inclusive cumulative counts cum[j] = sum_{i <= j} hist[i], j = 0..B
__global__ void __launch_bounds__(1024) triplet_cum_kernel(int* __restrict__ hist, int B) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int t = threadIdx.x, n = B + 1;
  const int per = (n + 1023) / 1024;
  const int i0 = min(n, t * per), i1 = min(n, i0 + per);
  int acc = 0;
  for (int i = i0; i < i1; ++i) acc += hist[i];
  int run;
  Scan(tmp).ExclusiveSum(acc, run);
  for (int i = i0; i < i1; ++i) {
    run += hist[i];
    hist[i] = run;
  }
}

__global__ void triplet_finish_kernel(const unsigned char* __restrict__ y, const float* __restrict__ p,
                                      const float* __restrict__ ps, const int* __restrict__ cum, int B, int C, int row0,
                                      int rows, float k, float* __restrict__ dscores, const double* __restrict__ partial,
                                      int nblocks, float* __restrict__ loss) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // over rows * C
  if (e < (long long)rows * C) {
    const int bl = (int)(e / C), c = (int)(e - (long long)bl * C);
    if (c == C - 1) {
      dscores[e] = 0.f;                                     // the gold slot is sliced off (utils.py:36-37)
    } else if (y[(long long)(row0 + bl) * (C - 1) + c]) {
      // N_b = # global scores with p_b < s' + margin = total - cum[pos_b]
      const int pos = lower_bound_f(ps, B, p[row0 + bl]);
      const int n_b = cum[B] - cum[pos];
      dscores[e] -= k * (float)y[(long long)(row0 + bl) * (C - 1) + c] * (float)n_b;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // only the blocks that cover local rows hold non-zero hinge partials; all are summed in block order
    double t = 0.0;
    for (int i = 0; i < nblocks; ++i) t += partial[i];
    loss[0] = (float)(t * (double)k);
  }
}

int triplet_loss(cudaStream_t stream, const float* scores, const unsigned char* labels, int B, int C, int row0,
                 int rows, float margin, float* loss, float* dscores, void* scratch) {
  prof::Scope prof_scope(stream, prof::LOSS);
  if (B <= 0 || C <= 1 || row0 < 0 || rows <= 0 || row0 + rows > B)
    return fail(DRIN_ERR_ARG, "triplet_loss: bad shape B=%d C=%d row0=%d rows=%d", B, C, row0, rows);
  if (!scores || !labels || !loss || !dscores || !scratch) return fail(DRIN_ERR_ARG, "triplet_loss: null argument");
  if ((long long)B * (C - 1) > 0x7fffffffLL) return fail(DRIN_ERR_ARG, "triplet_loss: batch too large");
  TripletScratch t = carve_triplet(scratch, B, C, nullptr);
  const long long ns = (long long)B * (C - 1);
  const float k = (float)(1.0 / ((double)B * (double)B * (double)(C - 1)));
  triplet_pos_kernel<<<(B + 255) / 256, 256, 0, stream>>>(scores, labels, B, C, t.p);
  DRIN_LAUNCH_CHECK();
  size_t need = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, need, t.p, t.p_sorted, B);
  if (need > t.cub_bytes) return fail(DRIN_ERR_WORKSPACE, "triplet_loss: sort temp %zu > %zu", need, t.cub_bytes);
  size_t tmp = t.cub_bytes;
  DRIN_CUDA(cub::DeviceRadixSort::SortKeys(t.cub_temp, tmp, t.p, t.p_sorted, B, 0, 32, stream));
  count_launch();
  triplet_prefix_kernel<<<1, 1024, 0, stream>>>(t.p_sorted, B, t.prefix, t.hist);
  DRIN_LAUNCH_CHECK();
  const int cblocks = (int)((ns + TL_THREADS - 1) / TL_THREADS);
  if (cblocks > 65536) return fail(DRIN_ERR_ARG, "triplet_loss: too many scores (%lld)", ns);
  triplet_count_kernel<<<cblocks, TL_THREADS, 0, stream>>>(scores, t.p_sorted, t.prefix, B, C, row0, rows, margin, k,
                                                           dscores, t.hist, t.partial);
  DRIN_LAUNCH_CHECK();
  triplet_cum_kernel<<<1, 1024, 0, stream>>>(t.hist, B);
  DRIN_LAUNCH_CHECK();
  const long long fe = (long long)rows * C;
  triplet_finish_kernel<<<(int)((fe + 255) / 256), 256, 0, stream>>>(labels, t.p, t.p_sorted, t.hist, B, C, row0, rows, k,
                                                                     dscores, t.partial, cblocks, loss);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// TopkAccuracy.update: gold is a hit for k when fewer than k real candidates score strictly higher
// (equivalent to gold >= k-th largest, ties count as hits).
__global__ void topk_hits_kernel(const float* __restrict__ s, const unsigned char* __restrict__ y, int B, int C,
                                 const int* __restrict__ topk, int nk, unsigned long long* __restrict__ hits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int g = 0; g < C - 1; ++g) {
    const unsigned char w = y[(long long)b * (C - 1) + g];
    if (!w) continue;
    const float sg = s[(long long)b * C + g];
    int gt = 0;
    for (int c = 0; c < C - 1; ++c) gt += s[(long long)b * C + c] > sg;
    for (int j = 0; j < nk; ++j)
      if (gt < topk[j]) atomicAdd(hits + j, (unsigned long long)w);
  }
}

int topk_hits(cudaStream_t stream, const float* scores, const unsigned char* labels, int B, int C, const int* topk,
              int nk, long long* hits) {
  if (nk <= 0 || nk > 16) return fail(DRIN_ERR_ARG, "topk_hits: 1..16 values of k");
  static int* d_topk = nullptr;
  if (!d_topk) DRIN_CUDA(cudaMalloc(&d_topk, 16 * sizeof(int)));
  DRIN_CUDA(cudaMemcpyAsync(d_topk, topk, nk * sizeof(int), cudaMemcpyHostToDevice, stream));
  topk_hits_kernel<<<(B + 127) / 128, 128, 0, stream>>>(scores, labels, B, C, d_topk, nk,
                                                        reinterpret_cast<unsigned long long*>(hits));
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
