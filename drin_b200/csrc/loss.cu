// TripletLoss forward + backward and the threshold top-k metric on the device.
//
// Reference: common/utils.py:35-43 (TripletLoss), :60-66 (TopkAccuracy.update).  The loss couples the
// whole batch -- line 42 subtracts the ENTIRE [B, C-1] score matrix from every mention's positive:
//     loss = k * sum_i sum_{b,c} max(s[b,c] - p_i + margin, 0),   k = 1 / (B * B * (C-1)),
//     p_i  = sum_c s[i,c] * y[i,c]   (0 when the gold entity is not among the candidates)
//     dL/ds[b,c] = k * ( A[b,c] - y[b,c] * N_b ),
//        A[b,c] = #{ i : s[b,c] - p_i + margin > 0 },  N_b = #{ (b',c') : s[b',c'] - p_b + margin > 0 }.
// Under data parallelism every rank holds the gathered global score matrix and evaluates its own rows.
// Counts come from binary searches in segment-sorted positives plus integer histograms of the global scores; the
// hinge sum is reduced in a fixed order in double precision, so the result is bit-reproducible.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace drin {

static constexpr int TL_THREADS = 256;

// O(N log N) evaluation without a global sort.  The B_glob positives are cut into segments of TL_SEG = 4096 (a rank's
// shard is one segment at the benchmark size) and every segment is sorted by ONE CTA (cub::BlockRadixSort), with its
// prefix sums P_g in double.  For a score s (m = margin):
//   A[b,c]      = sum_g lower_bound(seg_g, s + m)                    (# p_i < s + m)
//   hinge[b,c]  = sum_g ( A_g * (s + m) - P_g[A_g] )                 (sum_i max(s - p_i + m, 0), in double)
//   N_b         = # global scores s' with p_b < s' + m = sum_{j > pos_b} hist_g[j] for the segment g that holds p_b,
//                 hist_g[j] = # global scores whose lower_bound in seg_g equals j, pos_b = lower_bound(seg_g, p_b)
// so a rank bins all global scores only against its OWN segments (integer atomics: exact, order independent) and
// searches all segments only for its own scores: five launches whatever the number of GPUs.
static constexpr int TL_SEG = 4096;
static constexpr int TL_SORT_THREADS = 1024;
static constexpr int TL_SORT_ITEMS = TL_SEG / TL_SORT_THREADS;

struct TripletScratch {
  float* p;            // [B]
  float* p_sorted;     // [S][TL_SEG]       sorted per segment, padded with +inf
  double* prefix;      // [S][TL_SEG + 1]   exclusive prefix sums per segment
  int* hist;           // [S][TL_SEG + 1]   -> inclusive cumulative counts after triplet_cum_kernel (local segments)
  double* partial;     // [blocks]
};

static TripletScratch carve_triplet(void* scratch, int B, int C, size_t* bytes) {
  TripletScratch t;
  size_t off = 0;
  char* base = static_cast<char*>(scratch);
  const size_t S = ((size_t)B + TL_SEG - 1) / TL_SEG;
  auto take = [&](size_t nbytes) {
    char* ptr = base ? base + off : nullptr;
    off = align_up(off + nbytes, 256);
    return ptr;
  };
  t.p = reinterpret_cast<float*>(take(sizeof(float) * B));
  t.p_sorted = reinterpret_cast<float*>(take(sizeof(float) * S * TL_SEG));
  t.prefix = reinterpret_cast<double*>(take(sizeof(double) * S * (TL_SEG + 1)));
  t.hist = reinterpret_cast<int*>(take(sizeof(int) * S * (TL_SEG + 1)));
  t.partial = reinterpret_cast<double*>(take(sizeof(double) * 65536));
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

// one CTA per segment: sort its positives, exclusive prefix sums in double, clear its histogram
__global__ void __launch_bounds__(TL_SORT_THREADS) triplet_segsort_kernel(const float* __restrict__ p, int B,
                                                                          float* __restrict__ p_sorted,
                                                                          double* __restrict__ prefix,
                                                                          int* __restrict__ hist) {
  typedef cub::BlockRadixSort<float, TL_SORT_THREADS, TL_SORT_ITEMS> Sort;
  typedef cub::BlockScan<double, TL_SORT_THREADS> Scan;
  __shared__ union {
    typename Sort::TempStorage sort;
    typename Scan::TempStorage scan;
  } tmp;
  const int g = blockIdx.x, t = threadIdx.x;
  const int base = g * TL_SEG, n = min(TL_SEG, B - base);
  float keys[TL_SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < TL_SORT_ITEMS; ++i) {
    const int idx = t * TL_SORT_ITEMS + i;
    keys[i] = idx < n ? p[base + idx] : __int_as_float(0x7f800000);     // +inf padding sorts last
  }
  Sort(tmp.sort).Sort(keys);                                             // blocked arrangement, ascending
  __syncthreads();
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < TL_SORT_ITEMS; ++i) {
    const int idx = t * TL_SORT_ITEMS + i;
    p_sorted[(long long)g * TL_SEG + idx] = keys[i];
    if (idx < n) acc += (double)keys[i];
  }
  double run;
  Scan(tmp.scan).ExclusiveSum(acc, run);                                 // fixed tree order: deterministic
  double* pf = prefix + (long long)g * (TL_SEG + 1);
#pragma unroll
  for (int i = 0; i < TL_SORT_ITEMS; ++i) {
    const int idx = t * TL_SORT_ITEMS + i;
    if (idx < n) {
      pf[idx] = run;
      run += (double)keys[i];
      if (idx == n - 1) pf[n] = run;
    }
  }
  if (n == 0 && t == 0) pf[0] = 0.0;
  int* h = hist + (long long)g * (TL_SEG + 1);
  for (int i = t; i <= TL_SEG; i += TL_SORT_THREADS) h[i] = 0;
}

__device__ __forceinline__ int lower_bound_f(const float* __restrict__ a, int n, float v) {   // # a[i] < v
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// one thread per GLOBAL real-candidate score: bin it against the segments [g0, g1] that hold the local rows
// (hist_g[A_g] += 1, warp-aggregated integer atomics); the threads of the local rows also search the other segments
// and emit k * A into dscores and their hinge term
__global__ void __launch_bounds__(TL_THREADS) triplet_count_kernel(const float* __restrict__ s,
                                                                   const float* __restrict__ ps,
                                                                   const double* __restrict__ prefix, int B, int C,
                                                                   int row0, int rows, int g0, int g1, int S,
                                                                   float margin, float k,
                                                                   float* __restrict__ dscores, int* __restrict__ hist,
                                                                   double* __restrict__ partial) {
  __shared__ double sred[TL_THREADS / 32];
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // over B * (C-1)
  const long long total = (long long)B * (C - 1);
  const bool valid = e < total;
  double hinge = 0.0;
  int b = 0, c = 0;
  float sm = 0.f;
  if (valid) {
    b = (int)(e / (C - 1));
    c = (int)(e - (long long)b * (C - 1));
    sm = s[(long long)b * C + c] + margin;
  }
  const bool local = valid && b >= row0 && b < row0 + rows;
  int A = 0;
  for (int g = 0; g < S; ++g) {
    const bool mine = g >= g0 && g <= g1;
    if (!mine && !__any_sync(0xffffffffu, local)) continue;                  // warp-uniform skip
    int j = -1;
    if (valid && (mine || local)) {
      const int n = min(TL_SEG, B - g * TL_SEG);
      j = lower_bound_f(ps + (long long)g * TL_SEG, n, sm);              // s - p_i + margin > 0  <=>  p_i < s + margin
      if (local) {
        A += j;
        hinge += (double)j * (double)sm - prefix[(long long)g * (TL_SEG + 1) + j];
      }
    }
    if (mine) {   // lanes with the same bin combine into one atomic (at init every score lands in the same bin)
      const unsigned peers = __match_any_sync(0xffffffffu, j);
      if (j >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + (long long)g * (TL_SEG + 1) + j, __popc(peers));
    }
  }
  if (local) {
    if (sm != sm) hinge = (double)sm;                  // NaN scores poison the loss like upstream
    dscores[(long long)(b - row0) * C + c] = k * (float)A;
  }
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

// one CTA per local segment: hist_g[j] -> inclusive cumulative counts
__global__ void __launch_bounds__(1024) triplet_cum_kernel(int* __restrict__ hist, int g0) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int* h = hist + (long long)(g0 + blockIdx.x) * (TL_SEG + 1);
  const int t = threadIdx.x, n = TL_SEG + 1;
  const int per = (n + 1023) / 1024;
  const int i0 = min(n, t * per), i1 = min(n, i0 + per);
  int acc = 0;
  for (int i = i0; i < i1; ++i) acc += h[i];
  int run;
  Scan(tmp).ExclusiveSum(acc, run);
  for (int i = i0; i < i1; ++i) {
    run += h[i];
    h[i] = run;
  }
}

__global__ void triplet_finish_kernel(const unsigned char* __restrict__ y, const float* __restrict__ p,
                                      const float* __restrict__ ps, const int* __restrict__ cum, int B, int C, int row0,
                                      int rows, float k, float* __restrict__ dscores, const double* __restrict__ partial,
                                      int pb0, int pb1, float* __restrict__ loss) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // over rows * C
  if (e < (long long)rows * C) {
    const int bl = (int)(e / C), c = (int)(e - (long long)bl * C);
    if (c == C - 1) {
      dscores[e] = 0.f;                                     // the gold slot is sliced off (utils.py:36-37)
    } else if (y[(long long)(row0 + bl) * (C - 1) + c]) {
      // N_b = # global scores with p_b < s' + margin, from the cumulative histogram of the segment that holds p_b
      const int b = row0 + bl, g = b / TL_SEG;
      const int n = min(TL_SEG, B - g * TL_SEG);
      const int* cg = cum + (long long)g * (TL_SEG + 1);
      const int pos = lower_bound_f(ps + (long long)g * TL_SEG, n, p[b]);
      const int n_b = cg[TL_SEG] - cg[pos];
      dscores[e] -= k * (float)y[(long long)b * (C - 1) + c] * (float)n_b;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // only the count-kernel blocks [pb0, pb1] that cover local rows hold non-zero hinge partials: summed in block order
    // (the cost of this serial tail does not grow with the number of ranks)
    double t = 0.0;
    for (int i = pb0; i <= pb1; ++i) t += partial[i];
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
  const int S = (B + TL_SEG - 1) / TL_SEG;
  const int g0 = row0 / TL_SEG, g1 = (row0 + rows - 1) / TL_SEG;       // segments that hold the local rows
  const float k = (float)(1.0 / ((double)B * (double)B * (double)(C - 1)));
  triplet_pos_kernel<<<(B + 255) / 256, 256, 0, stream>>>(scores, labels, B, C, t.p);
  DRIN_LAUNCH_CHECK();
  triplet_segsort_kernel<<<S, TL_SORT_THREADS, 0, stream>>>(t.p, B, t.p_sorted, t.prefix, t.hist);
  DRIN_LAUNCH_CHECK();
  const int cblocks = (int)((ns + TL_THREADS - 1) / TL_THREADS);
  if (cblocks > 65536) return fail(DRIN_ERR_ARG, "triplet_loss: too many scores (%lld)", ns);
  triplet_count_kernel<<<cblocks, TL_THREADS, 0, stream>>>(scores, t.p_sorted, t.prefix, B, C, row0, rows, g0, g1, S, margin,
                                                           k, dscores, t.hist, t.partial);
  DRIN_LAUNCH_CHECK();
  triplet_cum_kernel<<<g1 - g0 + 1, 1024, 0, stream>>>(t.hist, g0);
  DRIN_LAUNCH_CHECK();
  const long long fe = (long long)rows * C;
  const int pb0 = (int)((long long)row0 * (C - 1) / TL_THREADS);
  const int pb1 = (int)((((long long)(row0 + rows)) * (C - 1) - 1) / TL_THREADS);
  triplet_finish_kernel<<<(int)((fe + 255) / 256), 256, 0, stream>>>(labels, t.p, t.p_sorted, t.hist, B, C, row0, rows, k,
                                                                     dscores, t.partial, pb0, pb1, loss);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// TopkAccuracy.update: gold is a hit for k when fewer than k real candidates score strictly higher
// (equivalent to gold >= k-th largest, ties count as hits).
struct TopkList {      // the k values travel by value in the launch parameters: no device buffer, no host copy to order
  int k[16];
};

__global__ void topk_hits_kernel(const float* __restrict__ s, const unsigned char* __restrict__ y, int B, int C,
                                 const TopkList topk_list, int nk, unsigned long long* __restrict__ hits) {
  const int* topk = topk_list.k;
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
  TopkList list{};
  for (int j = 0; j < nk; ++j) list.k[j] = topk[j];
  topk_hits_kernel<<<(B + 127) / 128, 128, 0, stream>>>(scores, labels, B, C, list, nk,
                                                        reinterpret_cast<unsigned long long*>(hits));
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
