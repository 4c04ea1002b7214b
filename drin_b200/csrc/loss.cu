// TripletLoss forward + backward and the threshold top-k metric on the device.
//
// Reference: common/utils.py:35-43 (TripletLoss), :60-66 (TopkAccuracy.update).  The loss couples the
// whole batch -- line 42 subtracts the ENTIRE [B, C-1] score matrix from every mention's positive:
//     loss = k * sum_i sum_{b,c} max(s[b,c] - p_i + margin, 0),   k = 1 / (B * B * (C-1)),
//     p_i  = sum_c s[i,c] * y[i,c]   (0 when the gold entity is not among the candidates)
//     dL/ds[b,c] = k * ( A[b,c] - y[b,c] * N_b ),
//        A[b,c] = #{ i : s[b,c] - p_i + margin > 0 },  N_b = #{ (b',c') : s[b',c'] - p_b + margin > 0 }.
// Under data parallelism every rank holds the gathered global score matrix and evaluates its own rows:
// counts are integers (order-independent atomics) and the hinge sum is reduced in a fixed order in
// double precision, so the result is bit-reproducible.
#include "kernels.cuh"

namespace drin {

static constexpr int TL_THREADS = 256;
static constexpr int TL_TILE = 2048;

struct TripletScratch {
  float* p;          // [B]
  int* n;            // [B]
  double* partial;   // [blocks]
};

static TripletScratch carve_triplet(void* scratch, int B, size_t* bytes) {
  TripletScratch t;
  size_t off = 0;
  char* base = static_cast<char*>(scratch);
  t.p = reinterpret_cast<float*>(base + off);
  off = align_up(off + sizeof(float) * B, 256);
  t.n = reinterpret_cast<int*>(base + off);
  off = align_up(off + sizeof(int) * B, 256);
  t.partial = reinterpret_cast<double*>(base + off);
  off = align_up(off + sizeof(double) * 65536, 256);
  if (bytes) *bytes = off;
  return t;
}

size_t triplet_scratch_bytes(int B, int C) {
  (void)C;
  size_t bytes = 0;
  carve_triplet(nullptr, B, &bytes);
  return bytes;
}

__global__ void triplet_pos_kernel(const float* __restrict__ s, const unsigned char* __restrict__ y, int B, int C,
                                   float* __restrict__ p, int* __restrict__ n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float acc = 0.f;
  for (int c = 0; c < C - 1; ++c) acc += s[(long long)i * C + c] * (float)y[(long long)i * (C - 1) + c];
  p[i] = acc;
  n[i] = 0;
}

// one thread per local score element; loops over all positives
__global__ void __launch_bounds__(TL_THREADS) triplet_count_kernel(const float* __restrict__ s,
                                                                   const float* __restrict__ p, int B, int C, int row0,
                                                                   int rows, float margin, float k,
                                                                   float* __restrict__ dscores,
                                                                   double* __restrict__ partial) {
  __shared__ float sp[TL_TILE];
  __shared__ double sred[TL_THREADS / 32];
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // over rows * (C-1)
  const long long total = (long long)rows * (C - 1);
  const bool live = e < total;
  const int bl = live ? (int)(e / (C - 1)) : 0;
  const int c = live ? (int)(e - (long long)bl * (C - 1)) : 0;
  const float sm = live ? s[(long long)(row0 + bl) * C + c] + margin : 0.f;
  int cnt = 0;
  float hinge = 0.f;
  double hinge_d = 0.0;
  for (int t0 = 0; t0 < B; t0 += TL_TILE) {
    const int nt = min(TL_TILE, B - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += blockDim.x) sp[i] = p[t0 + i];
    __syncthreads();
    if (live) {
      for (int i = 0; i < nt; ++i) {
        const float d = sm - sp[i];
        cnt += d > 0.f;
        hinge += fmaxf(d, 0.f);
      }
      hinge_d += (double)hinge;     // flush the fp32 tile sum into double every tile
      hinge = 0.f;
    }
  }
  if (live) dscores[(long long)bl * C + c] = k * (float)cnt;
  // fixed-order block reduction of the hinge sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hinge_d += __shfl_xor_sync(0xffffffffu, hinge_d, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = hinge_d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < TL_THREADS / 32; ++w) t += sred[w];
    partial[blockIdx.x] = t;
  }
}

// N_b for local b: blockIdx.y = tile of local rows (one thread per row), blockIdx.x = chunk of global scores
__global__ void __launch_bounds__(TL_THREADS) triplet_n_kernel(const float* __restrict__ s, const float* __restrict__ p,
                                                               int B, int C, int row0, int rows, float margin,
                                                               int* __restrict__ n) {
  __shared__ float ss[TL_TILE];
  const long long total = (long long)B * (C - 1);
  const long long e0 = (long long)blockIdx.x * TL_TILE;
  const int ne = (int)min((long long)TL_TILE, total - e0);
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const long long e = e0 + i;
    const long long b = e / (C - 1);
    ss[i] = s[b * C + (e - b * (C - 1))];
  }
  __syncthreads();
  const int bl = blockIdx.y * blockDim.x + threadIdx.x;
  if (bl >= rows) return;
  const float thr = p[row0 + bl] - margin;
  int cnt = 0;
  for (int i = 0; i < ne; ++i) cnt += ss[i] > thr;      // s - p + margin > 0
  if (cnt) atomicAdd(n + row0 + bl, cnt);
}

__global__ void triplet_finish_kernel(const unsigned char* __restrict__ y, const int* __restrict__ n, int C, int row0,
                                      int rows, float k, float* __restrict__ dscores, const double* __restrict__ partial,
                                      int nblocks, float* __restrict__ loss) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;     // over rows * C
  if (e < (long long)rows * C) {
    const int bl = (int)(e / C), c = (int)(e - (long long)bl * C);
    if (c == C - 1) {
      dscores[e] = 0.f;                                     // the gold slot is sliced off (utils.py:36-37)
    } else if (y[(long long)(row0 + bl) * (C - 1) + c]) {
      dscores[e] -= k * (float)y[(long long)(row0 + bl) * (C - 1) + c] * (float)n[row0 + bl];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
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
  TripletScratch t = carve_triplet(scratch, B, nullptr);
  const float k = (float)(1.0 / ((double)B * (double)B * (double)(C - 1)));
  triplet_pos_kernel<<<(B + 255) / 256, 256, 0, stream>>>(scores, labels, B, C, t.p, t.n);
  DRIN_LAUNCH_CHECK();
  const long long local = (long long)rows * (C - 1);
  const int cblocks = (int)((local + TL_THREADS - 1) / TL_THREADS);
  if (cblocks > 65536) return fail(DRIN_ERR_ARG, "triplet_loss: too many local scores (%lld)", local);
  triplet_count_kernel<<<cblocks, TL_THREADS, 0, stream>>>(scores, t.p, B, C, row0, rows, margin, k, dscores, t.partial);
  DRIN_LAUNCH_CHECK();
  const long long total = (long long)B * (C - 1);
  dim3 ngrid((unsigned)((total + TL_TILE - 1) / TL_TILE), (unsigned)((rows + TL_THREADS - 1) / TL_THREADS));
  triplet_n_kernel<<<ngrid, TL_THREADS, 0, stream>>>(scores, t.p, B, C, row0, rows, margin, t.n);
  DRIN_LAUNCH_CHECK();
  const long long fe = (long long)rows * C;
  triplet_finish_kernel<<<(int)((fe + 255) / 256), 256, 0, stream>>>(labels, t.n, C, row0, rows, k, dscores, t.partial,
                                                                     cblocks, loss);
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
