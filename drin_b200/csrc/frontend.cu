// Front end of the DRIN hot path: one HBM-bound pass over the raw cached features of a batch.
//
// Replaces (reference file:line, upstream starreeze/drin):
//   * Avg.avg                    baselines/ghmfc.py:55-60   span mean of the mention tokens (used twice upstream)
//   * EntityEncoder pooling      baselines/ghmfc.py:237-249 WikiMEL masked mean over tokens 1..n-2
//   * region mean                drin/model.py:41           mean over the 49 ResNet regions
//   * EdgeEncoder                drin/model.py:60-94        tt = cos(span, entity CLS), ii = score-weighted object cosines
//   * edge list                  drin/model.py:201-204      [tt, mtei/100, miet/100, ii] (enable mask: GCN layers, model.py:122)
// and emits the GEMM A-operands of the four input projections as bf16 planes (split hi/lo in fp32 mode),
// so every raw feature byte is read from HBM exactly once.
//
// One CTA per mention (grid-stride), NW warps.  Phase A (all threads): span mean, region mean and the
// mention object crops go to shared memory.  Phase B (warp per candidate): entity text pooling / CLS
// cosine, object cosines, plane conversion of the entity text and image rows.  All loads are 16-byte,
// lane-contiguous (fully coalesced); rows are 3 KB (768 fp32) or 8 KB (2048 fp32).
#include "kernels.cuh"

namespace drin {

template <typename T>
struct Vec;   // 16-byte vector of T -> floats
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
};
template <>
struct Vec<bf16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const bf16* p, float (&f)[8]) {
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(p));
    const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// write N consecutive values as bf16 planes (hi, optional lo); N is 4 or 8
template <int N>
__device__ __forceinline__ void store_planes(bf16* hi, bf16* lo, long long off, const float (&f)[N]) {
  uint32_t h[N / 2], l[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    bf16 h0, l0, h1, l1;
    split_bf16(f[2 * i], h0, l0);
    split_bf16(f[2 * i + 1], h1, l1);
    h[i] = pack_bf16x2(h0, h1);
    l[i] = pack_bf16x2(l0, l1);
  }
  if (N == 4) {
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(h[0], h[1]);
    if (lo) *reinterpret_cast<uint2*>(lo + off) = make_uint2(l[0], l[1]);
  } else {
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

static constexpr int FE_MAX_OM = 4;

// lane-strided row tile: vector i of the lane covers elements (i*32 + lane)*VN .. +VN-1 (coalesced 512-B requests)
template <typename T, int NV>
__device__ __forceinline__ void load_row_tile(const T* __restrict__ row, int lane, float (&f)[NV][Vec<T>::N]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) Vec<T>::load(row + (i * 32 + lane) * Vec<T>::N, f[i]);
}

// One work item = (mention b, candidate slice).  Slices let WikiMEL-sized mentions (101 candidates x up to
// 196 KB of entity tokens) spread over several CTAs; slice 0 also owns the mention-side outputs.
template <typename T, int NW, int D, int R>
__global__ void __launch_bounds__(NW * 32, 3) frontend_kernel(const FrontendArgs a) {
  constexpr int VN = Vec<T>::N;
  constexpr int DV = D / (32 * VN);          // vectors per lane for a D row  (6 fp32 / 3 bf16)
  constexpr int RV = R / (32 * VN);          // vectors per lane for an R row (16 fp32 / 8 bf16)
  extern __shared__ __align__(16) float sm[];
  float* s_span = sm;                        // [D]
  float* s_mo = s_span + D;                  // [Om][R]
  float* s_red = s_mo + a.Om * R;            // [(1 + Om) * NW] scratch
  __shared__ float s_span_norm;
  __shared__ float s_mo_norm[FE_MAX_OM];
  __shared__ float s_ms[FE_MAX_OM];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* mtf = static_cast<const T*>(a.mtf);
  const T* mif = static_cast<const T*>(a.mif);
  const T* mof = static_cast<const T*>(a.mof);
  const T* etf = static_cast<const T*>(a.etf);
  const T* eif = static_cast<const T*>(a.eif);
  const T* eof = static_cast<const T*>(a.eof);
  const long long BC = (long long)a.B * a.C;
  const int S = a.slices, cps = (a.C + S - 1) / S;
  const long long items = (long long)a.B * S;

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / S), slice = (int)(item - (long long)b * S);
    const bool owner = slice == 0;
    // table row of this mention (drin/data.py:99-108 done on device when the inputs are resident tables)
    const long long m = a.mention_index ? a.mention_index[b] : (long long)b;
    // ------------------------------ phase A: mention side ------------------------------
    // span mean (ghmfc.py:55-60): rows start..end-1 with Python slice clamping
    long long s = a.start[m], e = a.end[m];
    if (s < 0) s += a.Lm;
    if (e < 0) e += a.Lm;
    s = s < 0 ? 0 : (s > a.Lm ? a.Lm : s);
    e = e < 0 ? 0 : (e > a.Lm ? a.Lm : e);
    const float cnt = e > s ? (float)(e - s) : 0.f;          // 0 -> 0/0 = NaN like mean of an empty slice
    float nrm_part = 0.f;
    for (int v = tid; v < D / VN; v += NW * 32) {
      float acc[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] = 0.f;
      const T* base = mtf + m * a.Lm * D + v * VN;
      long long r = s;
      for (; r + 4 <= e; r += 4) {
        float f[4][VN];
#pragma unroll
        for (int j = 0; j < 4; ++j) Vec<T>::load(base + (r + j) * D, f[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[i] += f[j][i];
      }
      for (; r < e; ++r) {
        float f[VN];
        Vec<T>::load(base + r * D, f);
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] += f[i];
      }
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        acc[i] = acc[i] / cnt;
        s_span[v * VN + i] = acc[i];
        nrm_part += acc[i] * acc[i];
      }
      if (owner && a.span_hi) store_planes<VN>(a.span_hi, a.span_lo, (long long)b * D + v * VN, acc);
      if (owner && a.span_f) {
#pragma unroll
        for (int i = 0; i < VN; ++i) a.span_f[(long long)b * D + v * VN + i] = acc[i];
      }
    }
    nrm_part = warp_sum(nrm_part);
    if (lane == 0) s_red[warp] = nrm_part;

    // mention object crops (model.py:78-79; the singleton dim is already folded) -> smem + norms
    for (int o = 0; o < a.Om; ++o) {
      float part = 0.f;
      for (int v = tid; v < R / VN; v += NW * 32) {
        float f[VN];
        Vec<T>::load(mof + (m * a.Om + o) * R + v * VN, f);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          s_mo[o * R + v * VN + i] = f[i];
          part += f[i] * f[i];
        }
      }
      part = warp_sum(part);
      if (lane == 0) s_red[NW * (1 + o) + warp] = part;
    }

    // region mean (model.py:41): P rows of R, 7 independent 16-B loads in flight per thread
    if (owner) {
      for (int v = tid; v < R / VN; v += NW * 32) {
        float acc[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = 0.f;
        const T* base = mif + m * a.P * R + v * VN;
        int r = 0;
        for (; r + 7 <= a.P; r += 7) {
          float f[7][VN];
#pragma unroll
          for (int j = 0; j < 7; ++j) Vec<T>::load(base + (long long)(r + j) * R, f[j]);
#pragma unroll
          for (int j = 0; j < 7; ++j)
#pragma unroll
            for (int i = 0; i < VN; ++i) acc[i] += f[j][i];
        }
        for (; r < a.P; ++r) {
          float f[VN];
          Vec<T>::load(base + (long long)r * R, f);
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[i] += f[i];
        }
        const float inv = 1.0f / (float)a.P;
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] *= inv;
        if (a.mim_hi) store_planes<VN>(a.mim_hi, a.mim_lo, (long long)b * R + v * VN, acc);
        if (a.mim_f) {
#pragma unroll
          for (int i = 0; i < VN; ++i) a.mim_f[(long long)b * R + v * VN + i] = acc[i];
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < NW; ++w) t += s_red[w];
      s_span_norm = fmaxf(sqrtf(t), 1e-8f);
    }
    if (tid >= 32 && tid < 32 + a.Om) {
      const int o = tid - 32;
      float t = 0.f;
      for (int w = 0; w < NW; ++w) t += s_red[NW * (1 + o) + w];
      s_mo_norm[o] = fmaxf(sqrtf(t), 1e-8f);
      s_ms[o] = a.mos[m * a.Om + o];
    }
    __syncthreads();

    // ------------------------------ phase B: one warp per candidate ------------------------------
    const int c_end = min(a.C, (slice + 1) * cps);
    for (int c = slice * cps + warp; c < c_end; c += NW) {
      const long long r = (long long)b * a.C + c;         // output row
      const long long rm = m * a.C + c;                   // row in the mention-major [N, C] tables (CLIP similarities)
      const long long re = a.entity_index ? a.entity_index[r] : rm;   // row in the entity-side tables
      // all rows of the candidate that do not depend on anything are requested up front
      float eo[RV][VN], ec[DV][VN];
      load_row_tile<T, RV>(eof + re * a.Oe * R, lane, eo);                      // first entity object crop
      load_row_tile<T, DV>(etf + (a.Le ? re * (long long)a.Le * D : re * D), lane, ec);   // CLS row / pooled row

      // --- entity text: tt = cos(span, CLS) and the pooled vertex feature (ghmfc.py:237-249, model.py:73-76)
      float dot = 0.f, nrm = 0.f;
#pragma unroll
      for (int v = 0; v < DV; ++v) {
        const float* sp = s_span + (v * 32 + lane) * VN;
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          dot += ec[v][i] * sp[i];
          nrm += ec[v][i] * ec[v][i];
        }
      }
      if (a.Le == 0) {
#pragma unroll
        for (int v = 0; v < DV; ++v) {
          const long long off = r * D + (v * 32 + lane) * VN;
          if (a.ep_hi) store_planes<VN>(a.ep_hi, a.ep_lo, off, ec[v]);
          if (a.ep_f) {
#pragma unroll
            for (int i = 0; i < VN; ++i) a.ep_f[off + i] = ec[v][i];
          }
        }
      } else {
        long long n = 0;
        for (int t = lane; t < a.Le; t += 32) n += a.emask[re * a.Le + t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        long long t0 = 1, t1 = n - 1;                       // tokens 1 .. n-2 (drops CLS and SEP)
        if (t1 < 0) t1 += a.Le;                             // Python slice semantics for a negative stop
        if (t1 > a.Le) t1 = a.Le;
        const float cnt_e = t1 > t0 ? (float)(t1 - t0) : 0.f;
        const T* rowbase = etf + re * (long long)a.Le * D;
        float acc[DV][VN];
#pragma unroll
        for (int v = 0; v < DV; ++v)
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[v][i] = 0.f;
        long long t = t0;
        for (; t + 4 <= t1; t += 4) {                       // 4 rows x DV vectors in flight per lane
          float g[4][DV][VN];
#pragma unroll
          for (int j = 0; j < 4; ++j) load_row_tile<T, DV>(rowbase + (t + j) * D, lane, g[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < DV; ++v)
#pragma unroll
              for (int i = 0; i < VN; ++i) acc[v][i] += g[j][v][i];
        }
        for (; t < t1; ++t) {
          float g[DV][VN];
          load_row_tile<T, DV>(rowbase + t * D, lane, g);
#pragma unroll
          for (int v = 0; v < DV; ++v)
#pragma unroll
            for (int i = 0; i < VN; ++i) acc[v][i] += g[v][i];
        }
#pragma unroll
        for (int v = 0; v < DV; ++v) {
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[v][i] = acc[v][i] / cnt_e;
          const long long off = r * D + (v * 32 + lane) * VN;
          if (a.ep_hi) store_planes<VN>(a.ep_hi, a.ep_lo, off, acc[v]);
          if (a.ep_f) {
#pragma unroll
            for (int i = 0; i < VN; ++i) a.ep_f[off + i] = acc[v][i];
          }
        }
      }
      dot = warp_sum(dot);
      nrm = warp_sum(nrm);
      const float tt = dot / (s_span_norm * fmaxf(sqrtf(nrm), 1e-8f));

      // --- object crops: ii (model.py:84-92)
      float sim = 0.f, den = 0.f;
      for (int j = 0; j < a.Oe; ++j) {
        if (j > 0) load_row_tile<T, RV>(eof + (re * a.Oe + j) * R, lane, eo);
        float d[FE_MAX_OM] = {0.f, 0.f, 0.f, 0.f};
        float en2 = 0.f;
#pragma unroll
        for (int v = 0; v < RV; ++v) {
#pragma unroll
          for (int i = 0; i < VN; ++i) en2 += eo[v][i] * eo[v][i];
#pragma unroll
          for (int o = 0; o < FE_MAX_OM; ++o)
            if (o < a.Om) {
              const float* mo = s_mo + o * R + (v * 32 + lane) * VN;
#pragma unroll
              for (int i = 0; i < VN; ++i) d[o] += eo[v][i] * mo[i];
            }
        }
        en2 = warp_sum(en2);
        const float enorm = fmaxf(sqrtf(en2), 1e-8f);
        const float es = a.eos[re * a.Oe + j];
        for (int o = 0; o < a.Om; ++o) {                    // upstream loop order: i (mention) outer, j inner;
          const float cs = warp_sum(d[o]) / (s_mo_norm[o] * enorm);   // with Oe == 1 the orders coincide
          const float w = s_ms[o] * es;
          sim += cs * w;
          den += w;
        }
      }
      const float ii = sim / (den + 1e-9f);

      // --- entity image row -> planes (A operand of the entity-image projection, model.py:45)
      if (a.ei_hi) {
        load_row_tile<T, RV>(eif + re * R, lane, eo);
#pragma unroll
        for (int v = 0; v < RV; ++v) store_planes<VN>(a.ei_hi, a.ei_lo, r * R + (v * 32 + lane) * VN, eo[v]);
      }
      if (lane == 0 && a.edges) {                            // model.py:201-204 order tt, ti, it, ii
        a.edges[r] = tt;
        a.edges[BC + r] = a.mtei[rm] / 100.f;
        a.edges[2 * BC + r] = a.miet[rm] / 100.f;
        a.edges[3 * BC + r] = ii;
      }
    }
    __syncthreads();   // smem is reused by the next work item
  }
}

template <typename T>
static int launch_frontend(cudaStream_t stream, FrontendArgs a) {
  constexpr int D = 768, R = 2048;
  const size_t smem = (size_t)(D + a.Om * R + 8 * (1 + FE_MAX_OM)) * sizeof(float);
  // candidate slices: enough work items to fill the machine (3 CTAs/SM) when the batch alone cannot
  const int target = 148 * 3;
  int slices = 1;
  if (a.C >= 32 && a.B < 2 * target) {
    slices = (2 * target + a.B - 1) / a.B;
    const int max_slices = (a.C + 7) / 8;
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
  }
  a.slices = slices;
  const long long items = (long long)a.B * slices;
  const int grid = (int)(items < 148 * 6 ? items : 148 * 6);
  frontend_kernel<T, 4, D, R><<<grid, 128, smem, stream>>>(a);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

int frontend(cudaStream_t stream, const FrontendArgs& a, bool bf16_features) {
  prof::Scope prof_scope(stream, prof::FRONTEND);
  if (a.D != 768 || a.R != 2048)
    return fail(DRIN_ERR_ARG, "frontend: built for bert_embed_dim 768 / resnet_embed_dim 2048 (got %d / %d)", a.D, a.R);
  if (a.Om > FE_MAX_OM) return fail(DRIN_ERR_ARG, "frontend: at most %d mention objects", FE_MAX_OM);
  if (a.B <= 0 || a.C <= 0) return fail(DRIN_ERR_ARG, "frontend: empty batch");
  static bool attr_set_dev[64] = {};
  int dev = 0;
  DRIN_CUDA(cudaGetDevice(&dev));
  bool& attr_set = attr_set_dev[dev & 63];         // function attributes are per device
  if (!attr_set) {
    const int max_smem = 100 * 1024;
    DRIN_CUDA(cudaFuncSetAttribute(frontend_kernel<float, 4, 768, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    DRIN_CUDA(cudaFuncSetAttribute(frontend_kernel<bf16, 4, 768, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set = true;
  }
  return bf16_features ? launch_frontend<bf16>(stream, a) : launch_frontend<float>(stream, a);
}

}  // namespace drin
