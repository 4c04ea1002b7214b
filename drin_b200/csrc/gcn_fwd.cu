// Fused memory-bound forward stages of the GCN (reference drin/model.py:121-153, 207-209).
//
// The dense work of a layer is three GEMMs (gemm_tcgen05.cu); everything around them is fused here so
// every vertex row (768 fp32 = 3 KB) is read once per layer:
//
//   gcn_layer_fwd   per mention: enable mask (model.py:122), message aggregation for all four vertex
//                   types (model.py:124-127,139-146), the dynamic edge update
//                   e' = sigmoid(mean_D(W_u u * W_v v) + e) (model.py:131-134,148-153) and the split-bf16
//                   A operand z = agg + x of the shared W_h GEMM (model.py:128).
//                   LayerNorm + GELU of the PREVIOUS layer (model.py:128) is applied on the fly when the
//                   candidate rows are read, so activated vertices are never written to HBM.
//   mention_ln      LayerNorm + GELU of the 2B mention rows (they feed the small W_u GEMM).
//   score           final LayerNorm + GELU + cosine candidate scoring (model.py:207-209).
//
// Edge update without the big W_v GEMM: mean_D(fu_b * (W_v v + b_v)) = (v . (fu_b W_v) + fu_b . b_v) / D,
// so only g_b = fu_b W_v (2B rows) goes through the tensor cores and each candidate costs 4 dot products.
//
// Layout: a warp owns one vertex row; lane l holds elements (j*32 + l)*4 .. +3, j < D/128, i.e. every
// load/store is a fully coalesced 512-B warp transaction.  Row statistics use warp shuffles.
#include "kernels.cuh"
#include "pipe.cuh"
#include "rows.cuh"

namespace drin {

// ---------------------------------------------------------------------------------------------
// gcn_layer_fwd
// ---------------------------------------------------------------------------------------------
// Shared-memory staged version: a producer warp streams each mention's candidate rows (and its four
// mention-side vectors) into a 2-stage ring with 1-D bulk async copies (TMA engine, mbarrier completion);
// NW consumer warps process rows out of shared memory.  Bytes in flight are set by the ring (up to ~78 KB
// per SM), not by registers or occupancy, and the consumers never wait on HBM latency.
static constexpr int ST_CH = 11;       // candidate rows per array and chunk (one WikiDiverse mention)
static constexpr int ST_STAGES = 2;

template <int D, int NW, bool FULL>
__global__ void __launch_bounds__((NW + 1) * 32, 1) gcn_layer_fwd_kernel(const LayerFwdArgs a) {
  extern __shared__ __align__(128) float sm[];
  constexpr int NVEC = FULL ? 4 : 2;                          // mt, mi (, g_mt, g_mi)
  constexpr int STAGE_FLOATS = (2 * ST_CH + NVEC) * D;
  float* s_gamma = sm + ST_STAGES * STAGE_FLOATS;             // LayerNorm of the previous layer (if a.ln_gamma)
  float* s_beta = s_gamma + D;
  float* s_acc = s_beta + D;                                  // [2][NW][D] per-warp partial messages
  __shared__ __align__(8) unsigned long long bars[2 * ST_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long B = a.B, BC = (long long)a.B * a.C;
  const bool ln = a.ln_gamma != nullptr;
  const bool dyn = FULL && a.g != nullptr;                    // dynamic edge update (false: static edges)
  const int nchunks = (a.C + ST_CH - 1) / ST_CH;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[ST_STAGES + s]); };
  if (tid == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (ln) {
    for (int i = tid; i < D; i += (NW + 1) * 32) {
      s_gamma[i] = a.ln_gamma[i];
      s_beta[i] = a.ln_beta[i];
    }
  }
  __syncthreads();

  if (warp == NW) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        for (int k = 0; k < nchunks; ++k) {
          const int n = min(ST_CH, a.C - k * ST_CH);
          const long long r0 = (long long)b * a.C + k * ST_CH;
          mbar_wait(empty_bar(stage), phase ^ 1u, nullptr, 0);
          const uint32_t row_bytes = (uint32_t)(n * D * sizeof(float));
          mbar_arrive_expect_tx(full_bar(stage), 2 * row_bytes + (dyn ? 4u : 2u) * D * (uint32_t)sizeof(float));
          const uint32_t base = smem_u32(sm + stage * STAGE_FLOATS);
          bulk_copy_g2s(base, a.x_et + r0 * D, row_bytes, full_bar(stage));
          bulk_copy_g2s(base + ST_CH * D * 4, a.x_ei + r0 * D, row_bytes, full_bar(stage));
          const uint32_t vb = base + 2 * ST_CH * D * 4;
          bulk_copy_g2s(vb, a.xm + (long long)b * D, D * 4, full_bar(stage));
          bulk_copy_g2s(vb + D * 4, a.xm + (B + b) * D, D * 4, full_bar(stage));
          if (dyn) {
            bulk_copy_g2s(vb + 2 * D * 4, a.g + (long long)b * D, D * 4, full_bar(stage));
            bulk_copy_g2s(vb + 3 * D * 4, a.g + (B + b) * D, D * 4, full_bar(stage));
          }
          if (++stage == ST_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  const float invC = 1.0f / (float)a.C, invD = 1.0f / (float)D;
  float* acc_mt = s_acc + warp * D;
  float* acc_mi = s_acc + (NW + warp) * D;
  int stage = 0;
  uint32_t phase = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
#pragma unroll
    for (int j = 0; j < RowT<D>::NV; ++j) {
      *reinterpret_cast<float4*>(acc_mt + (j * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(acc_mi + (j * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float beta_mt = dyn ? a.beta_u[b] : 0.f;
    const float beta_mi = dyn ? a.beta_u[B + b] : 0.f;
    for (int k = 0; k < nchunks; ++k) {
      const int n = min(ST_CH, a.C - k * ST_CH);
      mbar_wait(full_bar(stage), phase, nullptr, 0);
      const float* st = sm + stage * STAGE_FLOATS;
      const float* rows_et = st;
      const float* rows_ei = st + ST_CH * D;
      const float* s_mt = st + 2 * ST_CH * D;
      const float* s_mi = s_mt + D;
      const float* s_gmt = s_mi + D;
      const float* s_gmi = s_gmt + D;
      for (int c = warp; c < n; c += NW) {
        const long long r = (long long)b * a.C + k * ST_CH + c;
        // enable mask (model.py:122); edge order tt (mt-et), ti (mt-ei), it (mi-et), ii (mi-ei)
        const float e0 = a.edges_in[r] * a.en[0], e1 = a.edges_in[BC + r] * a.en[1];
        const float e2 = a.edges_in[2 * BC + r] * a.en[2], e3 = a.edges_in[3 * BC + r] * a.en[3];
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        const long long zr_et = (FULL ? 2 * B : B) + r, zr_ei = 2 * B + BC + r;
        // both vertex rows of the candidate are processed together (independent instruction streams -> ILP);
        // messages: mt <- e0 et + e1 ei, mi <- e2 et + e3 ei ; z_et = et + e0 mt + e2 mi ; z_ei = ei + e1 mt + e3 mi
        // (model.py:124-128,139-146)
        RowT<D> xet, xei;
        row_load<D>(xet, rows_et + c * D, lane);
        row_load<D>(xei, rows_ei + c * D, lane);
        if (ln) {
          row_ln_gelu<D>(xet, s_gamma, s_beta, lane);
          row_ln_gelu<D>(xei, s_gamma, s_beta, lane);
        }
        if (dyn) {
          d0 = row_dot<D>(xet, s_gmt, lane); d1 = row_dot<D>(xei, s_gmt, lane);
          d2 = row_dot<D>(xet, s_gmi, lane); d3 = row_dot<D>(xei, s_gmi, lane);
        }
#pragma unroll
        for (int j = 0; j < RowT<D>::NV; ++j) {
          const int off = (j * 32 + lane) * 4;
          float4 am = *reinterpret_cast<float4*>(acc_mt + off);
          am.x += e0 * xet.v[4 * j] + e1 * xei.v[4 * j];
          am.y += e0 * xet.v[4 * j + 1] + e1 * xei.v[4 * j + 1];
          am.z += e0 * xet.v[4 * j + 2] + e1 * xei.v[4 * j + 2];
          am.w += e0 * xet.v[4 * j + 3] + e1 * xei.v[4 * j + 3];
          *reinterpret_cast<float4*>(acc_mt + off) = am;
          if (FULL) {
            float4 ai = *reinterpret_cast<float4*>(acc_mi + off);
            ai.x += e2 * xet.v[4 * j] + e3 * xei.v[4 * j];
            ai.y += e2 * xet.v[4 * j + 1] + e3 * xei.v[4 * j + 1];
            ai.z += e2 * xet.v[4 * j + 2] + e3 * xei.v[4 * j + 2];
            ai.w += e2 * xet.v[4 * j + 3] + e3 * xei.v[4 * j + 3];
            *reinterpret_cast<float4*>(acc_mi + off) = ai;
          }
          const float4 mt = *reinterpret_cast<const float4*>(s_mt + off);
          const float4 mi = *reinterpret_cast<const float4*>(s_mi + off);
          xet.v[4 * j] += e0 * mt.x + e2 * mi.x;
          xet.v[4 * j + 1] += e0 * mt.y + e2 * mi.y;
          xet.v[4 * j + 2] += e0 * mt.z + e2 * mi.z;
          xet.v[4 * j + 3] += e0 * mt.w + e2 * mi.w;
          if (FULL) {
            xei.v[4 * j] += e1 * mt.x + e3 * mi.x;
            xei.v[4 * j + 1] += e1 * mt.y + e3 * mi.y;
            xei.v[4 * j + 2] += e1 * mt.z + e3 * mi.z;
            xei.v[4 * j + 3] += e1 * mt.w + e3 * mi.w;
          }
        }
        row_store_planes<D>(xet, a.z_hi + zr_et * D, a.z_lo ? a.z_lo + zr_et * D : nullptr, lane);
        if (FULL) row_store_planes<D>(xei, a.z_hi + zr_ei * D, a.z_lo ? a.z_lo + zr_ei * D : nullptr, lane);
        if (dyn) {
          // dynamic edge update (model.py:131-134,148-153): e' = sigmoid((v . g_u + fu . b_v) / D + e)
          d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3);
          if (lane == 0) {
            a.edges_out[r] = 1.0f / (1.0f + __expf(-((d0 + beta_mt) * invD + e0)));
            a.edges_out[BC + r] = 1.0f / (1.0f + __expf(-((d1 + beta_mt) * invD + e1)));
            a.edges_out[2 * BC + r] = 1.0f / (1.0f + __expf(-((d2 + beta_mi) * invD + e2)));
            a.edges_out[3 * BC + r] = 1.0f / (1.0f + __expf(-((d3 + beta_mi) * invD + e3)));
          }
        }
      }
      if (k == nchunks - 1) {
        // cross-warp reduction of the mention messages (fixed order -> deterministic); mean over ALL C slots
        named_bar_sync(1, NW * 32);
        for (int i = tid; i < (FULL ? 2 : 1) * D; i += NW * 32) {
          const int which = i / D, col = i - which * D;
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < NW; ++w) t += s_acc[(which * NW + w) * D + col];
          const float zval = (which ? s_mi[col] : s_mt[col]) + t * invC;
          bf16 h, l;
          split_bf16(zval, h, l);
          const long long zr = which ? B + b : b;
          a.z_hi[zr * D + col] = h;
          if (a.z_lo) a.z_lo[zr * D + col] = l;
        }
        named_bar_sync(1, NW * 32);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(stage));
      if (++stage == ST_STAGES) { stage = 0; phase ^= 1u; }
    }
  }
}

// Warp-autonomous version (default when there is at least one mention per warp).  A warp owns a mention: its four
// mention-side vectors sit in a private shared-memory slice, the messages to the two mention vertices are summed over
// the candidates IN REGISTERS (candidate order -> deterministic), and the next candidate's two rows are prefetched
// into registers while the current candidate is processed.  No CTA barrier and no cross-warp reduction in the loop.
template <int D, int NW, bool FULL>
__global__ void __launch_bounds__(NW * 32, 1) gcn_layer_fwd_warp_kernel(const LayerFwdArgs a) {
  constexpr int NV = RowT<D>::NV, NE = NV * 4;
  extern __shared__ __align__(16) float sm[];
  float* s_gamma = sm;
  float* s_beta = s_gamma + D;
  float* s_vec = s_beta + D;                                  // [NW][4][D]: mt, mi, g_mt, g_mi of the warp's mention
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long B = a.B, BC = (long long)a.B * a.C;
  const bool ln = a.ln_gamma != nullptr;
  const bool dyn = FULL && a.g != nullptr;
  if (ln) {
    for (int i = tid; i < D; i += NW * 32) {
      s_gamma[i] = a.ln_gamma[i];
      s_beta[i] = a.ln_beta[i];
    }
  }
  __syncthreads();
  float* v_mt = s_vec + warp * 4 * D;
  float* v_mi = v_mt + D;
  float* v_gmt = v_mi + D;
  float* v_gmi = v_gmt + D;
  const float invC = 1.0f / (float)a.C, invD = 1.0f / (float)D;
  const float en0 = a.en[0], en1 = a.en[1], en2 = a.en[2], en3 = a.en[3];
  const long long gwarp = (long long)blockIdx.x * NW + warp, nwarps = (long long)gridDim.x * NW;
  const int S = a.slices, cps = (a.C + S - 1) / S;            // work unit = (mention, candidate slice)
  const long long units = B * S;

  for (long long u = gwarp; u < units; u += nwarps) {
    const long long b = u / S;
    const int c0 = (int)(u - b * S) * cps, c1 = min(a.C, c0 + cps);
    if (c0 >= a.C) {                                          // empty trailing slice (C not a multiple of the slice size)
      float* part = a.acc_part + u * 2 * D;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        *reinterpret_cast<float4*>(part + (j * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(part + D + (j * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    RowT<D> pet, pei;                                         // prefetched rows of the next candidate
    row_load<D>(pet, a.x_et + (b * a.C + c0) * D, lane);
    row_load<D>(pei, a.x_ei + (b * a.C + c0) * D, lane);
    {
      RowT<D> t;
      row_load<D>(t, a.xm + b * D, lane);
      row_store<D>(t, v_mt, lane);
      row_load<D>(t, a.xm + (B + b) * D, lane);
      row_store<D>(t, v_mi, lane);
      if (dyn) {
        row_load<D>(t, a.g + b * D, lane);
        row_store<D>(t, v_gmt, lane);
        row_load<D>(t, a.g + (B + b) * D, lane);
        row_store<D>(t, v_gmi, lane);
      }
    }
    const float beta_mt = dyn ? a.beta_u[b] : 0.f, beta_mi = dyn ? a.beta_u[B + b] : 0.f;
    __syncwarp();
    float acc_mt[NE], acc_mi[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) acc_mt[i] = acc_mi[i] = 0.f;
    long long r = b * a.C + c0;
    float n0 = a.edges_in[r], n1 = a.edges_in[BC + r], n2 = a.edges_in[2 * BC + r], n3 = a.edges_in[3 * BC + r];
    for (int c = c0; c < c1; ++c, ++r) {
      RowT<D> xet = pet, xei = pei;
      // enable mask (model.py:122); edge order tt (mt-et), ti (mt-ei), it (mi-et), ii (mi-ei)
      const float e0 = n0 * en0, e1 = n1 * en1, e2 = n2 * en2, e3 = n3 * en3;
      if (c + 1 < c1) {
        row_load<D>(pet, a.x_et + (r + 1) * D, lane);
        row_load<D>(pei, a.x_ei + (r + 1) * D, lane);
        n0 = a.edges_in[r + 1]; n1 = a.edges_in[BC + r + 1]; n2 = a.edges_in[2 * BC + r + 1]; n3 = a.edges_in[3 * BC + r + 1];
      }
      if (lane == 0) {        // DRAM latency runs two candidates ahead of the register loads (bulk L2 prefetch)
        long long pr = -1;
        if (c + 2 < c1) {
          pr = r + 2;
        } else if (u + nwarps < units) {                      // first rows of this warp's next unit
          const long long nb = (u + nwarps) / S;
          const int nc0 = (int)(u + nwarps - nb * S) * cps;
          pr = nb * a.C + min(nc0 + (c + 2 - c1), a.C - 1);
        }
        if (pr >= 0) {
          l2_prefetch(a.x_et + pr * D, D * 4);
          l2_prefetch(a.x_ei + pr * D, D * 4);
        }
      }
      if (ln) {
        row_ln_gelu<D>(xet, s_gamma, s_beta, lane);
        row_ln_gelu<D>(xei, s_gamma, s_beta, lane);
      }
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      if (dyn) {
        d0 = row_dot<D>(xet, v_gmt, lane); d1 = row_dot<D>(xei, v_gmt, lane);
        d2 = row_dot<D>(xet, v_gmi, lane); d3 = row_dot<D>(xei, v_gmi, lane);
      }
      // messages (model.py:124-128,139-146): mt <- e0 et + e1 ei, mi <- e2 et + e3 ei ;
      // z_et = et + e0 mt + e2 mi ; z_ei = ei + e1 mt + e3 mi
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 mt4 = *reinterpret_cast<const float4*>(v_mt + (j * 32 + lane) * 4);
        const float4 mi4 = *reinterpret_cast<const float4*>(v_mi + (j * 32 + lane) * 4);
        const float mt[4] = {mt4.x, mt4.y, mt4.z, mt4.w}, mi[4] = {mi4.x, mi4.y, mi4.z, mi4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 4 * j + k;
          acc_mt[i] += e0 * xet.v[i] + e1 * xei.v[i];
          if (FULL) acc_mi[i] += e2 * xet.v[i] + e3 * xei.v[i];
          xet.v[i] += e0 * mt[k] + e2 * mi[k];
          if (FULL) xei.v[i] += e1 * mt[k] + e3 * mi[k];
        }
      }
      const long long zr_et = (FULL ? 2 * B : B) + r, zr_ei = 2 * B + BC + r;
      row_store_planes<D>(xet, a.z_hi + zr_et * D, a.z_lo ? a.z_lo + zr_et * D : nullptr, lane);
      if (FULL) row_store_planes<D>(xei, a.z_hi + zr_ei * D, a.z_lo ? a.z_lo + zr_ei * D : nullptr, lane);
      if (dyn) {
        // dynamic edge update (model.py:131-134,148-153): e' = sigmoid((v . g_u + fu . b_v) / D + e)
        d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3);
        if (lane == 0) {
          a.edges_out[r] = 1.0f / (1.0f + __expf(-((d0 + beta_mt) * invD + e0)));
          a.edges_out[BC + r] = 1.0f / (1.0f + __expf(-((d1 + beta_mt) * invD + e1)));
          a.edges_out[2 * BC + r] = 1.0f / (1.0f + __expf(-((d2 + beta_mi) * invD + e2)));
          a.edges_out[3 * BC + r] = 1.0f / (1.0f + __expf(-((d3 + beta_mi) * invD + e3)));
        }
      }
    }
    if (S > 1) {
      // sliced mention: leave the partial messages to layer_fwd_mention_finish (fixed slice order -> deterministic)
      float* part = a.acc_part + u * 2 * D;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int off = (j * 32 + lane) * 4;
        *reinterpret_cast<float4*>(part + off) = make_float4(acc_mt[4 * j], acc_mt[4 * j + 1], acc_mt[4 * j + 2], acc_mt[4 * j + 3]);
        if (FULL)
          *reinterpret_cast<float4*>(part + D + off) = make_float4(acc_mi[4 * j], acc_mi[4 * j + 1], acc_mi[4 * j + 2], acc_mi[4 * j + 3]);
      }
      __syncwarp();
      continue;
    }
    // mention rows: z_m = x_m + mean over ALL C slots of the messages
    RowT<D> zm;
    row_load<D>(zm, v_mt, lane);
#pragma unroll
    for (int i = 0; i < NE; ++i) zm.v[i] = zm.v[i] + acc_mt[i] * invC;
    row_store_planes<D>(zm, a.z_hi + b * D, a.z_lo ? a.z_lo + b * D : nullptr, lane);
    if (FULL) {
      row_load<D>(zm, v_mi, lane);
#pragma unroll
      for (int i = 0; i < NE; ++i) zm.v[i] = zm.v[i] + acc_mi[i] * invC;
      row_store_planes<D>(zm, a.z_hi + (B + b) * D, a.z_lo ? a.z_lo + (B + b) * D : nullptr, lane);
    }
    __syncwarp();                                             // the vector slice is rewritten for the next mention
  }
}

// sliced mentions: z_m = x_m + (sum over slices of the partial messages) / C, one warp per mention row
template <int D>
__global__ void __launch_bounds__(256) layer_fwd_mention_finish_kernel(const LayerFwdArgs a) {
  constexpr int NE = RowT<D>::NV * 4;
  const int lane = threadIdx.x & 31;
  const long long B = a.B;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows = a.full ? 2 * B : B;                  // mt rows, then mi rows
  const float invC = 1.0f / (float)a.C;
  for (long long r = warp0; r < rows; r += nwarps) {
    const long long b = r < B ? r : r - B;
    const int which = r < B ? 0 : 1;
    RowT<D> z, t;
    row_load<D>(z, a.xm + r * D, lane);
    RowT<D> acc;
#pragma unroll
    for (int i = 0; i < NE; ++i) acc.v[i] = 0.f;
    for (int sl = 0; sl < a.slices; ++sl) {
      row_load<D>(t, a.acc_part + ((b * a.slices + sl) * 2 + which) * D, lane);
#pragma unroll
      for (int i = 0; i < NE; ++i) acc.v[i] += t.v[i];
    }
#pragma unroll
    for (int i = 0; i < NE; ++i) z.v[i] = z.v[i] + acc.v[i] * invC;
    row_store_planes<D>(z, a.z_hi + r * D, a.z_lo ? a.z_lo + r * D : nullptr, lane);
  }
}

template <int D, int NW>
static int launch_layer_fwd(cudaStream_t stream, const LayerFwdArgs& a) {
  const int nvec = a.full ? 4 : 2;
  const size_t smem = (size_t)(ST_STAGES * (2 * ST_CH + nvec) * D + 2 * D + 2 * NW * D) * sizeof(float);
  const int grid = a.B < 148 ? a.B : 148;
  if (a.full) {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_fwd_kernel<D, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    gcn_layer_fwd_kernel<D, NW, true><<<grid, (NW + 1) * 32, smem, stream>>>(a);
  } else {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_fwd_kernel<D, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    gcn_layer_fwd_kernel<D, NW, false><<<grid, (NW + 1) * 32, smem, stream>>>(a);
  }
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

template <int D, int NW>
static int launch_layer_fwd_warp(cudaStream_t stream, const LayerFwdArgs& a) {
  const size_t smem = (size_t)(2 * D + NW * 4 * D) * sizeof(float);
  if (a.full) {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_fwd_warp_kernel<D, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    gcn_layer_fwd_warp_kernel<D, NW, true><<<148, NW * 32, smem, stream>>>(a);
  } else {
    DRIN_CUDA(cudaFuncSetAttribute(gcn_layer_fwd_warp_kernel<D, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    gcn_layer_fwd_warp_kernel<D, NW, false><<<148, NW * 32, smem, stream>>>(a);
  }
  DRIN_LAUNCH_CHECK();
  if (a.slices > 1) {
    const long long rows = a.full ? 2LL * a.B : a.B;
    const int grid = (int)((rows + 7) / 8 < 148 * 4 ? (rows + 7) / 8 : 148 * 4);
    layer_fwd_mention_finish_kernel<D><<<grid, 256, 0, stream>>>(a);
    DRIN_LAUNCH_CHECK();
  }
  return DRIN_OK;
}

static int g_row_slice_min = 8;          // fewest candidates a slice may hold (test / A-B hook)
void debug_set_row_slice_min(int v) { g_row_slice_min = v > 0 ? v : 8; }

int row_kernel_slices(int B, int C) {
  const int max_slices = C / g_row_slice_min;         // at least 8 candidates per slice
  if (max_slices < 2) return 1;
  const long long warps = 148LL * 8;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_slices; ++s) {
    const long long units = (long long)B * s;
    const long long rounds = (units + warps - 1) / warps;
    const double eff = (double)units / (double)(rounds * warps);
    if (eff > best_eff + 1e-9) { best = s; best_eff = eff; }     // ties: fewer slices (less per-slice overhead)
  }
  return best;
}

static int g_layer_fwd_variant = -1;     // -1 auto, 0 staged CTA-per-mention, >= 1 warp-per-mention (test hook)
void debug_set_layer_fwd_variant(int v) { g_layer_fwd_variant = v; }

int gcn_layer_fwd(cudaStream_t stream, const LayerFwdArgs& a) {
  prof::Scope prof_scope(stream, prof::GCN_FWD);
  if (a.D != 768) return fail(DRIN_ERR_ARG, "gcn_layer_fwd: gcn_embed_dim %d not built (768 only)", a.D);
  const int slices = a.slices > 0 ? a.slices : 1;
  const int variant = g_layer_fwd_variant < 0 ? ((long long)a.B * slices >= 148 * 8 ? 1 : 0) : g_layer_fwd_variant;
  if (variant >= 1) {                                        // 12 warps / SM measured slower (spills, tail)
    if (slices > 1 && !a.acc_part) return fail(DRIN_ERR_ARG, "gcn_layer_fwd: sliced launch without a partial buffer");
    LayerFwdArgs b = a;
    b.slices = slices;
    return launch_layer_fwd_warp<768, 8>(stream, b);
  }
  return launch_layer_fwd<768, 8>(stream, a);
}

// ---------------------------------------------------------------------------------------------
// mention_ln: x = gelu(LN(h)) for `rows` rows (warp per row) -> fp32 and optional planes
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) mention_ln_kernel(const float* __restrict__ h, long long rows,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ x,
                                                          bf16* __restrict__ x_hi, bf16* __restrict__ x_lo) {
  __shared__ __align__(16) float s_gamma[D];
  __shared__ __align__(16) float s_beta[D];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    s_gamma[i] = gamma[i];
    s_beta[i] = beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    RowT<D> row;
    row_load<D>(row, h + r * D, lane);
    row_ln_gelu<D>(row, s_gamma, s_beta, lane);
    if (x) row_store<D>(row, x + r * D, lane);
    if (x_hi) row_store_planes<D>(row, x_hi + r * D, x_lo ? x_lo + r * D : nullptr, lane);
  }
}

int mention_ln(cudaStream_t stream, int D, const float* h, long long rows, const float* gamma, const float* beta,
               float* x, bf16* x_hi, bf16* x_lo) {
  prof::Scope prof_scope(stream, prof::GCN_FWD);
  if (D != 768) return fail(DRIN_ERR_ARG, "mention_ln: gcn_embed_dim %d not built (768 only)", D);
  const long long blocks = (rows + 7) / 8;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  mention_ln_kernel<768><<<grid, 256, 0, stream>>>(h, rows, gamma, beta, x, x_hi, x_lo);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// rowdot: out[r] = x[r] . w   (beta_u = fu . b_v)
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ x, long long rows,
                                                     const float* __restrict__ w, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    RowT<D> row;
    row_load<D>(row, x + r * D, lane);
    const float d = warp_sum(row_dot<D>(row, w, lane));
    if (lane == 0) out[r] = d;
  }
}

int rowdot(cudaStream_t stream, int D, const float* x, long long rows, const float* w, float* out) {
  prof::Scope prof_scope(stream, prof::GCN_FWD);
  if (D != 768) return fail(DRIN_ERR_ARG, "rowdot: gcn_embed_dim %d not built (768 only)", D);
  const long long blocks = (rows + 7) / 8;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  rowdot_kernel<768><<<grid, 256, 0, stream>>>(x, rows, w, out);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// score: scores[b, c] = cos(gelu(LN(h_mt[b])), gelu(LN(h_et[b, c])))   (model.py:207-209)
// ---------------------------------------------------------------------------------------------
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32) score_kernel(const float* __restrict__ h_mt, const float* __restrict__ h_et,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int B, int C,
                                                         float* __restrict__ scores) {
  __shared__ __align__(16) float s_gamma[D];
  __shared__ __align__(16) float s_beta[D];
  __shared__ __align__(16) float s_m[D];
  __shared__ float s_mnorm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += NW * 32) {
    s_gamma[i] = gamma[i];
    s_beta[i] = beta[i];
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    if (warp == 0) {
      RowT<D> m;
      row_load<D>(m, h_mt + (long long)b * D, lane);
      row_ln_gelu<D>(m, s_gamma, s_beta, lane);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < RowT<D>::NV * 4; ++i) q += m.v[i] * m.v[i];
      q = warp_sum(q);
      row_store<D>(m, s_m, lane);
      if (lane == 0) s_mnorm = fmaxf(sqrtf(q), 1e-8f);
    }
    __syncthreads();
    for (int c = warp; c < C; c += NW) {
      const long long r = (long long)b * C + c;
      RowT<D> e;
      row_load<D>(e, h_et + r * D, lane);
      row_ln_gelu<D>(e, s_gamma, s_beta, lane);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < RowT<D>::NV * 4; ++i) q += e.v[i] * e.v[i];
      q = warp_sum(q);
      const float d = warp_sum(row_dot<D>(e, s_m, lane));
      if (lane == 0) scores[r] = d / (s_mnorm * fmaxf(sqrtf(q), 1e-8f));
    }
  }
}

// Warp-autonomous version (default when there is at least one mention per warp): a warp keeps the activated mention
// row in registers and walks its C candidate rows, prefetching the next row while it normalises the current one --
// no CTA barrier, one coalesced 3 KB row read per candidate.
template <int D, int NW>
__global__ void __launch_bounds__(NW * 32, 2) score_warp_kernel(const float* __restrict__ h_mt, const float* __restrict__ h_et,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int B, int C, int S,
                                                              float* __restrict__ scores) {
  constexpr int NE = RowT<D>::NV * 4;
  __shared__ __align__(16) float s_gamma[D];
  __shared__ __align__(16) float s_beta[D];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += NW * 32) {
    s_gamma[i] = gamma[i];
    s_beta[i] = beta[i];
  }
  __syncthreads();
  const long long gwarp = (long long)blockIdx.x * NW + warp, nwarps = (long long)gridDim.x * NW;
  const int cps = (C + S - 1) / S;                            // work unit = (mention, candidate slice)
  const long long units = (long long)B * S;
  for (long long u = gwarp; u < units; u += nwarps) {
    const long long b = u / S;
    const int c0 = (int)(u - b * S) * cps, c1 = min(C, c0 + cps);
    if (c0 >= C) continue;                                    // empty trailing slice
    RowT<D> m, hn;
    row_load<D>(m, h_mt + b * D, lane);
    row_load<D>(hn, h_et + (b * C + c0) * D, lane);
    row_ln_gelu<D>(m, s_gamma, s_beta, lane);
    float qm = 0.f;
#pragma unroll
    for (int i = 0; i < NE; ++i) qm = fmaf(m.v[i], m.v[i], qm);
    const float nm = fmaxf(sqrtf(warp_sum(qm)), 1e-8f);
    for (int c = c0; c < c1; ++c) {
      const long long r = b * C + c;
      RowT<D> e = hn;
      if (c + 1 < c1) row_load<D>(hn, h_et + (r + 1) * D, lane);
      if (lane == 0) {
        long long pr = -1;
        if (c + 3 < c1) {
          pr = r + 3;
        } else if (u + nwarps < units) {
          const long long nb = (u + nwarps) / S;
          pr = nb * C + min((int)(u + nwarps - nb * S) * cps + (c + 3 - c1), C - 1);
        }
        if (pr >= 0) l2_prefetch(h_et + pr * D, D * 4);
      }
      row_ln_gelu<D>(e, s_gamma, s_beta, lane);
      float q = 0.f, d = 0.f;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        q = fmaf(e.v[i], e.v[i], q);
        d = fmaf(e.v[i], m.v[i], d);
      }
      q = warp_sum(q);
      d = warp_sum(d);
      if (lane == 0) scores[r] = d / (nm * fmaxf(sqrtf(q), 1e-8f));
    }
  }
}

static int g_score_fwd_variant = -1;     // -1 auto, 0 CTA-per-mention, 1 warp-per-mention (test hook)
void debug_set_score_fwd_variant(int v) { g_score_fwd_variant = v; }

int score_fwd(cudaStream_t stream, int D, const float* h_mt, const float* h_et, const float* gamma, const float* beta,
              int B, int C, float* scores) {
  prof::Scope prof_scope(stream, prof::SCORE);
  if (D != 768) return fail(DRIN_ERR_ARG, "score: gcn_embed_dim %d not built (768 only)", D);
  constexpr int WNW = 8, WGRID = 148 * 2;     // 2 CTAs / SM (<= 128 registers)
  const int S = row_kernel_slices(B, C);
  const bool warp_kernel = g_score_fwd_variant < 0 ? (long long)B * S >= 148 * 8 : g_score_fwd_variant >= 1;
  if (warp_kernel) {
    score_warp_kernel<768, WNW><<<WGRID, WNW * 32, 0, stream>>>(h_mt, h_et, gamma, beta, B, C, S, scores);
    DRIN_LAUNCH_CHECK();
    return DRIN_OK;
  }
  const int grid = B < 148 * 8 ? B : 148 * 8;
  if (C < 32) score_kernel<768, 4><<<grid, 128, 0, stream>>>(h_mt, h_et, gamma, beta, B, C, scores);
  else score_kernel<768, 8><<<grid, 256, 0, stream>>>(h_mt, h_et, gamma, beta, B, C, scores);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
