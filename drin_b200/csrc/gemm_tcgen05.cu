// tcgen05 / TMEM / TMA GEMM family for the DRIN projections and GCN updates (sm_100a only).
//
//   C[M,N] (+bias) = op(A) * op(B), bf16 operands, fp32 accumulation in TMEM.
//
// Two numeric modes share one kernel:
//   * planes == 1  : plain bf16 operands, one tensor-core pass            ("bf16 mode")
//   * planes == 2  : every fp32 operand is carried as two bf16 planes x = hi + lo; the kernel issues
//                    hi*hi + hi*lo + lo*hi into the same TMEM accumulator ("split-bf16", fp32 parity:
//                    per-product relative error <= 3 * 2^-18 ~ 1.1e-5).
// Three layouts (gemm.cuh): NT (forward), NN (data gradient), TN (weight gradient, split-K).  The
// transposed operands are fed MN-major straight from the row-major tensors -- no transposes in HBM.
//
// Structure: persistent CTAs (grid = min(tiles, #SM)), 128x256 output tile, BLOCK_K = 64,
// warp 0 = TMA producer, warp 1 = MMA issuer (single thread), warps 2..5 = epilogue
// (tcgen05.ld -> bias -> global / split planes).  Two 256-column TMEM accumulators let the epilogue
// of tile i overlap the MMAs of tile i+1.  All mbarrier waits carry a watchdog so a protocol bug traps
// instead of hanging the GPU.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "gemm.cuh"
#include "kernels.cuh"
#include "pipe.cuh"
#include "rows.cuh"

namespace drin {

static constexpr int BM = 128;
static constexpr int BN = 256;
static constexpr int BK = 64;                 // 64 bf16 = 128 B = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;
static constexpr int A_PLANE_BYTES = BM * BK * 2;   // 16 KB
static constexpr int B_PLANE_BYTES = BN * BK * 2;   // 32 KB (16 KB per CTA in cta_group::2 mode)
static constexpr int ACC_STAGES = 2;
static constexpr int TMEM_COLS = ACC_STAGES * BN;   // 512
static constexpr int MAX_STAGES = 6;
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue.  EIGHT epilogue warps: two per TMEM lane quarter
// (a warp may only read the quarter warp_id % 4), the first four take columns 0..127 of the 256-column accumulator,
// the other four columns 128..255.  With four warps the store path of a 128 x 256 tile (128 KB of fp32 per CTA) was
// on the critical path of the single-pass (bf16 mode) and K = 768 GEMMs: 126 us vs 84 us without stores at
// M = 90112, N = K = 768 (scripts/gemm_bf16_probe.py).
static constexpr int EPI_WARPS = 8;
static constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
static constexpr int SMEM_TILE_BYTES = 192 * 1024;
static constexpr int EPI_LD = 32;                                    // staging row (floats): XOR-swizzled, no padding
static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * EPI_LD * 4;  // one 32 x 32 fp32 staging tile per epilogue warp
static constexpr int SMEM_TOTAL_BYTES = SMEM_TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;

struct GemmKernelParams {
  long long M;          // output rows
  int N;                // output cols
  int planes;           // 1 or 2
  int stages;           // smem ring depth
  int ksplit;           // number of contraction slices
  int kblocks_total;    // ceil(K / BK)
  int kblocks_per_split;
  int m_tiles, n_tiles;
  float* C;
  long long c_split_stride;   // elements between split-K partials
  int ldc;
  const float* bias;
  bf16* out_hi;
  bf16* out_lo;
  int ld_planes;
  uint32_t mn_lbo, mn_sbo;    // MN-major descriptor strides (bytes)
  int* error_flag;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier / bulk-copy helpers live in pipe.cuh)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// cta_group::2 forms: one MMA spans the CTA pair (M = 256: 128 rows of A and half of B from each CTA's smem,
// 128 accumulator rows in each CTA's TMEM); issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// TMA load issued by either CTA of the pair; transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 rem;\n\t"
      "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format, cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//  [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (2ull << 61);
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN, int CTAS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                    const GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-B alignment
  const uint32_t bar_base = smem_base + SMEM_TILE_BYTES;
  // barrier block: full[6] empty[6] tmem_full[2] tmem_empty[2] (8 B each) + tmem base (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + ACC_STAGES + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 2 * ACC_STAGES);
  const uint32_t stage_smem = bar_base + 256u;           // epilogue transpose staging (16-B aligned)
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  constexpr int BN_CTA = BN / CTAS;                      // B rows this CTA stages (the pair's MMA spans all BN)
  constexpr int B_PLANE = BN_CTA * BK * 2;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int stage_bytes = p.planes * (A_PLANE_BYTES + B_PLANE);
  const int total_tiles = p.m_tiles * p.n_tiles * p.ksplit;      // m_tiles counts (CTAS * 128)-row tiles
  const int worker = blockIdx.x / CTAS, nworkers = gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p.planes == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS * CTAS);     // one arrive per epilogue warp of every CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();          // peer barriers are initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_coords = [&](int t, int& m_blk, int& n_blk, int& ks) {
    const int per_split = p.m_tiles * p.n_tiles;
    ks = t / per_split;
    const int r = t - ks * per_split;
    m_blk = r / p.n_tiles;
    n_blk = r - m_blk * p.n_tiles;
  };
  auto kblock_range = [&](int ks, int& kb0, int& kb1) {
    kb0 = ks * p.kblocks_per_split;
    kb1 = min(p.kblocks_total, kb0 + p.kblocks_per_split);
  };

  if (warp == 0) {
    // ================================ TMA producer (every CTA loads its own half) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = worker; t < total_tiles; t += nworkers) {
        int m_blk, n_blk, ks, kb0, kb1;
        tile_coords(t, m_blk, n_blk, ks);
        kblock_range(ks, kb0, kb1);
        const int m0 = (m_blk * CTAS + (int)cta_rank) * BM;
        const int n0 = n_blk * BN + (int)cta_rank * BN_CTA;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.error_flag, 1);
          const uint32_t fb = full_bar(stage);
          if (leader) mbar_arrive_expect_tx(fb, (uint32_t)(CTAS * stage_bytes));
          const uint32_t sA = smem_base + stage * stage_bytes;
          const uint32_t sB = sA + p.planes * A_PLANE_BYTES;
          const int k0 = kb * BK;
          auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
            if (CTAS == 1) tma_load_2d(dst, map, fb, c0, c1);
            else tma_load_2d_2cta(dst, map, fb, c0, c1);
          };
          for (int pl = 0; pl < p.planes; ++pl) {
            const CUtensorMap* ma = pl ? &tmA1 : &tmA0;
            const CUtensorMap* mb = pl ? &tmB1 : &tmB0;
            const uint32_t a_dst = sA + pl * A_PLANE_BYTES;
            const uint32_t b_dst = sB + pl * B_PLANE;
            if (!A_MN) {
              load(a_dst, ma, k0, m0);                                   // box {64 k, 128 rows}
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)                          // box {64 m, 64 k-rows}
                load(a_dst + j * (BK * 128), ma, m0 + 64 * j, k0);
            }
            if (!B_MN) {
              load(b_dst, mb, k0, n0);                                   // box {64 k, BN_CTA rows}
            } else {
#pragma unroll
              for (int j = 0; j < BN_CTA / 64; ++j)
                load(b_dst + j * (BK * 128), mb, n0 + 64 * j, k0);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA, one thread) ==================================
    if (lane == 0 && leader) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): fp32 accum, bf16 x bf16
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((BM * CTAS) >> 4) << 24);
      const uint32_t a_lbo = A_MN ? p.mn_lbo : 16u, a_sbo = A_MN ? p.mn_sbo : 1024u;
      const uint32_t b_lbo = B_MN ? p.mn_lbo : 16u, b_sbo = B_MN ? p.mn_sbo : 1024u;
      const uint32_t a_kstep = A_MN ? UMMA_K * 128u : UMMA_K * 2u;      // bytes per UMMA_K step
      const uint32_t b_kstep = B_MN ? UMMA_K * 128u : UMMA_K * 2u;
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
        if (CTAS == 1) umma_bf16(d, ad, bd, idesc, acc);
        else umma_bf16_2cta(d, ad, bd, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {
        if (CTAS == 1) umma_commit(bar);
        else umma_commit_2cta(bar);
      };
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = worker; t < total_tiles; t += nworkers) {
        int m_blk, n_blk, ks, kb0, kb1;
        tile_coords(t, m_blk, n_blk, ks);
        kblock_range(ks, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.error_flag, 2);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accumulate = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, p.error_flag, 3);
          tcgen05_fence_after();
          const uint32_t sA = smem_base + stage * stage_bytes;
          const uint32_t sB = sA + p.planes * A_PLANE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a0 = make_smem_desc(sA + k * a_kstep, a_lbo, a_sbo);
            const uint64_t b0 = make_smem_desc(sB + k * b_kstep, b_lbo, b_sbo);
            mma(d_tmem, a0, b0, accumulate);
            accumulate = 1;
            if (p.planes == 2) {
              const uint64_t a1 = make_smem_desc(sA + A_PLANE_BYTES + k * a_kstep, a_lbo, a_sbo);
              const uint64_t b1 = make_smem_desc(sB + B_PLANE + k * b_kstep, b_lbo, b_sbo);
              mma(d_tmem, a0, b1, 1u);     // hi * lo
              mma(d_tmem, a1, b0, 1u);     // lo * hi
            }
          }
          commit(empty_bar(stage));                 // frees the smem slot (in both CTAs) when these MMAs retire
          if (kb == kb1 - 1) commit(tfull_bar(acc));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) commit(tfull_bar(acc));     // empty slice (cannot happen; keeps protocol live)
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue (8 warps: 4 lane quarters x 2 column halves) =================
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;                   // column half of the 256-column accumulator (0 | 1)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = worker; t < total_tiles; t += nworkers) {
      int m_blk, n_blk, ks;
      tile_coords(t, m_blk, n_blk, ks);
      mbar_wait(tfull_bar(acc), acc_phase, p.error_flag, 4);
      tcgen05_fence_after();
      // Each thread owns one accumulator row in TMEM (32x32b shape); an XOR-swizzled 32 x 32 smem transpose (float4
      // slot s of row r lives at slot s ^ (r & 7): conflict-free both ways, no padding) turns that into row-contiguous
      // 128-B global stores (4 rows x 128 B per warp instruction).
      const long long row_base = ((long long)m_blk * CTAS + cta_rank) * BM + q * 32;
      float* stg = reinterpret_cast<float*>(smem_raw + (stage_smem - smem_u32(smem_raw))) + (warp - 2) * (32 * EPI_LD);
      const int sub = lane >> 3, l8 = lane & 7;
      const long long rows_left = p.M - row_base;                     // rows of this warp that exist (may be <= 0)
      const int rows_valid = rows_left >= 32 ? 32 : (rows_left > 0 ? (int)rows_left : 0);
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const int ncols = p.N - n_blk * BN;                             // columns of this tile that exist
      const int nchunks = ncols >= BN ? BN / 32 : (ncols + 31) / 32;  // warp-uniform
      // base pointers of the first row this thread stores (row sub, columns 4*l8 .. +3 of the chunk)
      float* c_ptr = p.C ? p.C + (long long)ks * p.c_split_stride + (row_base + sub) * p.ldc + n_blk * BN + 4 * l8 : nullptr;
      bf16* hi_ptr = p.out_hi ? p.out_hi + (row_base + sub) * p.ld_planes + n_blk * BN + 4 * l8 : nullptr;
      bf16* lo_ptr = p.out_lo ? p.out_lo + (row_base + sub) * p.ld_planes + n_blk * BN + 4 * l8 : nullptr;

      // one chunk = 32 accumulator columns: registers -> swizzled smem transpose -> 128-B row segments in global.
      // The TMEM load of chunk i + 1 is issued before chunk i is stored, so its latency hides behind the stores;
      // once the last load has landed the accumulator is handed back to the MMA warp.
      auto store_chunk = [&](int chunk, const uint32_t (&v)[32]) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(stg + lane * EPI_LD + ((((i >> 2) ^ lane) & 7) << 2)) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
        const int ccol = chunk * 32 + 4 * l8;                         // column inside the tile
        const int col = n_blk * BN + ccol;
        const bool full_cols = chunk * 32 + 32 <= ncols;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          if (full_cols || col + 3 < p.N) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          else {
            if (col < p.N) bv.x = __ldg(p.bias + col);
            if (col + 1 < p.N) bv.y = __ldg(p.bias + col + 1);
            if (col + 2 < p.N) bv.z = __ldg(p.bias + col + 2);
          }
        }
        if (full_cols) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + sub;
            float4 f = *reinterpret_cast<const float4*>(stg + r * EPI_LD + (((l8 ^ r) & 7) << 2));
            f.x += bv.x; f.y += bv.y; f.z += bv.z; f.w += bv.w;
            if (r < rows_valid) {
              if (c_ptr) *reinterpret_cast<float4*>(c_ptr + (long long)it * 4 * p.ldc + chunk * 32) = f;
              if (hi_ptr) {
                uint32_t h0, l0, h1, l1;
                split_bf16x2(f.x, f.y, h0, l0);
                split_bf16x2(f.z, f.w, h1, l1);
                *reinterpret_cast<uint2*>(hi_ptr + (long long)it * 4 * p.ld_planes + chunk * 32) = make_uint2(h0, h1);
                if (lo_ptr) *reinterpret_cast<uint2*>(lo_ptr + (long long)it * 4 * p.ld_planes + chunk * 32) = make_uint2(l0, l1);
              }
            }
          }
        } else {                                                       // ragged last chunk of the tile (N % 32 != 0)
#pragma unroll 1
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + sub;
            const long long row = row_base + r;
            float4 f = *reinterpret_cast<const float4*>(stg + r * EPI_LD + (((l8 ^ r) & 7) << 2));
            f.x += bv.x; f.y += bv.y; f.z += bv.z; f.w += bv.w;
            const float ff[4] = {f.x, f.y, f.z, f.w};
            if (r < rows_valid) {
              for (int j = 0; j < 4 && col + j < p.N; ++j) {
                if (p.C) p.C[(long long)ks * p.c_split_stride + row * p.ldc + col + j] = ff[j];
                if (p.out_hi) {
                  bf16 h, l;
                  split_bf16(ff[j], h, l);
                  p.out_hi[row * p.ld_planes + col + j] = h;
                  if (p.out_lo) p.out_lo[row * p.ld_planes + col + j] = l;
                }
              }
            }
          }
        }
        __syncwarp();
      };

      // this warp's chunks: column half `chalf` of the accumulator, clipped to the columns that exist
      const int cbeg = chalf * (BN / 64);
      const int cend = nchunks < cbeg + BN / 64 ? nchunks : cbeg + BN / 64;
      auto release_acc = [&]() {                                       // this warp has read all it needs from TMEM
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 1) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(tempty_bar(acc), 0u);               // the leader's MMA thread waits for both CTAs
        }
      };
      uint32_t va[32], vb[32];
      if (cbeg < cend) tmem_ld_32x32(taddr0 + (uint32_t)(cbeg * 32), va);
      else release_acc();                                              // ragged last N tile: nothing in this half
#pragma unroll 1
      for (int chunk = cbeg; chunk < cend; chunk += 2) {
        tmem_ld_wait();                                                // va = chunk
        if (chunk + 1 < cend) tmem_ld_32x32(taddr0 + (uint32_t)((chunk + 1) * 32), vb);
        else release_acc();
        store_chunk(chunk, va);
        if (chunk + 1 < cend) {
          tmem_ld_wait();                                              // vb = chunk + 1
          if (chunk + 2 < cend) tmem_ld_32x32(taddr0 + (uint32_t)((chunk + 2) * 32), va);
          else release_acc();
          store_chunk(chunk + 1, vb);
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();          // the peer may still be read / written by the leader's MMAs
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTAS == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                   : "memory");
  }
}

// deterministic split-K reduction: C[i] = bias + sum_s partial[s][i]
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, long long split_stride,
                                     float* __restrict__ C, long long M, int N, int ldc,
                                     const float* __restrict__ bias) {
  const long long total4 = M * (long long)(N / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (N / 4);
    const int c = (int)(i - r * (N / 4)) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) acc = *reinterpret_cast<const float4*>(bias + c);
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(partial + s * split_stride + r * ldc + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(C + r * ldc + c) = acc;
  }
}

// all deferred split-K reductions of a backward pass in one launch: blockIdx.y selects the job
__global__ void splitk_reduce_multi_kernel(const SplitKJobs jobs) {
  const SplitKJob j = jobs.job[blockIdx.y];
  const int n4 = j.N / 4;
  const long long total4 = j.M * (long long)n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n4;
    const int c = (int)(i - r * n4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j.bias) acc = *reinterpret_cast<const float4*>(j.bias + c);
    for (int s = 0; s < j.ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(j.partial + s * j.split_stride + r * j.ldc + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(j.C + r * j.ldc + c) = acc;
  }
}

int splitk_reduce_multi(cudaStream_t stream, const SplitKJobs& jobs) {
  if (jobs.count <= 0) return DRIN_OK;
  prof::Scope prof_scope(stream, prof::GEMM);
  splitk_reduce_multi_kernel<<<dim3(148, jobs.count), 256, 0, stream>>>(jobs);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct TmapKey {
  const void* ptr; long long rows; int cols, ld, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<long long>()(k.rows);
    h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)k.box_rows;
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

// bf16 row-major [rows, cols] (leading dim ld) -> 2-D map with box {64 cols, box_rows}, SWIZZLE_128B
static int make_tmap(const bf16* ptr, long long rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  if (((uintptr_t)ptr & 15) || (ld % 8) != 0)
    return fail(DRIN_ERR_ARG, "GEMM operand must be 16-byte aligned with ld %% 8 == 0 (ptr=%p ld=%d)", ptr, ld);
  TmapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lock(g_tmap_mutex);
  auto it = g_tmap_cache.find(key);
  if (it != g_tmap_cache.end()) {
    *out = it->second;
    return DRIN_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(DRIN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DRIN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d ld=%d box_rows=%d", (int)r, rows,
                cols, ld, box_rows);
  if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return DRIN_OK;
}

static int g_mn_lbo = 0, g_mn_sbo = 0;
void gemm_debug_set_mn_desc(int lbo_bytes, int sbo_bytes) {
  g_mn_lbo = lbo_bytes;
  g_mn_sbo = sbo_bytes;
}

// Per-device state: one process may drive several GPUs (function attributes, the SM count and the device-side
// error flag all belong to the device that was current when they were set up).
constexpr int MAX_DEVICES = 64;
struct GemmDeviceState {
  int status = -1;
  int num_sms = 0;
  int* error_flag = nullptr;
};
static GemmDeviceState g_dev[MAX_DEVICES];
static std::mutex g_dev_mutex;
static bool g_force_1cta = false;
static int g_sm_cap = 0;      // > 0: persistent GEMM grids use at most this many SMs (the rest stay free for kernels of a
                              // concurrent stream); 0 = all
void debug_set_gemm_sm_cap(int v) { g_sm_cap = v > 0 ? v : 0; }

template <bool A_MN, bool B_MN, int CTAS>
static cudaError_t launch_gemm(int grid, cudaStream_t stream, const CUtensorMap& tA0, const CUtensorMap& tA1,
                               const CUtensorMap& tB0, const CUtensorMap& tB1, const GemmKernelParams& p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<A_MN, B_MN, CTAS>, tA0, tA1, tB0, tB1, p);
}

static int gemm_init_device(GemmDeviceState** out) {
  int dev = 0;
  DRIN_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) return fail(DRIN_ERR_ARG, "gemm: device ordinal %d out of range", dev);
  GemmDeviceState& st = g_dev[dev];
  *out = &st;
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (st.status >= 0) return st.status;
  DRIN_CUDA(cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, dev));
#define DRIN_GEMM_ATTR(A, B, C)                                                                               \
  DRIN_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<A, B, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 SMEM_TOTAL_BYTES))
  DRIN_GEMM_ATTR(false, false, 1);
  DRIN_GEMM_ATTR(false, true, 1);
  DRIN_GEMM_ATTR(true, true, 1);
  DRIN_GEMM_ATTR(false, false, 2);
  DRIN_GEMM_ATTR(false, true, 2);
  DRIN_GEMM_ATTR(true, true, 2);
#undef DRIN_GEMM_ATTR
  {
    const char* e = getenv("DRIN_GEMM_1CTA");
    g_force_1cta = e && e[0] == '1';
  }
  DRIN_CUDA(cudaMalloc(&st.error_flag, sizeof(int)));
  DRIN_CUDA(cudaMemset(st.error_flag, 0, sizeof(int)));
  st.status = DRIN_OK;
  return st.status;
}

int gemm_tcgen05(cudaStream_t stream, GemmLayout layout, const Operand& A, const Operand& B, long long M, int N,
                 long long K, const GemmEpilogue& ep, int ksplit, float* partial, SplitKJobs* defer) {
  prof::Scope prof_scope(stream, prof::GEMM, 2.0 * (double)M * (double)N * (double)K, 0);
  GemmDeviceState* dstate = nullptr;
  DRIN_TRY(gemm_init_device(&dstate));
  const int g_num_sms = dstate->num_sms;
  if (M <= 0 || N <= 0 || K <= 0) return fail(DRIN_ERR_ARG, "gemm: empty problem M=%lld N=%d K=%lld", M, N, K);
  const int planes = A.lo ? 2 : 1;
  if ((A.lo != nullptr) != (B.lo != nullptr)) return fail(DRIN_ERR_ARG, "gemm: A and B must have the same planes");
  if (!ep.C && !ep.out_hi) return fail(DRIN_ERR_ARG, "gemm: no output requested");
  if (ep.C && (ep.ldc % 4)) return fail(DRIN_ERR_ARG, "gemm: ldc must be a multiple of 4");
  if (ep.out_hi && (ep.ld_planes % 8)) return fail(DRIN_ERR_ARG, "gemm: ld_planes must be a multiple of 8");
  const bool a_mn = layout == GEMM_TN, b_mn = layout != GEMM_NT;
  // operand shape checks
  const long long a_rows = a_mn ? K : M;
  const long long a_cols = a_mn ? M : K;
  const long long b_rows = b_mn ? K : N;
  const long long b_cols = b_mn ? N : K;
  if (A.rows != a_rows || A.cols != a_cols || B.rows != b_rows || B.cols != b_cols)
    return fail(DRIN_ERR_ARG, "gemm: operand shapes do not match layout %d (A %lldx%d, B %lldx%d, M=%lld N=%d K=%lld)",
                (int)layout, A.rows, A.cols, B.rows, B.cols, M, N, K);

  // cta_group::2 (256 x 256 tile per CTA pair, each CTA stages half of B) whenever there are >= 2 row tiles
  const int ctas = (!g_force_1cta && M > BM) ? 2 : 1;
  CUtensorMap tA0, tA1, tB0, tB1;
  const int a_box = a_mn ? BK : BM, b_box = b_mn ? BK : BN / ctas;
  DRIN_TRY(make_tmap(A.hi, A.rows, A.cols, A.ld, a_box, &tA0));
  DRIN_TRY(make_tmap(B.hi, B.rows, B.cols, B.ld, b_box, &tB0));
  if (planes == 2) {
    DRIN_TRY(make_tmap(A.lo, A.rows, A.cols, A.ld, a_box, &tA1));
    DRIN_TRY(make_tmap(B.lo, B.rows, B.cols, B.ld, b_box, &tB1));
  } else {
    tA1 = tA0;
    tB1 = tB0;
  }

  GemmKernelParams p{};
  p.M = M;
  p.N = N;
  p.planes = planes;
  p.stages = ctas == 2 ? (planes == 2 ? 3 : 6) : (planes == 2 ? 2 : 4);      // 192 KB ring either way
  p.kblocks_total = (int)((K + BK - 1) / BK);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.kblocks_total) ksplit = p.kblocks_total;
  p.kblocks_per_split = (p.kblocks_total + ksplit - 1) / ksplit;
  ksplit = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;   // no empty slices
  p.ksplit = ksplit;
  p.m_tiles = (int)((M + (long long)BM * ctas - 1) / ((long long)BM * ctas));   // tiles of the CTA pair
  p.n_tiles = (N + BN - 1) / BN;
  p.bias = ksplit > 1 ? nullptr : ep.bias;
  p.ldc = ep.ldc;
  p.out_hi = ep.out_hi;
  p.out_lo = ep.out_lo;
  p.ld_planes = ep.ld_planes;
  p.mn_lbo = g_mn_lbo ? (uint32_t)g_mn_lbo : (uint32_t)(BK * 128);
  p.mn_sbo = g_mn_sbo ? (uint32_t)g_mn_sbo : 1024u;
  p.error_flag = dstate->error_flag;
  if (ksplit > 1) {
    if (!partial || !ep.C || ep.out_hi || (N % 4))
      return fail(DRIN_ERR_ARG, "gemm: split-K needs a partial buffer, an fp32 output with N %% 4 == 0 and no planes");
    p.C = partial;
    p.c_split_stride = M * (long long)ep.ldc;
  } else {
    p.C = ep.C;
    p.c_split_stride = 0;
  }
  const long long tiles = (long long)p.m_tiles * p.n_tiles * ksplit;
  const int usable_sms = (g_sm_cap > 0 && g_sm_cap < g_num_sms) ? g_sm_cap : g_num_sms;
  const int workers = usable_sms / ctas > 0 ? usable_sms / ctas : 1;
  const int grid = ctas * (int)(tiles < workers ? tiles : workers);
  cudaError_t le;
  if (ctas == 1) {
    le = layout == GEMM_NT   ? launch_gemm<false, false, 1>(grid, stream, tA0, tA1, tB0, tB1, p)
         : layout == GEMM_NN ? launch_gemm<false, true, 1>(grid, stream, tA0, tA1, tB0, tB1, p)
                             : launch_gemm<true, true, 1>(grid, stream, tA0, tA1, tB0, tB1, p);
  } else {
    le = layout == GEMM_NT   ? launch_gemm<false, false, 2>(grid, stream, tA0, tA1, tB0, tB1, p)
         : layout == GEMM_NN ? launch_gemm<false, true, 2>(grid, stream, tA0, tA1, tB0, tB1, p)
                             : launch_gemm<true, true, 2>(grid, stream, tA0, tA1, tB0, tB1, p);
  }
  DRIN_CUDA(le);
  DRIN_LAUNCH_CHECK();
  if (ksplit > 1 && defer) {
    if (defer->count >= SplitKJobs::MAX) return fail(DRIN_ERR_ARG, "gemm: too many deferred split-K reductions");
    defer->job[defer->count++] = SplitKJob{partial, p.c_split_stride, ep.C, ep.bias, M, N, ep.ldc, ksplit};
  } else if (ksplit > 1) {
    const long long total4 = M * (long long)(N / 4);
    const int rgrid = (int)((total4 + 255) / 256 < 4 * g_num_sms ? (total4 + 255) / 256 : 4 * g_num_sms);
    splitk_reduce_kernel<<<rgrid, 256, 0, stream>>>(partial, ksplit, p.c_split_stride, ep.C, M, N, ep.ldc, ep.bias);
    DRIN_LAUNCH_CHECK();
  }
  return DRIN_OK;
}

// ---------------------------------------------------------------------------------------------
// CUDA-core fp32 reference (tests / bring-up only)
// ---------------------------------------------------------------------------------------------
__global__ void gemm_reference_kernel(int layout, const bf16* __restrict__ a_hi, const bf16* __restrict__ a_lo, int lda,
                                      const bf16* __restrict__ b_hi, const bf16* __restrict__ b_lo, int ldb,
                                      long long M, int N, long long K, float* __restrict__ C, int ldc,
                                      const float* __restrict__ bias, bf16* out_hi, bf16* out_lo, int ldp) {
  const long long m = blockIdx.y * (long long)blockDim.y + threadIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (long long k = 0; k < K; ++k) {
    const long long ai = layout == GEMM_TN ? k * lda + m : m * lda + k;
    const long long bi = layout == GEMM_NT ? (long long)n * ldb + k : k * ldb + n;
    float a = __bfloat162float(a_hi[ai]);
    float b = __bfloat162float(b_hi[bi]);
    if (a_lo) a += __bfloat162float(a_lo[ai]);
    if (b_lo) b += __bfloat162float(b_lo[bi]);
    acc = fmaf(a, b, acc);
  }
  if (bias) acc += bias[n];
  if (C) C[m * ldc + n] = acc;
  if (out_hi) {
    bf16 h, l;
    split_bf16(acc, h, l);
    out_hi[m * ldp + n] = h;
    if (out_lo) out_lo[m * ldp + n] = l;
  }
}

int gemm_reference_simt(cudaStream_t stream, GemmLayout layout, const Operand& A, const Operand& B, long long M,
                        int N, long long K, const GemmEpilogue& ep) {
  dim3 block(32, 8);
  dim3 grid((N + 31) / 32, (unsigned)((M + 7) / 8));
  gemm_reference_kernel<<<grid, block, 0, stream>>>((int)layout, A.hi, A.lo, A.ld, B.hi, B.lo, B.ld, M, N, K, ep.C,
                                                    ep.ldc, ep.bias, ep.out_hi, ep.out_lo, ep.ld_planes);
  DRIN_LAUNCH_CHECK();
  return DRIN_OK;
}

}  // namespace drin
