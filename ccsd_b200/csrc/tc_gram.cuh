// tc_gram.cuh -- Gram product of the rank-2 tensor on the 5th-gen tensor cores.
//
//   G[b] = F[b] . [F[b] ; Wp]^T     (E x K) . (K x (E + PR0))
//   H    = G[:, :E] * (1 - I)        hodge_laplacian + default_mask   (cc_utils.py:917-942, 964-969)
//   P0   = G[:, E:]                  rank2 @ W_{q,k} of hodge layer 0 (hodge_layers.py:185)
//
// fp32 parity needs more than one bf16 pass: every operand is split x = hi + lo (two bf16) and the
// product is accumulated as hi.hi + hi.lo + lo.hi in the fp32 TMEM accumulator (dropped lo.lo term
// ~2^-18 relative) -- "bf16x3".
//
// One persistent CTA per SM, 9 warps:
//   warps 0-3  producers: coalesced fp32 loads of F rows (+ the projection rows of the weight blob),
//              hi/lo split, st.shared into the canonical K-major SWIZZLE_128B operand layout
//              (row r, 16-byte chunk c at (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) * 16)), one stage =
//              64 k-values x 256 rows x {hi, lo} = 64 KB, 3-stage mbarrier ring;
//   warp  8    one elected thread issues tcgen05.mma (M=128, N=ncols, K=16; 2 M tiles x 3 terms x 4
//              k-steps per stage), tcgen05.commit frees the stage / publishes the accumulator;
//   warps 4-7  epilogue: tcgen05.ld (32 lanes x 16 columns) -> masked stores of H and P0.
// A and B read the SAME shared-memory rows (rows < E are F; rows wp0.. are Wp), so F is converted once.
#pragma once
#include "plan_dev.h"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TG_THREADS = 288;
constexpr int TG_STAGES = 3;
constexpr int TG_BK = 64;
constexpr uint32_t TG_HALF = 256 * 128;   // bytes of the hi (or lo) half of a stage
constexpr uint32_t TG_STAGE = 2 * TG_HALF;
constexpr size_t TG_SMEM = (size_t)TG_STAGES * TG_STAGE + 1024 /*align*/ + 256 /*barriers*/;

struct TcGramArgs {
  const float *r2;  // [B,E,K]
  float *H;         // [B,E,E]
  float *P0;        // [B,E,PR0]
};

static inline int tc_gram_wp0(int E) { return (E + 7) & ~7; }
static inline int tc_gram_ncols(int E, int PR0) { return (tc_gram_wp0(E) + PR0 + 15) & ~15; }
static inline int tc_gram_supported(int E, int K, int PR0) {
  (void)K;
  return E >= 8 && E <= 248 && tc_gram_ncols(E, PR0) <= 256;
}

__global__ void __launch_bounds__(TG_THREADS, 1) tc_gram_kernel(const DevPlan *__restrict__ P, TcGramArgs a) {
  extern __shared__ uint8_t tg_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const int E = d.E, K = d.K, PR0 = P->PR0, Kw = P->Kp, B = d.B;
  const int wp0 = (E + 7) & ~7;
  const int ncols = (wp0 + PR0 + 15) & ~15;
  const int mtiles = E > 128 ? 2 : 1;
  const int nkb = (K + TG_BK - 1) / TG_BK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t raw = tc::smem_u32(tg_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                       // operand stages (1024-aligned)
  uint8_t *gen_base = tg_smem_raw + (base - raw);
  const uint32_t bars = base + TG_STAGES * TG_STAGE;                  // mbarriers
  const uint32_t full0 = bars, empty0 = bars + 8 * TG_STAGES, tfull = bars + 16 * TG_STAGES,
                 tempty = tfull + 8, tslot = tempty + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen_base + TG_STAGES * TG_STAGE + 16 * TG_STAGES + 16);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      tc::mbar_init(full0 + 8 * s, 128);  // every producer thread arrives
      tc::mbar_init(empty0 + 8 * s, 1);   // tcgen05.commit
    }
    tc::mbar_init(tfull, 1);
    tc::mbar_init(tempty, 128);           // every epilogue thread arrives
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(tslot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;

  if (warp < 4) {
    // ===================== producers =====================
    const float *Wp = P->W + d.neta.proj_w;
    const int tasks = (E + PR0) * 8;
    const bool vec = (K & 3) == 0;
    uint32_t it = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      const float *Fb = a.r2 + (size_t)b * E * K;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % TG_STAGES;
        const uint32_t ph = (it / TG_STAGES) & 1;
        tc::mbar_wait(empty0 + 8 * s, ph ^ 1);
        uint8_t *st = gen_base + (size_t)s * TG_STAGE;
        for (int t = threadIdx.x; t < tasks; t += 128) {
          const int rr = t >> 3, c = t & 7;
          const int k = kb * TG_BK + c * 8;
          float x[8];
          int row;
          if (rr < E) {
            row = rr;
            const float *src = Fb + (size_t)rr * K + k;
            if (vec && k + 8 <= K) {
              const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src));
              const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
              x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) x[q] = (k + q < K) ? __ldg(src + q) : 0.f;
            }
          } else {
            row = wp0 + (rr - E);
            const float *src = Wp + (size_t)(rr - E) * Kw + k;   // rows are zero padded to Kw (multiple of 4)
            if (k + 8 <= Kw) {
              const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src));
              const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
              x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) x[q] = (k + q < Kw) ? __ldg(src + q) : 0.f;
            }
          }
          uint4 hi, lo;
          tc::split8(x, hi, lo);
          const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
          *reinterpret_cast<uint4 *>(st + off) = hi;
          *reinterpret_cast<uint4 *>(st + TG_HALF + off) = lo;
        }
        tc::fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
        tc::mbar_arrive(full0 + 8 * s);
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::make_idesc_bf16(128, ncols, 0, 0);
    uint32_t it = 0, tile = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++tile) {
      tc::mbar_wait(tempty, (tile & 1) ^ 1);   // epilogue has drained the previous accumulators
      tc::tc_fence_after_sync();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % TG_STAGES;
        const uint32_t ph = (it / TG_STAGES) & 1;
        tc::mbar_wait(full0 + 8 * s, ph);
        tc::tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t sb = base + (uint32_t)s * TG_STAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t ko = (uint32_t)k4 * 32u;   // 16 bf16 = 32 bytes inside the 128-byte swizzle span
            const uint64_t b_hi = tc::make_smem_desc(sb + ko, 0, 1024);
            const uint64_t b_lo = tc::make_smem_desc(sb + TG_HALF + ko, 0, 1024);
            for (int mt = 0; mt < mtiles; ++mt) {
              const uint64_t a_hi = tc::make_smem_desc(sb + (uint32_t)mt * 16384u + ko, 0, 1024);
              const uint64_t a_lo = tc::make_smem_desc(sb + TG_HALF + (uint32_t)mt * 16384u + ko, 0, 1024);
              const uint32_t dcol = tmem + (uint32_t)(mt * ncols);
              tc::umma_bf16(dcol, a_hi, b_hi, idesc, (kb | k4) != 0);
              tc::umma_bf16(dcol, a_hi, b_lo, idesc, 1);
              tc::umma_bf16(dcol, a_lo, b_hi, idesc, 1);
            }
          }
          tc::umma_commit(empty0 + 8 * s);             // stage may be refilled once these MMAs retire
          if (kb == nkb - 1) tc::umma_commit(tfull);   // accumulators complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 4-7 -> TMEM lane quarters 0-3) =====================
    const int q = warp & 3;
    const int mask_diag = d.netf.use_hodge_mask;
    uint32_t tile = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++tile) {
      tc::mbar_wait(tfull, tile & 1);
      tc::tc_fence_after_sync();
      for (int mt = 0; mt < mtiles; ++mt) {
        const int row = mt * 128 + q * 32 + lane;
        float *Hrow = a.H + ((size_t)b * E + row) * E;
        float *Prow = a.P0 + ((size_t)b * E + row) * PR0;
        for (int c0 = 0; c0 < ncols; c0 += 16) {
          float v[16];
          tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * ncols + c0), v);
          if (row < E) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = c0 + j;
              if (col < E) Hrow[col] = (mask_diag && col == row) ? 0.f : v[j];
              else if (col >= wp0 && col - wp0 < PR0) Prow[col - wp0] = v[j];
            }
          }
        }
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(tempty);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

static inline int tc_gram_prepare() {
  return cudaFuncSetAttribute(tc_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM) == cudaSuccess ? 0 : -1;
}

static inline int tc_gram_launch(const DevPlan *dP, const DevPlan &hp, const float *r2, float *H, float *P0, void *stream) {
  TcGramArgs a;
  a.r2 = r2; a.H = H; a.P0 = P0;
  int grid = hp.d.B < 148 ? hp.d.B : 148;
  tc_gram_kernel<<<grid, TG_THREADS, TG_SMEM, (cudaStream_t)stream>>>(dP, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccsd
