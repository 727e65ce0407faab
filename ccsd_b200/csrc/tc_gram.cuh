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
// One persistent CTA per SM, 13 warps:
//   warps 0-7  producers (all loads of a k-block are issued before the split, >= 56 KB in flight per SM): coalesced fp32 loads of F rows (+ the projection rows of the weight blob),
//              hi/lo split, st.shared into the canonical K-major SWIZZLE_128B operand layout
//              (row r, 16-byte chunk c at (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) * 16)), one stage =
//              64 k-values x 256 rows x {hi, lo} = 64 KB, 3-stage mbarrier ring;
//   warp  12   one elected thread issues tcgen05.mma (M=128, N=ncols, K=16; 2 M tiles x 3 terms x 4
//              k-steps per stage), tcgen05.commit frees the stage / publishes the accumulator;
//   warps 8-11 epilogue: tcgen05.ld (32 lanes x 16 columns) -> masked stores of H and P0.
// A and B read the SAME shared-memory rows (rows < E are F; rows wp0.. are Wp), so F is converted once.
#pragma once
#include "plan_dev.h"
#include "tc_common.cuh"

namespace ccsd {

#ifndef TG_PRODUCER_WARPS
#define TG_PRODUCER_WARPS 16
#endif
constexpr int TG_PROD_WARPS = TG_PRODUCER_WARPS;   // two groups of TG_PROD_WARPS / 2 warps fill alternate k-blocks
constexpr int TG_GRP = TG_PROD_WARPS * 16;          // threads per producer group
constexpr int TG_RSTEP = TG_GRP / 8;               // rows covered by one sweep of a group
constexpr int TG_PROD = TG_PROD_WARPS * 32;
// epilogue warps: TMEM lane quarter = warp % 4, alternate 16-column chunks by warp / 4.  One sample per unit: 4 (its epilogue
// hides behind the next unit's MMAs -- double-buffered accumulators -- and 672 threads leave the producers 96 registers);
// stacked samples: 8 (the accumulators do not always fit twice, so the epilogue is on the critical path)
__host__ __device__ constexpr int tg_epi_warps(bool grouped) { return grouped ? 8 : 4; }
__host__ __device__ constexpr int tg_threads(bool grouped) { return TG_PROD + tg_epi_warps(grouped) * 32 + 32; }   // producers, epilogue warps, 1 MMA warp
constexpr int TG_STAGES = 3;
constexpr int TG_BK = 64;
constexpr uint32_t TG_HALF = 256 * 128;   // bytes of the hi (or lo) half of a stage
constexpr uint32_t TG_STAGE = 2 * TG_HALF;
constexpr size_t TG_SMEM = (size_t)TG_STAGES * TG_STAGE + 1024 /*align*/ + 256 /*barriers*/;

struct TcGramArgs {
  const float *r2;  // [B,E,K]
  float *H;         // [B,E,E]
  float *P0;        // [B,E,PR0]
  const uint8_t *wimg;   // projection rows of the weight blob as bf16 hi / lo operand chunks: [k-block][row][chunk][hi 16 B | lo 16 B]
                         // (tc_gram_prep_kernel, once per run), or nullptr: converted from the fp32 blob in the loop
  long long *trace;  // debug timeline (ccsd_debug_apply_trace): CTA 0, [k-block index < 512][16] clock stamps, or nullptr
  float *Dg, *Rs;   // optional [B,E]: diag(F F^T) (before the (1 - I) mask) and the row sums F 1 (one more all-ones
                    // projection row) -- the Gram quantities the Langevin norm of an affine ScoreNetworkF needs (tc_hnorm.cuh)
};

static inline int tc_gram_wp0(int E) { return (E + 7) & ~7; }
static inline int tc_gram_ncols(int E, int PR0) { return (tc_gram_wp0(E) + PR0 + 15) & ~15; }   // PR0 incl. the ones row
static inline int tc_gram_supported(int E, int K, int PR0) {
  (void)K;
  return E >= 8 && E <= 192 && PR0 <= 64 && tc_gram_ncols(E, PR0) <= 256;
}

// The projection rows (hodge q / k weights, K-major like F) are the same for every sample: converted once per run into the
// operand chunks the producers would otherwise rebuild for every k-block of every sample -- behind a synchronous L2 load
// that sat between the stage wait and the arrive (3 k of the 5.5 k cycles of that section, tools/gram_trace.py).
static __global__ void tc_gram_prep_kernel(const DevPlan *__restrict__ P, uint8_t *__restrict__ img) {
  const ccsd_plan_desc_t &d = P->d;
  const int PR0 = P->PR0, Kw = P->Kp, kb = blockIdx.x;
  const float *Wp = P->W + d.neta.proj_w;
  for (int t = threadIdx.x; t < PR0 * 8; t += blockDim.x) {
    const int rw = t >> 3, c = t & 7, k = kb * TG_BK + c * 8;
    float y[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) y[q] = (k + q < Kw) ? __ldg(Wp + (size_t)rw * Kw + k + q) : 0.f;
    uint4 hi, lo;
    tc::split8(y, hi, lo);
    uint4 *dst = reinterpret_cast<uint4 *>(img + ((size_t)(kb * PR0 + rw) * 8 + c) * 32);
    dst[0] = hi; dst[1] = lo;
  }
}
static inline size_t tc_gram_img_bytes(int K, int PR0) { return (size_t)((K + TG_BK - 1) / TG_BK) * (size_t)(PR0 > 0 ? PR0 : 1) * 256; }

#define TG_STAMP(slot_, i_) do { if (a.trace && blockIdx.x == 0 && (i_) < 512) a.trace[(size_t)(i_) * 16 + (slot_)] = clock64(); } while (0)

template <bool GROUPED>   // GROUPED = false: one sample per work unit (G = 1 folds away at compile time)
__global__ void __launch_bounds__(tg_threads(GROUPED), 1) tc_gram_kernel(const DevPlan *__restrict__ P, TcGramArgs a) {
  extern __shared__ uint8_t tg_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  constexpr int TG_EPI_WARPS = tg_epi_warps(GROUPED);
  const int E = d.E, K = d.K, PR0 = P->PR0, Kw = P->Kp, B = d.B;
  // small complexes: G consecutive samples form one work unit -- their rows are contiguous in the state, so the operand
  // tile is simply taller; the Gram of the stacked rows holds the G per-sample Grams as its diagonal blocks (the
  // cross-sample blocks are computed and never stored)
  const int G = GROUPED ? P->gram_group : 1, EG = G * E, NV = (B + G - 1) / G;
  const int wp0 = (EG + 7) & ~7;
  const int rsum = a.Rs != nullptr;                 // one more operand row: all ones (Gram column = row sums of F)
  const int rcol = wp0 > EG ? EG : wp0 + PR0;       // it takes a pad row of the F block when there is one, else a new column
  const int ncols = (wp0 + PR0 + (rsum && rcol >= wp0 ? 1 : 0) + 15) & ~15;
  const int mtiles = EG > 128 ? 2 : 1;
  // The second M tile (rows 128 ..) only needs the columns from c1 on: for one sample per unit H is symmetric, so the block
  // [128:, :128] is the transpose of what the first tile holds (c1 = 128, the first tile's epilogue writes both orientations);
  // for stacked samples its rows' diagonal blocks start at the block that straddles row 128.
  const int c1 = mtiles == 1 ? 0 : (GROUPED ? ((128 / E) * E) & ~15 : 128), N1 = ncols - c1;
  // accumulators in tensor memory: the first tile's are double buffered when 2 ncols + N1 <= 512 columns, so the epilogue of
  // one unit overlaps the MMAs of the next (the second tile's are drained first and waited for at the start of a unit)
  const int dbuf = 2 * ncols + (mtiles == 2 ? N1 : 0) <= 512;
  const uint32_t acc1 = (uint32_t)((dbuf ? 2 : 1) * ncols);
  const int nkb = (K + TG_BK - 1) / TG_BK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t raw = tc::smem_u32(tg_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                       // operand stages (1024-aligned)
  uint8_t *gen_base = tg_smem_raw + (base - raw);
  const uint32_t bars = base + TG_STAGES * TG_STAGE;                  // mbarriers
  const uint32_t full0 = bars, empty0 = bars + 8 * TG_STAGES, tfull = bars + 16 * TG_STAGES /* [2] */,
                 tempty0 = tfull + 16 /* [2] */, tempty1 = tempty0 + 16, tslot = tempty1 + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen_base + TG_STAGES * TG_STAGE + 16 * TG_STAGES + 40);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      tc::mbar_init(full0 + 8 * s, TG_GRP);   // every thread of the producer group that fills it
      tc::mbar_init(empty0 + 8 * s, 1);   // tcgen05.commit
    }
    for (int q2 = 0; q2 < 2; ++q2) {
      tc::mbar_init(tfull + 8 * q2, 1);
      tc::mbar_init(tempty0 + 8 * q2, TG_EPI_WARPS * 32);   // every epilogue thread arrives
    }
    tc::mbar_init(tempty1, TG_EPI_WARPS * 32);
    tc::mbar_fence_init();
  }
  if (warp == TG_PROD_WARPS + TG_EPI_WARPS) tc::tmem_alloc(tslot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;

  if (warp < TG_PROD_WARPS) {
    // ===================== producers =====================
    // Two independent groups of 4 warps fill alternate k-blocks, so two k-blocks (~100 KB of loads)
    // are in flight per SM without register double buffering.  Within a group, thread t owns the
    // 16-byte chunk c = t & 7 of rows (t >> 3) + 16 j: global and shared addresses advance by constant
    // strides and the 24 loads of a k-block are independent.
    const float *Wp = P->W + d.neta.proj_w;
    const bool vec = (K & 3) == 0;
    const bool vec2 = (K & 1) == 0 && (reinterpret_cast<uintptr_t>(a.r2) & 7) == 0;
    constexpr int NT = 192 / TG_RSTEP;          // RSTEP * NT = 192 rows of F per k-block (E <= 192)
    const int grp = warp / (TG_PROD_WARPS / 2), tg = threadIdx.x - grp * TG_GRP;
    const int c = tg & 7, r0 = tg >> 3;
    const uint32_t off0 = (uint32_t)(r0 >> 3) * 1024u + (uint32_t)(r0 & 7) * 128u + (uint32_t)((c ^ (r0 & 7)) << 4);
    // k-block g of this CTA = (unit vb, block kb); this group takes every other one.  32-bit incremental bookkeeping (the
    // 64-bit divisions of the first version were a quarter of the producers' instructions)
    int vb = (int)blockIdx.x, kb = grp;
    while (kb >= nkb && vb < NV) { kb -= nkb; vb += (int)gridDim.x; }
    int s = grp % TG_STAGES;
    uint32_t ph = (uint32_t)(grp / TG_STAGES) & 1;
    int gi = grp;   // k-block index of this CTA (debug timeline)
    // Software pipeline over the group's k-blocks, in two halves of NT / 2 rows: while one half of the CURRENT k-block is split
    // and stored, the same half of the NEXT k-block is already being loaded into the registers it frees.  (The first version
    // issued a k-block's loads, waited for them, converted, and only then issued the next ones: 6.3 k cycles per k-block and
    // group against 1.9 k of MMA time per k-block -- tools/gram_trace.py.)
    constexpr int NH = NT / 2;
    auto load_half = [&](float (&xh)[NH][8], int vb_, int kb_, int h) {
      const float *Fb = a.r2 + (size_t)vb_ * EG * K;
      const int rows_valid = ((B - vb_ * G < G) ? B - vb_ * G : G) * E;   // the last unit may hold fewer samples
      const int k = kb_ * TG_BK + c * 8;
      const bool fast = vec && (kb_ * TG_BK + TG_BK <= K);
      const float *src = Fb + (size_t)r0 * K + k;
#pragma unroll
      for (int j3 = 0; j3 < NH; ++j3) {
        const int j = h * NH + j3;
        if (r0 + TG_RSTEP * j < rows_valid) {
          const float *sj = src + (size_t)(TG_RSTEP * j) * K;
          if (fast) {
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(sj));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(sj + 4));
            xh[j3][0] = v0.x; xh[j3][1] = v0.y; xh[j3][2] = v0.z; xh[j3][3] = v0.w;
            xh[j3][4] = v1.x; xh[j3][5] = v1.y; xh[j3][6] = v1.z; xh[j3][7] = v1.w;
          } else if (vec2) {   // even K (QM9_CC: 466): rows are 8-byte aligned -- four 8-byte loads instead of eight scalar ones
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
              float2 v2 = make_float2(0.f, 0.f);
              if (k + q < K) v2 = __ldg(reinterpret_cast<const float2 *>(sj + q));   // (K even, k + q even: the pair is whole)
              xh[j3][q] = v2.x; xh[j3][q + 1] = v2.y;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) xh[j3][q] = (k + q < K) ? __ldg(sj + q) : 0.f;
          }
        }
      }
    };
    auto store_half = [&](const float (&xh)[NH][8], uint8_t *st, int rows_valid, int h) {
#pragma unroll
      for (int j3 = 0; j3 < NH; ++j3) {
        const int j = h * NH + j3;
        if (r0 + TG_RSTEP * j < rows_valid) {
          uint4 hi, lo;
          tc::split8(xh[j3], hi, lo);
          *reinterpret_cast<uint4 *>(st + off0 + j * (TG_RSTEP * 128u)) = hi;
          *reinterpret_cast<uint4 *>(st + TG_HALF + off0 + j * (TG_RSTEP * 128u)) = lo;
        }
      }
    };
    float xa[NH][8], xb[NH][8];
    if (vb < NV) { load_half(xa, vb, kb, 0); load_half(xb, vb, kb, 1); }
    for (; vb < NV; gi += 2) {
      if (tg == 0) TG_STAMP(grp * 4 + 0, gi);
      const int rows_valid = ((B - vb * G < G) ? B - vb * G : G) * E;
      const int k = kb * TG_BK + c * 8;
      int vbn = vb, kbn = kb + 2;
      while (kbn >= nkb && vbn < NV) { kbn -= nkb; vbn += (int)gridDim.x; }
      if (tg == 0) TG_STAMP(grp * 4 + 1, gi);
      tc::mbar_wait(empty0 + 8 * s, ph ^ 1);
      if (tg == 0) TG_STAMP(grp * 4 + 2, gi);
      uint8_t *st = gen_base + (size_t)s * TG_STAGE;
      // projection rows: asynchronous copies of the prepared operand chunks (no registers, no wait until the arrive) ...
      const bool img_row = a.wimg != nullptr && r0 < PR0;
      if (img_row) {
        const int row = wp0 + r0;
        const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
        const uint8_t *src = a.wimg + ((size_t)(kb * PR0 + r0) * 8 + c) * 32;
        const uint32_t sdst = base + (uint32_t)s * TG_STAGE + off;
        tc::cp_async16(sdst, src, 16u);
        tc::cp_async16(sdst + TG_HALF, src + 16, 16u);
      }
      store_half(xa, st, rows_valid, 0);
      if (tg == 0 && grp == 0) TG_STAMP(4, gi);
      if (vbn < NV) load_half(xa, vbn, kbn, 0);
      store_half(xb, st, rows_valid, 1);
      if (tg == 0 && grp == 0) TG_STAMP(5, gi);
      if (vbn < NV) load_half(xb, vbn, kbn, 1);
      if (tg == 0 && grp == 0) TG_STAMP(6, gi);
      // ... or, without the image (and for rows past the first sweep), converted from the fp32 blob (zero padded to Kw)
      for (int rw = a.wimg ? r0 + TG_RSTEP : r0; rw < PR0; rw += TG_RSTEP) {
        const float *sw = Wp + (size_t)rw * Kw + k;
        float y[8];
        if (k + 8 <= Kw) {
          const float4 v0 = __ldg(reinterpret_cast<const float4 *>(sw));
          const float4 v1 = __ldg(reinterpret_cast<const float4 *>(sw + 4));
          y[0] = v0.x; y[1] = v0.y; y[2] = v0.z; y[3] = v0.w; y[4] = v1.x; y[5] = v1.y; y[6] = v1.z; y[7] = v1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) y[q] = (k + q < Kw) ? __ldg(sw + q) : 0.f;
        }
        const int row = wp0 + rw;
        const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
        uint4 hi, lo;
        tc::split8(y, hi, lo);
        *reinterpret_cast<uint4 *>(st + off) = hi;
        *reinterpret_cast<uint4 *>(st + TG_HALF + off) = lo;
      }
      if (rsum && r0 == 0) {   // the all-ones row (row sums of F as one more Gram column): exact in bf16, lo = 0
        const int row = rcol;
        const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = (k + 2 * q < K ? 0x3F80u : 0u) | (k + 2 * q + 1 < K ? 0x3F800000u : 0u);
        *reinterpret_cast<uint4 *>(st + off) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4 *>(st + TG_HALF + off) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (tg == 0 && grp == 0) TG_STAMP(7, gi);
      tc::cp_async_commit();
      tc::cp_async_wait<0>();
      tc::fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
      tc::mbar_arrive(full0 + 8 * s);
      if (tg == 0) TG_STAMP(grp * 4 + 3, gi);
      vb = vbn; kb = kbn;
      s += 2;
      if (s >= TG_STAGES) { s -= TG_STAGES; ph ^= 1u; }
    }
  } else if (warp == TG_PROD_WARPS + TG_EPI_WARPS) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::make_idesc_bf16(128, ncols, 0, 0);
    const uint32_t idesc1 = tc::make_idesc_bf16(128, N1 > 0 ? N1 : 16, 0, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);   // warp-uniform copy (uniform-register MMA operands)
    uint32_t it = 0, tile = 0;
    for (int vb = blockIdx.x; vb < NV; vb += gridDim.x, ++tile) {
      const uint32_t buf = dbuf ? (tile & 1u) : 0u, use = dbuf ? (tile >> 1) : tile;
      tc::mbar_wait(tempty0 + 8 * buf, (use & 1u) ^ 1u);   // the epilogue has drained this buffer's previous unit
      if (mtiles == 2) tc::mbar_wait(tempty1, (tile & 1u) ^ 1u);
      tc::tc_fence_after_sync();
      const uint32_t acc0 = tmem_u + buf * (uint32_t)ncols;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % TG_STAGES;
        const uint32_t ph = (it / TG_STAGES) & 1;
        if (lane == 0) TG_STAMP(8, it);
        tc::mbar_wait(full0 + 8 * s, ph);
        tc::tc_fence_after_sync();
        if (lane == 0) TG_STAMP(9, it);
        if (tc::elect_one()) {
          const uint32_t sb = base + (uint32_t)s * TG_STAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t ko = (uint32_t)k4 * 32u;   // 16 bf16 = 32 bytes inside the 128-byte swizzle span
            const uint64_t b_hi = tc::make_smem_desc(sb + ko, 0, 1024);
            const uint64_t b_lo = tc::make_smem_desc(sb + TG_HALF + ko, 0, 1024);
            {
              const uint64_t a_hi = tc::make_smem_desc(sb + ko, 0, 1024);
              const uint64_t a_lo = tc::make_smem_desc(sb + TG_HALF + ko, 0, 1024);
              tc::umma_bf16(acc0, a_hi, b_hi, idesc, (kb | k4) != 0);
              tc::umma_bf16(acc0, a_hi, b_lo, idesc, 1);
              tc::umma_bf16(acc0, a_lo, b_hi, idesc, 1);
            }
            if (mtiles == 2) {   // rows 128 .. against the operand rows c1 .. (8-row atoms of 1024 bytes)
              const uint64_t a_hi = tc::make_smem_desc(sb + 16384u + ko, 0, 1024);
              const uint64_t a_lo = tc::make_smem_desc(sb + TG_HALF + 16384u + ko, 0, 1024);
              const uint64_t b1_hi = tc::make_smem_desc(sb + (uint32_t)(c1 >> 3) * 1024u + ko, 0, 1024);
              const uint64_t b1_lo = tc::make_smem_desc(sb + TG_HALF + (uint32_t)(c1 >> 3) * 1024u + ko, 0, 1024);
              const uint32_t dcol = tmem_u + acc1;
              tc::umma_bf16(dcol, a_hi, b1_hi, idesc1, (kb | k4) != 0);
              tc::umma_bf16(dcol, a_hi, b1_lo, idesc1, 1);
              tc::umma_bf16(dcol, a_lo, b1_hi, idesc1, 1);
            }
          }
          tc::umma_commit(empty0 + 8 * s);             // stage may be refilled once these MMAs retire
          if (kb == nkb - 1) tc::umma_commit(tfull + 8 * buf);   // accumulators complete
        }
        __syncwarp();
        if (lane == 0) TG_STAMP(10, it);
      }
    }
  } else {
    // ===================== epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter alternate chunks) =====================
    const int q = warp & 3, half = (warp - TG_PROD_WARPS) >> 2;
    const int mask_diag = d.netf.use_hodge_mask;
    uint32_t tile = 0;
    const size_t ep = (size_t)P->Ep;
    for (int vb = blockIdx.x; vb < NV; vb += gridDim.x, ++tile) {
      const uint32_t buf = dbuf ? (tile & 1u) : 0u, use = dbuf ? (tile >> 1) : tile;
      tc::mbar_wait(tfull + 8 * buf, use & 1u);
      tc::tc_fence_after_sync();
      if (threadIdx.x == TG_PROD) TG_STAMP(12, tile * nkb);
      for (int mi = 0; mi < mtiles; ++mi) {
        const int mt = mtiles - 1 - mi;   // the second tile first: its single accumulator is what the next unit's MMAs wait for
        const int row = mt * 128 + q * 32 + lane;
        const int gs = G == 1 ? 0 : row / E, er = row - gs * E, b = vb * G + gs;   // sample of this row inside the unit, edge row
        const bool live = row < EG && b < B;
        // H is symmetric: lane `row` holds H[row][c0 .. c0+15]; it is stored as H[c0+j][row], so that for
        // every j the 32 lanes of the warp write 32 CONSECUTIVE floats (one coalesced 128-byte store)
        // instead of 32 rows 760 bytes apart.
        float *Hcol = a.H + (size_t)b * E * ep + er;
        float *Hrow = a.H + ((size_t)b * E + er) * ep;   // G == 1, first tile: the columns >= 128 are ALSO stored untransposed
        const bool both = !GROUPED && mtiles == 2 && mt == 0;
        float *Prow = a.P0 + ((size_t)b * E + er) * PR0;
        const int cb0 = gs * E, cb1 = cb0 + E;   // this sample's diagonal block of the stacked Gram
        // warp-uniform range of columns that hold a diagonal block of the warp's 32 rows (G = 1: all of [0, E))
        const int rlo = mt * 128 + q * 32, rhi = rlo + 31 < EG ? rlo + 31 : EG - 1;
        const int blo = G == 1 ? 0 : (rlo / E) * E, bhi = G == 1 ? E : (rhi / E) * E + E;
        const int cfirst = mt == 0 ? 0 : c1;
        const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + (mt == 0 ? buf * (uint32_t)ncols : acc1 - (uint32_t)c1);
        for (int c0 = cfirst + half * 16; c0 < ncols; c0 += 16 * (TG_EPI_WARPS / 4)) {
          if (!GROUPED && c0 + 16 <= E) {
            // a chunk of H columns only: 16 unconditional coalesced stores (the diagonal is patched first, by the one warp
            // whose rows the chunk crosses)
            float v[16];
            tc::tmem_ld16(tcol + (uint32_t)c0, v);
            if (live) {
              if (both && c0 >= 128) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(Hrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
              if (c0 < rlo + 32 && c0 + 16 > rlo) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (c0 + j == row) {
                    if (a.Dg) a.Dg[(size_t)b * E + er] = v[j];
                    if (mask_diag) v[j] = 0.f;
                  }
                }
              }
              float *hc = Hcol + (size_t)c0 * ep;
#pragma unroll
              for (int j = 0; j < 16; ++j) hc[(size_t)j * ep] = v[j];
            }
            continue;
          }
          if (!((c0 < bhi && c0 + 16 > blo) || c0 + 16 > wp0 || (rsum && c0 <= rcol && rcol < c0 + 16))) continue;   // neither a diagonal block nor projections / row sums
          float v[16];
          tc::tmem_ld16(tcol + (uint32_t)c0, v);
          if (live) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = c0 + j;
              if (col >= cb0 && col < cb1) {
                Hcol[(size_t)(col - cb0) * ep] = (mask_diag && col == row) ? 0.f : v[j];
                if (both && col >= 128) Hrow[col] = v[j];
                if (col == row && a.Dg) a.Dg[(size_t)b * E + er] = v[j];
              } else if (rsum && col == rcol) a.Rs[(size_t)b * E + er] = v[j];
              else if (col >= wp0 && col - wp0 < PR0) Prow[col - wp0] = v[j];
            }
          }
        }
        tc::tc_fence_before_sync();
        if (mt == 1) tc::mbar_arrive(tempty1);
        else tc::mbar_arrive(tempty0 + 8 * buf);
        if (threadIdx.x == TG_PROD) TG_STAMP(13 + mt, tile * nkb);
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TG_PROD_WARPS + TG_EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}

static inline int tc_gram_prepare() {
  static CcsdSmemAttr a0, a1;
  if (ccsd_ensure_smem(tc_gram_kernel<false>, TG_SMEM, a0)) return -1;
  return ccsd_ensure_smem(tc_gram_kernel<true>, TG_SMEM, a1);
}

static inline int tc_gram_launch(const DevPlan *dP, const DevPlan &hp, const float *r2, float *H, float *P0, float *Dg, float *Rs,
                                 void *stream, long long *trace = nullptr, const uint8_t *wimg = nullptr) {
  TcGramArgs a;
  a.r2 = r2; a.H = H; a.P0 = P0; a.Dg = Dg; a.Rs = Rs; a.trace = trace; a.wimg = wimg;
  const int nv = (hp.d.B + hp.gram_group - 1) / hp.gram_group;
  int grid = nv < 148 ? nv : 148;
  if (hp.gram_group > 1) tc_gram_kernel<true><<<grid, tg_threads(true), TG_SMEM, (cudaStream_t)stream>>>(dP, a);
  else tc_gram_kernel<false><<<grid, tg_threads(false), TG_SMEM, (cudaStream_t)stream>>>(dP, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccsd
