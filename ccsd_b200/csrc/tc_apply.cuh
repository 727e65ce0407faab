// tc_apply.cuh -- ScoreNetworkF's second contraction H . F on the 5th-gen tensor cores, fused with the
// per-entry network, the masks, the score scaling and every sampler update that touches the rank-2
// state (Langevin norms, Langevin correction, predictor, Philox noise).  The state is read from HBM
// ONCE and written ONCE per pass.
//
//   D[m = edge e, n = cell k] = sum_{e'} H[e, e'] * F[e', k]            (= (H F)[e, k])
//
//   A = H (E x E, E <= 192) lives in TENSOR MEMORY for the whole sample (tcgen05.mma A-from-TMEM form):
//       row e on a TMEM lane, element e' in 32-bit column e'/2 as a bf16 pair; two M tiles of 128 lanes
//       (edges 0..127 | 128..191) x {hi, lo} x 96 columns = 384 columns.  This frees the 144 KB of
//       shared memory a resident smem operand would need.
//   B = F tile: 32 cells x all E edge rows, MN-major (cells contiguous -- the state's own layout in
//       HBM), bf16 hi/lo in the canonical SWIZZLE_128B layout; two tiles share one 64-cell-wide
//       operand buffer (tile slot = 64-byte offset of the descriptor start address).
//   D: 128 lanes x 32 fp32 columns per M tile, double buffered in the remaining 128 TMEM columns.
//   bf16x3: D += Ahi.Bhi + Ahi.Blo + Alo.Bhi with fp32 accumulation (dropped term ~2^-18 relative).
//
// The SAME fp32 tile that feeds the operand conversion stays in a shared-memory staging ring
// (cp.async, 6 stages x 24 KB, 16-byte chunks XOR-swizzled by the row so that both the row-wise loader
// accesses and the column-wise epilogue accesses are conflict free).  The epilogue threads -- one per
// edge row, i.e. per TMEM lane -- read their row's 16 cells from the staging tile, combine them with the
// accumulator, Philox noise and masks, and write the NEW state back into the staging tile in place;
// the loader warps then copy the finished tile out with coalesced 16-byte stores.
//
// Small complexes: when E <= 96 several consecutive samples form one work group (G = min(8, 192 / E) samples,
// G * E edge rows): their rows are contiguous in the state, so the F tile is simply taller, and H becomes
// block diagonal in tensor memory.  Every per-row quantity (flags, Philox key, norm partial) follows the
// row's own sample.
//
// Warp roles (17 warps): 0-3 loaders (cp.async ring, bf16 split, copy-out), 4-15 epilogue (TMEM lane
// quarter q = warp % 4; role = (warp - 4) / 4: M tile 0 cells 0-15 | M tile 0 cells 16-31 | M tile 1, where
// the 16 edges of the quarter are processed by lane pairs (l, l + 16) taking 16 cells each, so every lane
// of every epilogue warp has work), 16-17 MMA issuers (one per M tile: the issue rate of one thread, not the
// tensor pipe, bounds a 72-instruction tile), 18 TMA warp.  When the state's row pitch allows it (K % 4 == 0)
// the fp32 tiles move by TMA: one elected thread issues a 2-D tensor load per tile into the staging ring
// (SWIZZLE_128B is exactly the ring's chunk-XOR layout; cells past K are zero-filled) and a tensor store of
// every finished tile (clipped at K), so the loader warps only convert operands.  Otherwise the loader
// warps move the tiles themselves with cp.async and coalesced stores.  The kernel is templated on the ScoreNetworkF entry
// path AND the pass mode, so the per-entry code has no mode branches.
#pragma once
#include "r2_kernels.cuh"
#include "tc_common.cuh"
#include <cuda.h>   // CUtensorMap

namespace ccsd {

#ifndef TA_LOAD_THREADS
#define TA_LOAD_THREADS 128
#endif
#ifndef TA_PREFETCH
#define TA_PREFETCH 2
#endif
#ifndef TA_FILL_NL
#define TA_FILL_NL 32   // loads in flight per column group of the H refill (probe: 8 -> +25 % refill time: it is issue bound, not latency bound)
#endif
constexpr int TA_LOAD = TA_LOAD_THREADS;
constexpr int TA_RSTEP = TA_LOAD / 8;              // rows covered by one sweep of the loader threads
constexpr int TA_EPI = 384;
constexpr int TA_THREADS = TA_LOAD + TA_EPI + 96;   // + one MMA-issuing warp per M tile + the TMA warp
constexpr int TA_TN = 32;                         // cells per tile
constexpr int TA_NE = 192;                        // padded edge count
constexpr int TA_NS = 6;                          // staging stages (the loaders only block on a tile 4 behind)
constexpr int TA_PD = TA_PREFETCH;                          // cp.async prefetch distance (tiles)
constexpr uint32_t TA_OPHALF = TA_NE * 128u;      // 24576: hi (or lo) operand rows, 64 cells (2 tile slots) x bf16
constexpr uint32_t TA_OPER = 2u * TA_OPHALF;      // 49152
constexpr uint32_t TA_STAGE = TA_NE * 128u;       // 24576: 192 rows x 32 fp32
constexpr uint32_t TA_BARS = TA_OPER + TA_NS * TA_STAGE;   // 147456
constexpr int TA_GMAX = 8;                        // samples per work group (E <= 24)
constexpr uint32_t TA_FCS = TA_BARS + 256;        // [NS][GMAX][32] cell flags
constexpr uint32_t TA_RED = TA_FCS + TA_NS * TA_GMAX * 32 * 4;   // [2 cell halves][192 rows][2] norm partials
constexpr int TA_MMAW = (TA_LOAD + TA_EPI) / 32;     // index of the MMA warp
constexpr uint32_t TA_FW = TA_RED + 2 * TA_NE * 2 * 4;   // staged ScoreNetworkF weights (FMODE 2)
constexpr int TA_FW_FLOATS = 2048;
constexpr size_t TA_SMEM = (size_t)TA_FW + TA_FW_FLOATS * 4 + 1024 /*alignment slack*/;
constexpr uint32_t TA_COL_D = 384;                // first accumulator column

static inline int tc_apply_supported(int E, int K) { return E >= 8 && E <= TA_NE && K >= 8; }

#ifdef TC_APPLY_KERNEL_TU
// Per-pass constants of the entry update, folded so that the affine ScoreNetworkF path is a handful of FMAs:
//   score s = sc * o,  o = m * (a0 f + a1 hf + a2)  (FMODE 1)  ->  s = m * (k0 f + k1 hf + k2)
struct R2Fold {
  float k0, k1, k2;    // sc * aff (FMODE 1)
  float sc;            // score scale
  float cs, cn;        // Langevin step sizes (CORR)
  float pa, pb, pc;    // predictor
};

// one entry (edge e, cell k): network -> score -> mode-specific value.  Returns the value that replaces
// the entry in the staging tile (raw output, scaled score or new state).
template <int FMODE, int MODE>
__device__ __forceinline__ float r2_entry(const R2Epi &c, const R2Fold &w, float f, float hf, float zraw, float m,
                                          float &s2, float &z2, float &mu_out, float o_pre) {
  const DevPlan *P = c.P;
  float s;
  if (FMODE == 1) {
    s = m * (w.k0 * f + w.k1 * hf + w.k2);     // MODE_EVAL: the fold uses sc = 1
  } else {
    float o;
    if (FMODE == 3) o = o_pre;                 // network evaluated four entries at a time by the caller
    else if (FMODE == 2) o = netf_entry_w8(P->d.netf, c.fw, c.f_nlin, f, hf, m);
    else if (FMODE == 4) o = netf_entry_w8f2(P->d.netf, c.fw, c.f_nlin, f, hf, m);
    else o = netf_entry(P->d.netf, P->W, f, hf, m);
    s = w.sc * o;
  }
  if (MODE == MODE_EVAL) return s;
  const float z = zraw * m;
  if (MODE == MODE_SCORE || MODE == MODE_NORM) {
    s2 += s * s;
    z2 += z * z;
    return s;
  }
  if (MODE == MODE_CORR) return f + w.cs * s + w.cn * z;   // Langevin (solver.py:784-785)
  const float mu = w.pa * f + w.pb * s;                    // predictor (solver.py:283-300, 433-450)
  mu_out = mu;
  return mu + w.pc * z;
}

// debug timeline (ccsd_debug_apply_trace): stamp `slot` of tile g, CTA 0 only
#define TA_STAMP(slot_, g_) do { if (a.trace && blockIdx.x == 0 && (g_) < 512) a.trace[(size_t)(g_) * 16 + (slot_)] = clock64(); } while (0)

// H of a work group -> the TMEM A operand (called once per group by every epilogue warp; inlined: as a real call its
// register save / restore cost the main kernel 4 %).
// The A operand is 2 M tiles x 6 groups of 32 e' columns (16 TMEM columns for hi, 16 for lo) per lane quarter.  The three
// warps of a quarter (roles 0-2) take every third (M tile, group) pair, whichever tile their entries belong to:
// M tile 0 holds edge row 32 q + lane on lane 32 q + lane, M tile 1 holds row 128 + 16 q + lane on the lanes < 16 of
// the quarter (zero rows above).  Block-diagonal operand: H of the row's sample in columns [sg E, sg E + E).
__device__ __forceinline__ void ta_fill_h(const float *__restrict__ H, int b0, int gsz, int E, int Ep, int EB, int mtiles, int role, int q,
                                       int lane, uint32_t tmem_lane) {
  const int npair = mtiles * 6;
  for (int idx = role; idx < npair; idx += 3) {
    const int fm = idx / 6, c16 = idx - fm * 6;
    int erf = fm == 0 ? q * 32 + lane : (lane < 16 ? 128 + q * 16 + lane : -1);
    if (erf >= EB) erf = -1;
    const int sgf = erf >= 0 ? erf / E : 0, ef = erf >= 0 ? erf - sgf * E : 0;
    const bool frow = erf >= 0 && sgf < gsz;
    // H is symmetric: element (e, k) is read as (k, e), so that the lanes of a warp (consecutive rows e) read consecutive
    // addresses -- one 128-byte wavefront per load instead of 32 (a row-wise float4 per lane made the refill at every
    // sample boundary cost 20-30 k cycles: 16 % of the pass)
    const float *Hcolp = H + (size_t)(b0 + (frow ? sgf : 0)) * E * Ep + ef;
    const int c_lo = sgf * E;
    const int kb = c16 * 32;                       // first e' column of this 32-element (16-column) group
    uint32_t hw[16], lw[16];
    if (!frow || kb + 32 <= c_lo || kb >= c_lo + E) {
      // outside the row's own diagonal block (or no row): zeros, no loads, no split.  Per-lane branch: both sides end in
      // the same two warp-wide stores below.
#pragma unroll
      for (int j = 0; j < 16; ++j) { hw[j] = 0u; lw[j] = 0u; }
    } else {
      const bool inner = kb >= c_lo && kb + 32 <= c_lo + E;   // the whole group lies inside the block: no per-element bounds
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += TA_FILL_NL) {
        float vv[TA_FILL_NL];
        if (inner) {
          const float *hp = Hcolp + (size_t)(kb + j0 - c_lo) * Ep;
#pragma unroll
          for (int j = 0; j < TA_FILL_NL; ++j, hp += Ep) vv[j] = __ldg(hp);
        } else {
#pragma unroll
          for (int j = 0; j < TA_FILL_NL; ++j) {
            const int k = kb + j0 + j - c_lo;        // column inside the sample's own H
            vv[j] = (k >= 0 && k < E) ? __ldg(Hcolp + (size_t)k * Ep) : 0.f;
          }
        }
#pragma unroll
        for (int j4 = 0; j4 < TA_FILL_NL / 4; ++j4) {
          uint2 hi, lo;
          tc::split4(make_float4(vv[4 * j4], vv[4 * j4 + 1], vv[4 * j4 + 2], vv[4 * j4 + 3]), hi, lo);
          hw[(j0 >> 1) + 2 * j4] = hi.x; hw[(j0 >> 1) + 2 * j4 + 1] = hi.y;
          lw[(j0 >> 1) + 2 * j4] = lo.x; lw[(j0 >> 1) + 2 * j4 + 1] = lo.y;
        }
      }
    }
    const uint32_t col = (uint32_t)(fm * 192 + c16 * 16);
    tc::tmem_st16(tmem_lane + col, hw);
    tc::tmem_st16(tmem_lane + col + 96u, lw);
  }
}

template <int FMODE, int MODE>
__global__ void __launch_bounds__(TA_THREADS, 1) tc_apply_kernel(const DevPlan *__restrict__ P, ApplyArgs a,
                                                                 const __grid_constant__ CUtensorMap tm_in,
                                                                 const __grid_constant__ CUtensorMap tm_out, int use_tma) {
  extern __shared__ uint8_t ta_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, E = d.E, K = d.K, B = d.B, Ep = P->Ep;
  const int ntile = (K + TA_TN - 1) / TA_TN;
  const int G = P->ap_group;                    // samples per work group
  const int EB = G * E;                         // edge rows of a full group (<= 192)
  const int ngroups = (B + G - 1) / G;
  const int nk = (EB + 15) >> 4;                // 16-wide k steps over e'
  const int mtiles = EB > 128 ? 2 : 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t raw = tc::smem_u32(ta_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = ta_smem_raw + (base - raw);
  const uint32_t sOp = base, sStage = base + TA_OPER, bars = base + TA_BARS;
  // barriers (8 bytes each): full[2] op_empty[2] t_full[2] d_empty[2] epi_done[NS] stage_full[NS] h_ready, then
  // the TMEM slot.  Every barrier's next phase is gated by all of its waiters having passed the previous one
  // (a parity wait must never fall two phases behind): `full` (operand slot, 2-deep) is waited on by the MMA
  // warp only; the epilogue learns that a STAGING tile is complete from stage_full (4-deep), whose next
  // phase needs the stage to have been copied out, i.e. every epilogue thread to be done with it.
  const uint32_t full = bars, op_empty = bars + 16, t_full = bars + 32, d_empty = bars + 48, epi_done = bars + 64,
                 stage_full = epi_done + 8 * TA_NS, tma_full = stage_full + 8 * TA_NS, h_ready = tma_full + 8 * TA_NS,
                 tslot = h_ready + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TA_BARS + 64 + 24 * TA_NS + 8);
  float *fcs = reinterpret_cast<float *>(gen + TA_FCS);
  float *red = reinterpret_cast<float *>(gen + TA_RED);
  float *fw = reinterpret_cast<float *>(gen + TA_FW);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(full + 8 * s, TA_LOAD);
      tc::mbar_init(op_empty + 8 * s, mtiles);   // one tcgen05.commit per MMA-issuing warp
      tc::mbar_init(t_full + 8 * s, mtiles);
      tc::mbar_init(d_empty + 8 * s, TA_EPI);
    }
    for (int s = 0; s < TA_NS; ++s) {
      tc::mbar_init(epi_done + 8 * s, TA_EPI);
      tc::mbar_init(stage_full + 8 * s, TA_LOAD);
      tc::mbar_init(tma_full + 8 * s, 1);
    }
    tc::mbar_init(h_ready, TA_EPI);
    tc::mbar_fence_init();
  }
  if (warp == TA_MMAW) tc::tmem_alloc(tslot, 512);
  if (FMODE == 2 || FMODE == 3) netf_stage_w8(d.netf, P->W, fw);
  if (FMODE == 4) netf_stage_w8f2(d.netf, P->W, fw);
  // operand rows that no edge fills stay zero for the whole kernel (their A columns are zero too, but
  // 0 * NaN from uninitialised shared memory would poison the accumulator)
  for (uint32_t o = threadIdx.x * 16u; o < TA_OPER; o += TA_THREADS * 16u)
    *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const int nmine = (ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // groups of this CTA
  const int ntot = nmine * ntile;

  if (warp < TA_LOAD / 32) {
    // ===================== loaders =====================
    const bool vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.r2) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
    const bool vec2 = ((K & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.r2) & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 7) == 0);
    const int cu = threadIdx.x & 7, r0 = threadIdx.x >> 3;   // 16-byte chunk (4 cells) of rows r0 + RSTEP j
    const bool writes = MODE != MODE_NORM;
    // rows advance by a multiple of 8, so the XOR swizzle term (r & 7) is a per-thread constant and every
    // address below is base + j * stride
    constexpr int NJ = TA_NE / TA_RSTEP;
    constexpr uint32_t JSTR = TA_RSTEP * 128u;
    const int nj = (EB - r0 + TA_RSTEP - 1) / TA_RSTEP;         // rows of this thread in a full group
    const uint32_t soff = (uint32_t)r0 * 128u + (uint32_t)((cu ^ (r0 & 7)) << 4);
    // tile g -> first sample b of its group, rows Eg of the group that exist (the last group may be short), first cell
    auto tile_of = [&](int g, int &b, int &Eg, int &k0) {
      const int si = g / ntile;
      b = ((int)blockIdx.x + si * (int)gridDim.x) * G;
      Eg = (B - b < G ? B - b : G) * E;
      k0 = (g - si * ntile) * TA_TN;
    };
    auto issue = [&](int g) {
      int b, Eg, k0;
      tile_of(g, b, Eg, k0);
      const float *Fb = a.r2 + (size_t)b * E * K;
      const uint32_t st = sStage + (uint32_t)(g % TA_NS) * TA_STAGE + soff;
      const int k = k0 + 4 * cu;
      const float *src = Fb + (size_t)r0 * K + k;
      const size_t gstr = (size_t)TA_RSTEP * K;
      // rows past the group's last sample (short last group) are zero-filled (src-size 0)
      if (vec) {
        const int nb = k + 4 <= K ? 16 : (k < K ? (K - k) * 4 : 0);
        if (nb == 0) src = Fb;
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
          const bool ok = nb && r0 + j * TA_RSTEP < Eg;
          tc::cp_async16(st + (uint32_t)j * JSTR, ok ? (const void *)(src + j * gstr) : (const void *)a.r2, ok ? (uint32_t)nb : 0u);
        }
      } else {
        // Row pitch not 16-byte aligned (QM9_CC: K = 466, ENZYMES_small_CC: K = 715).  The loader warps run at a fraction of an
        // issue slot beside the epilogue warps, so what counts is instructions per row: the row / cell bounds are hoisted
        // (njv valid rows, nv valid cells of this thread's chunk), the pointers advance by constant strides, and the copies
        // are 8-byte ones when K is even.
        int njv = Eg > r0 ? (Eg - r0 + TA_RSTEP - 1) / TA_RSTEP : 0;
        if (njv > nj) njv = nj;
        const int nv = k + 4 <= K ? 4 : (k < K ? K - k : 0);
        const float *sp = src;
        uint32_t sa = st;
        if (vec2 && nv == 4) {
#pragma unroll 4
          for (int j = 0; j < njv; ++j, sa += JSTR, sp += gstr) {
            tc::cp_async8(sa, sp, 8u);
            tc::cp_async8(sa + 8u, sp + 2, 8u);
          }
          for (int j = njv; j < nj; ++j, sa += JSTR)   // rows past the group's last sample (short last group): zero fill
#pragma unroll
            for (int i = 0; i < 4; ++i) tc::cp_async4(sa + 4u * i, (const void *)a.r2, 0u);
        } else {
          for (int j = 0; j < nj; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const bool ok = k + i < K && r0 + j * TA_RSTEP < Eg;
              tc::cp_async4(st + (uint32_t)j * JSTR + 4u * i, ok ? (const void *)(src + j * gstr + i) : (const void *)a.r2, ok ? 4u : 0u);
            }
        }
      }
    };
    auto copy_out = [&](int g) {
      tc::mbar_wait_relaxed(epi_done + 8 * (g % TA_NS), (uint32_t)(g / TA_NS) & 1u);
      if (!writes) return;
      int b, Eg, k0;
      tile_of(g, b, Eg, k0);
      const uint8_t *st = gen + TA_OPER + (size_t)(g % TA_NS) * TA_STAGE + soff;
      const int k = k0 + 4 * cu;
      if (k >= K) return;
      const int njw = Eg > r0 ? (Eg - r0 + TA_RSTEP - 1) / TA_RSTEP : 0;   // rows of this thread that exist
      float *dst = a.out + (size_t)b * E * K + (size_t)r0 * K + k;
      const size_t gstr = (size_t)TA_RSTEP * K;
      if (vec) {
        for (int j0 = 0; j0 < njw; j0 += 4) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + u < njw) v[u] = *reinterpret_cast<const float4 *>(st + (size_t)(j0 + u) * JSTR);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + u < njw) __stcs(reinterpret_cast<float4 *>(dst + (j0 + u) * gstr), v[u]);
        }
      } else {
        const int nv = k + 4 <= K ? 4 : K - k;   // (k < K here)
        const uint8_t *sp = st;
        float *dp = dst;
        if (vec2 && nv == 4) {
#pragma unroll 4
          for (int j = 0; j < njw; ++j, sp += JSTR, dp += gstr) {
            const float4 v = *reinterpret_cast<const float4 *>(sp);
            __stcs(reinterpret_cast<float2 *>(dp), make_float2(v.x, v.y));
            __stcs(reinterpret_cast<float2 *>(dp + 2), make_float2(v.z, v.w));
          }
        } else {
          for (int j = 0; j < njw; ++j) {
            const float4 v = *reinterpret_cast<const float4 *>(st + (size_t)j * JSTR);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (k + i < K) __stcs(dst + j * gstr + i, vv[i]);
          }
        }
      }
    };
    (void)NJ;
    if (!use_tma)
      for (int t = 0; t < TA_PD; ++t) {
        if (t < ntot) issue(t);
        tc::cp_async_commit();
      }

    for (int g = 0; g < ntot; ++g) {
      if (use_tma) {
        if (threadIdx.x == 0) TA_STAMP(0, g);
        tc::mbar_wait(tma_full + 8 * (g % TA_NS), (uint32_t)(g / TA_NS) & 1u);   // the tensor load of tile g has landed
      } else {
        if (g + TA_PD < ntot) {
          if (g + TA_PD - TA_NS >= 0) copy_out(g + TA_PD - TA_NS);   // frees the stage tile g + PD lands in
          issue(g + TA_PD);
        }
        tc::cp_async_commit();
        if (threadIdx.x == 0) TA_STAMP(0, g);
        tc::cp_async_wait<TA_PD>();                                   // this thread's chunks of tile g have landed
      }
      if (threadIdx.x == 0) TA_STAMP(1, g);
      int b, Eg, k0;
      tile_of(g, b, Eg, k0);
      // (the stage is this tile's since its load landed: the cell flags go first, so that their two dependent global loads
      // overlap the wait for the operand slot instead of sitting between the conversion and the arrive)
      for (int t = threadIdx.x; t < G * TA_TN; t += TA_LOAD) {   // cell flags of every sample of the group
        const int sg = t >> 5, k = k0 + (t & 31);
        const bool on = b + sg < B && k < K && !(P->cell_mask[k] & a.zmask[b + sg]);
        fcs[((g % TA_NS) * TA_GMAX + sg) * TA_TN + (t & 31)] = on ? 1.f : 0.f;
      }
      if (g >= 2) tc::mbar_wait(op_empty + 8 * (g & 1), (uint32_t)(((g >> 1) & 1) ^ 1));
      if (threadIdx.x == 0) TA_STAMP(2, g);
      const uint8_t *st = gen + TA_OPER + (size_t)(g % TA_NS) * TA_STAGE + soff;
      uint8_t *op = gen + (uint32_t)r0 * 128u + ((((uint32_t)(g & 1) * 4u + (uint32_t)(cu >> 1)) ^ (uint32_t)(r0 & 7)) << 4) +
                    (uint32_t)(cu & 1) * 8u;
      for (int j0 = 0; j0 < nj; j0 += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < nj) v[u] = *reinterpret_cast<const float4 *>(st + (size_t)(j0 + u) * JSTR);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < nj) {
            uint2 hi, lo;
            tc::split4(v[u], hi, lo);
            *reinterpret_cast<uint2 *>(op + (size_t)(j0 + u) * JSTR) = hi;
            *reinterpret_cast<uint2 *>(op + TA_OPHALF + (size_t)(j0 + u) * JSTR) = lo;
          }
      }
      tc::fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
      tc::mbar_arrive(full + 8 * (g & 1));
      tc::mbar_arrive(stage_full + 8 * (g % TA_NS));
      if (threadIdx.x == 0) TA_STAMP(3, g);
    }
    // drain: tiles [ntot - NS, ntot) are still in the ring (the loop copied out tile g + PD - NS)
    if (!use_tma)
      for (int g = ntot > TA_NS ? ntot - TA_NS : 0; g < ntot; ++g) copy_out(g);
  } else if (warp == TA_MMAW + 2) {
    // ===================== TMA warp: tensor loads into / tensor stores out of the staging ring =====================
    if (use_tma && lane == 0) {
      const bool writes = MODE != MODE_NORM;
      const uint32_t tile_bytes = (uint32_t)EB * 128u;   // the box is EB rows; rows past the tensor are zero-filled / clipped
      auto coords = [&](int g, int &col, int &row) {
        const int si = g / ntile;
        row = ((int)blockIdx.x + si * (int)gridDim.x) * EB;
        col = (g - si * ntile) * TA_TN;
      };
      auto store = [&](int g) {   // tile g is final in its stage once every epilogue thread arrived (after a proxy fence)
        tc::mbar_wait(epi_done + 8 * (g % TA_NS), (uint32_t)(g / TA_NS) & 1u);
        if (writes) {
          int col, row;
          coords(g, col, row);
          tc::tma_store_2d(&tm_out, col, row, sStage + (uint32_t)(g % TA_NS) * TA_STAGE);
        }
        tc::bulk_commit();
      };
      for (int g = 0; g < ntot; ++g) {
        // Stage g % NS held tile g - NS.  Its store is issued one iteration EARLY (with tile g - NS + 1 ... see
        // below), so that by now only its shared-memory read has to have finished, not its issue.
        if (g >= TA_NS - 1) store(g - (TA_NS - 1));
        if (g >= TA_NS) tc::bulk_wait_read<1>();   // every store but the newest has read its stage: tile g - NS is out
        int col, row;
        coords(g, col, row);
        tc::mbar_arrive_expect_tx(tma_full + 8 * (g % TA_NS), tile_bytes);
        tc::tma_load_2d(sStage + (uint32_t)(g % TA_NS) * TA_STAGE, &tm_in, col, row, tma_full + 8 * (g % TA_NS));
      }
      for (int g = ntot > TA_NS - 1 ? ntot - (TA_NS - 1) : 0; g < ntot; ++g) store(g);
      tc::bulk_wait<0>();   // the stores must have completed before the CTA exits
    }
  } else if (warp >= TA_MMAW) {
    // ===================== MMA issuers: warp TA_MMAW + mt drives M tile mt =====================
    const int mt = warp - TA_MMAW;   // 0 or 1 (the TMA warp, TA_MMAW + 2, is handled above)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);   // warp-uniform copy of the TMEM base
    const uint32_t idesc = tc::make_idesc_bf16(128, TA_TN, /*A from TMEM*/ 0, /*B MN-major*/ 1);
    if (mt < mtiles) {
      for (int g = 0; g < ntot; ++g) {
        const int slot = g & 1;
        const int si = g / ntile;
        if (g - si * ntile == 0) tc::mbar_wait(h_ready, (uint32_t)si & 1u);
        tc::mbar_wait(full + 8 * slot, (uint32_t)(g >> 1) & 1u);
        tc::mbar_wait(d_empty + 8 * slot, (uint32_t)(((g >> 1) & 1) ^ 1));
        tc::tc_fence_after_sync();
        if (lane == 0 && mt == 0) TA_STAMP(4, g);
        if (tc::elect_one()) {
          // descriptors advance by 16 e' rows (2048 bytes = 128 descriptor units) per k step
          uint64_t b_hi = tc::make_smem_desc(sOp + (uint32_t)slot * 64u, 8192, 1024);
          uint64_t b_lo = tc::make_smem_desc(sOp + TA_OPHALF + (uint32_t)slot * 64u, 8192, 1024);
          const uint32_t dcol = tmem_u + TA_COL_D + (uint32_t)(slot * 64 + mt * 32);
          uint32_t a_hi = tmem_u + (uint32_t)(mt * 192);
#pragma unroll 1
          for (int k4 = 0; k4 < nk; ++k4) {
            tc::umma_bf16_ts_coll<1, 0>(dcol, a_hi, b_hi, idesc, k4 != 0);   // A_hi kept in the collector ...
            tc::umma_bf16_ts_coll<0, 1>(dcol, a_hi, b_lo, idesc, 1);         // ... and reused
            tc::umma_bf16_ts(dcol, a_hi + 96u, b_hi, idesc, 1);
            a_hi += 8u; b_hi += 128u; b_lo += 128u;
          }
          tc::umma_commit(op_empty + 8 * slot);   // operand slot may be refilled once these MMAs retire
          tc::umma_commit(t_full + 8 * slot);     // accumulators complete
          if (mt == 0) TA_STAMP(5, g);   // (stamped by the elected lane)
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 4-15) =====================
    const int ew = warp - TA_LOAD / 32;   // 0..11
    const int q = ew & 3;                // TMEM lane quarter (= warp % 4)
    const int role = ew >> 2;            // 0: M tile 0, cells 0-15   1: M tile 0, cells 16-31   2: M tile 1
    const int mt = role == 2 ? 1 : 0;
    const int et = threadIdx.x - TA_LOAD;
    const int Kg = P->Kp >> 2;
    // The edge row whose ENTRIES this thread processes, and its 16 cells of every tile.  M tile 0 holds
    // edges 0..127 on lanes 0..127.  M tile 1 holds edge 128 + 16 q + l on lane 32 q + l for l < 16 (every
    // lane quarter carries the same load); there lane l >= 16 processes cells 16-31 of lane (l - 16)'s row.
    int er, chalf;   // edge ROW of the group (sample-in-group * E + edge), cell half
    if (role < 2) { er = q * 32 + lane; chalf = role; }
    else { er = 128 + q * 16 + (lane & 15); chalf = lane >> 4; }
    if (er >= EB) er = -1;
    const int sg = er >= 0 ? er / E : 0;          // sample within the group
    const int e = er >= 0 ? er - sg * E : -1;     // edge of that sample
    // the A-operand row (TMEM lane 32 q + lane) this thread FILLS with H: the same row, except that the
    // upper lanes of M tile 1 hold no edge (zero rows)
    const bool fills = er >= 0 && !(role == 2 && lane >= 16);
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int ei = e >= 0 ? P->edge_ij[2 * e] : 0, ej = e >= 0 ? P->edge_ij[2 * e + 1] : 0;
    R2Fold w;
    {
      ccsd_objcoef_t co;
      memset(&co, 0, sizeof co);
      co.score_scale = 1.f;
      if (MODE != MODE_EVAL) co = P->sched[nz_step(a.nz) * 3 + 2];
      w.sc = co.score_scale;
      w.k0 = w.sc * d.netf.aff[0]; w.k1 = w.sc * d.netf.aff[1]; w.k2 = w.sc * d.netf.aff[2];
      w.cs = 0.f; w.cn = 0.f;
      if (MODE == MODE_CORR) { w.cs = a.coef[4]; w.cn = a.coef[5]; }
      w.pa = co.pa; w.pb = co.pb; w.pc = co.pc;
    }
    const uint32_t did = draw_id(2, nz_step(a.nz), a.slot);
    for (int si = 0; si < nmine; ++si) {
      const int b0 = ((int)blockIdx.x + si * (int)gridDim.x) * G;   // first sample of the group
      const int gsz = B - b0 < G ? B - b0 : G;
      const int b = b0 + sg;                                         // this thread's sample
      const bool live = er >= 0 && sg < gsz;
      const float *fl = a.flags + (size_t)(live ? b : b0) * N;
      R2Epi c;
      c.P = P; c.fl = fl; c.fw = fw; c.zm = 0ull;
      c.gs = (unsigned long long)(a.nz.sample_offset + b);
      c.b = b; c.E = E; c.K = K; c.Kg = Kg; c.f_nlin = P->f_nlin;
      const float fe = live ? fl[ei] * fl[ej] : 0.f;
      const bool side = MODE == MODE_PRED && (a.write_mean || (a.traj != nullptr && b == 0));
      // ---- H of this sample -> TMEM (A operand).  Every MMA of the previous sample has completed: this
      // warp waited on t_full of its last tile. ----
      if (lane == 0 && ew == 0) TA_STAMP(14, si * ntile);
      {
        ta_fill_h(a.H, b0, gsz, E, Ep, EB, mtiles, role, q, lane, tmem + lane_base);
        tc::tmem_st_wait();
        if (lane == 0 && ew == 0) TA_STAMP(13, si * ntile);
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(h_ready);
      if (lane == 0 && ew == 0) TA_STAMP(15, si * ntile);

      float s2 = 0.f, z2 = 0.f;
      const float *Nb = a.noise ? a.noise + (size_t)(live ? b : b0) * E * K : nullptr;
      for (int ct = 0; ct < ntile; ++ct) {
        const int g = si * ntile + ct;
        const int slot = g & 1, stg = g % TA_NS;
        if (ct == (ntile > 3 ? ntile - 3 : 0) && si + 1 < nmine && fills && mt < mtiles) {
          // the next group's H row of this thread -> L2, so the refill at the group boundary (which drains the MMA pipeline)
          // does not wait for HBM three times
          const int bn = b0 + (int)gridDim.x * G + sg;
          if (bn < B) {
            const char *hp = reinterpret_cast<const char *>(a.H + ((size_t)bn * E + e) * Ep);
            for (int o = 0; o < E * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + o));
          }
        }
        tc::mbar_wait_relaxed(stage_full + 8 * stg, (uint32_t)(g / TA_NS) & 1u);   // staging tile + cell flags visible
        if (lane == 0 && ew == 0) TA_STAMP(6, g);
        tc::mbar_wait_relaxed(t_full + 8 * slot, (uint32_t)(g >> 1) & 1u);         // accumulators complete
        tc::tc_fence_after_sync();
        if (lane == 0 && ew == 0) TA_STAMP(7, g);
        float v[16];
        const uint32_t dcol = tmem + lane_base + TA_COL_D + (uint32_t)(slot * 64 + mt * 32);
        if (role < 2) {
          tc::tmem_ld16(dcol + (uint32_t)(role * 16), v);
        } else {
          uint32_t va[16], vb[16];
          tc::tmem_ld16_nowait(dcol, va);
          tc::tmem_ld16_nowait(dcol + 16u, vb);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t t = __shfl_sync(0xffffffffu, vb[j], lane & 15);
            v[j] = __uint_as_float(lane < 16 ? va[j] : t);
          }
        }
        tc::tc_fence_before_sync();
        tc::mbar_arrive(d_empty + 8 * slot);                          // accumulator slot may be overwritten
        if (lane == 0 && ew == 0) TA_STAMP(8, g);
        if (live) {
          uint8_t *row = gen + TA_OPER + (size_t)stg * TA_STAGE + (size_t)er * 128;
          const float *fcp = fcs + (stg * TA_GMAX + sg) * TA_TN + chalf * 16;
          const int kbase = ct * TA_TN + chalf * 16;
          const bool inside = kbase + 16 <= K;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int k = kbase + 4 * c4;
            float4 *cell = reinterpret_cast<float4 *>(row + (((chalf * 4 + c4) ^ (er & 7)) << 4));
            const float4 fc4 = *reinterpret_cast<const float4 *>(fcp + 4 * c4);   // 0 for cells >= K
            // Four cells that are all masked for this sample (or a masked edge row): the state there is zero and stays
            // zero, the score and the masked noise are zero -- no network, no Philox, no store.  The test is uniform
            // over the lanes that share a sample (every lane of the warp when one sample fills the group).
            if (!(MODE == MODE_PRED && side) && (fe == 0.f || (fc4.x == 0.f && fc4.y == 0.f && fc4.z == 0.f && fc4.w == 0.f))) {
              if (MODE == MODE_EVAL || MODE == MODE_SCORE) *cell = make_float4(0.f, 0.f, 0.f, 0.f);   // out != in: the caller's input may be unmasked
              continue;
            }
            const float4 f4 = *cell;
            const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
            const float mv[4] = {fe * fc4.x, fe * fc4.y, fe * fc4.z, fe * fc4.w};
            float z4[4] = {0.f, 0.f, 0.f, 0.f};
            if (MODE != MODE_EVAL) {
              if (Nb) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (k + i < K) z4[i] = Nb[(size_t)e * K + k + i];
              } else {
                normal4(a.nz.seed, c.gs, did, (uint32_t)(e * Kg + (k >> 2)), z4);
              }
            }
            float o4[4], mu4[4], pre[4] = {0.f, 0.f, 0.f, 0.f};
            if (FMODE == 3) netf_entry_w4x4(d.netf, fw, c.f_nlin, fv, v + 4 * c4, mv, pre);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = r2_entry<FMODE, MODE>(c, w, fv[i], v[4 * c4 + i], z4[i], mv[i], s2, z2, mu4[i], pre[i]);
            if (MODE != MODE_NORM) *cell = make_float4(o4[0], o4[1], o4[2], o4[3]);
            if (MODE == MODE_PRED && side) {
              const bool tr = a.traj && b == 0;
              if (inside && !(K & 3)) {   // whole 16-byte groups (rows are 16-byte aligned when K % 4 == 0)
                const float4 m4 = make_float4(mu4[0], mu4[1], mu4[2], mu4[3]);
                if (a.write_mean) *reinterpret_cast<float4 *>(a.mean + ((size_t)b * E + e) * K + k) = m4;
                if (tr) *reinterpret_cast<float4 *>(a.traj + (size_t)e * K + k) = a.denoise ? m4 : make_float4(o4[0], o4[1], o4[2], o4[3]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (inside || k + i < K) {
                    if (a.write_mean) a.mean[((size_t)b * E + e) * K + k + i] = mu4[i];
                    if (tr) a.traj[(size_t)e * K + k + i] = a.denoise ? mu4[i] : o4[i];
                  }
              }
            }
          }
        }
        if (MODE != MODE_NORM) tc::fence_proxy_async_smem();          // generic writes -> visible to the TMA store
        tc::mbar_arrive(epi_done + 8 * stg);                          // (release) tile may be copied out
        if (lane == 0 && ew == 0) TA_STAMP(9, g);
      }
      if (MODE == MODE_SCORE || MODE == MODE_NORM) {
        // per-sample squared norms: every (cell half, row) partial goes to shared memory, then one thread per
        // sample of the group sums its rows in a fixed order (bit-reproducible); named barrier 1, 384 threads
        if (er >= 0) { red[(chalf * TA_NE + er) * 2] = live ? s2 : 0.f; red[(chalf * TA_NE + er) * 2 + 1] = live ? z2 : 0.f; }
        asm volatile("bar.sync 1, 384;" ::: "memory");
        if (et < gsz) {
          float ts = 0.f, tz = 0.f;
          for (int h = 0; h < 2; ++h)
            for (int r = et * E; r < et * E + E; ++r) { ts += red[(h * TA_NE + r) * 2]; tz += red[(h * TA_NE + r) * 2 + 1]; }
          float *np = a.norm_part + ((size_t)(2 * d.B + b0 + et) * P->ntile_max) * 2;
          np[0] = ts; np[1] = tz;
          for (int t = 1; t < P->ntile_r2; ++t) { np[2 * t] = 0.f; np[2 * t + 1] = 0.f; }
        }
        asm volatile("bar.sync 1, 384;" ::: "memory");
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TA_MMAW) tc::tmem_dealloc(tmem, 512);
}

#endif  // TC_APPLY_KERNEL_TU

struct TcApplyMaps {   // host side: tensor maps of the state read and the state / score written, or use_tma = 0
  CUtensorMap in, out;
  int use_tma;
};

#ifdef TC_APPLY_KERNEL_TU
template <int FMODE, int MODE>
static inline int tc_apply_launch_fm(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream) {
  static CcsdSmemAttr attr;   // per instantiation
  if (ccsd_ensure_smem(tc_apply_kernel<FMODE, MODE>, TA_SMEM, attr)) return -1;
  tc_apply_kernel<FMODE, MODE><<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(dP, a, m.in, m.out, m.use_tma);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
template <int FMODE>
static inline int tc_apply_launch_f(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream) {
  switch (a.mode) {
    case MODE_EVAL: return tc_apply_launch_fm<FMODE, MODE_EVAL>(dP, grid, a, m, stream);
    case MODE_SCORE: return tc_apply_launch_fm<FMODE, MODE_SCORE>(dP, grid, a, m, stream);
    case MODE_PRED: return tc_apply_launch_fm<FMODE, MODE_PRED>(dP, grid, a, m, stream);
    case MODE_NORM: return tc_apply_launch_fm<FMODE, MODE_NORM>(dP, grid, a, m, stream);
    case MODE_CORR: return tc_apply_launch_fm<FMODE, MODE_CORR>(dP, grid, a, m, stream);
  }
  return -1;
}

#endif  // TC_APPLY_KERNEL_TU

// One translation unit per ScoreNetworkF entry path (tc_apply_tu.cu compiled with -DTA_FMODE=k): the 25
// (FMODE, MODE) instantiations are most of the library's compile time, so they build in parallel.
int tc_apply_launch_f0(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream);
int tc_apply_launch_f1(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream);
int tc_apply_launch_f2(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream);
int tc_apply_launch_f3(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream);
int tc_apply_launch_f4(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream);

// 2-D tensor map of a [B*E rows][K cols] fp32 tensor with a box of E rows x 32 columns, SWIZZLE_128B.
// Returns 0 on success; fails (-> cp.async path) when the pitch or the base is not 16-byte aligned.
static inline int tc_apply_make_map(CUtensorMap *m, const float *base, int B, int E, int K, int box_rows) {
  if ((K & 3) || (reinterpret_cast<uintptr_t>(base) & 15) || box_rows > 256) return -1;
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    fn = (EncodeFn)p;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)B * E};
  const cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TA_TN, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : -1;
}

static inline int tc_apply_prepare() { return 0; }   // attributes are set per instantiation at first launch

static inline int tc_apply_launch(const DevPlan *dP, const DevPlan &hp, const ApplyArgs &a, void *stream) {
  const int ngroups = (hp.d.B + hp.ap_group - 1) / hp.ap_group;
  const int grid = ngroups < 148 ? ngroups : 148;
  const int box_rows = hp.ap_group * hp.d.E;
  TcApplyMaps m;
  memset(&m, 0, sizeof m);
  m.use_tma = 0;
  static const bool no_tma = getenv("CCSD_B200_NO_TMA") != nullptr;   // A/B switch for tests and profiling
  if (!no_tma && tc_apply_make_map(&m.in, a.r2, hp.d.B, hp.d.E, hp.d.K, box_rows) == 0 &&
      (a.mode == MODE_NORM || tc_apply_make_map(&m.out, a.out, hp.d.B, hp.d.E, hp.d.K, box_rows) == 0))
    m.use_tma = 1;
  if (a.mode == MODE_NORM) m.out = m.in;
  if (hp.f_mode == 1) return tc_apply_launch_f1(dP, grid, a, m, stream);
  if (hp.f_mode == 2) return tc_apply_launch_f2(dP, grid, a, m, stream);
  if (hp.f_mode == 3) return tc_apply_launch_f3(dP, grid, a, m, stream);
  if (hp.f_mode == 4) return tc_apply_launch_f4(dP, grid, a, m, stream);
  return tc_apply_launch_f0(dP, grid, a, m, stream);
}

}  // namespace ccsd
