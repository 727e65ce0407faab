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
// (cp.async, 4 stages x 24 KB, 16-byte chunks XOR-swizzled by the row so that both the row-wise loader
// accesses and the column-wise epilogue accesses are conflict free).  The epilogue threads -- one per
// edge row, i.e. per TMEM lane -- read their row's 16 cells from the staging tile, combine them with the
// accumulator, Philox noise and masks, and write the NEW state back into the staging tile in place;
// the loader warps then copy the finished tile out with coalesced 16-byte stores.
//
// Warp roles (21 warps): 0-3 loaders (cp.async ring, bf16 split, copy-out), 4-19 epilogue (TMEM lane
// quarter q = warp % 4; M tile and cell half from the warp index), 20 MMA issuer.
#pragma once
#include "r2_kernels.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TA_LOAD = 128;
constexpr int TA_EPI = 512;
constexpr int TA_THREADS = TA_LOAD + TA_EPI + 32;
constexpr int TA_TN = 32;                         // cells per tile
constexpr int TA_NE = 192;                        // padded edge count
constexpr int TA_NS = 4;                          // staging stages
constexpr int TA_PD = 2;                          // cp.async prefetch distance (tiles)
constexpr uint32_t TA_OPHALF = TA_NE * 128u;      // 24576: hi (or lo) operand rows, 64 cells (2 tile slots) x bf16
constexpr uint32_t TA_OPER = 2u * TA_OPHALF;      // 49152
constexpr uint32_t TA_STAGE = TA_NE * 128u;       // 24576: 192 rows x 32 fp32
constexpr uint32_t TA_BARS = TA_OPER + TA_NS * TA_STAGE;   // 147456
constexpr uint32_t TA_FCS = TA_BARS + 256;        // [NS][32] cell flags
constexpr uint32_t TA_RED = TA_FCS + TA_NS * 32 * 4;   // [2][16] warp partials
constexpr uint32_t TA_FW = TA_RED + 256;          // staged ScoreNetworkF weights (FMODE 2)
constexpr size_t TA_SMEM = (size_t)TA_FW + 6144 + 1024 /*alignment slack*/;
constexpr uint32_t TA_COL_D = 384;                // first accumulator column

static inline int tc_apply_supported(int E, int K) { return E >= 8 && E <= TA_NE && K >= 8; }

// one entry (edge e, cell k): network -> score -> mode-specific value.  Returns the value that replaces
// the entry in the staging tile (raw output, scaled score or new state).
template <int FMODE>
__device__ __forceinline__ float r2_entry(const R2Epi &c, const ApplyArgs &a, int e, int k, float f, float hf, float zraw,
                                          float m, float cs, float cn, float &s2, float &z2) {
  const DevPlan *P = c.P;
  float o;
  if (FMODE == 1) o = m * (c.aff0 * f + c.aff1 * hf + c.aff2);
  else if (FMODE == 2) o = netf_entry_w8(P->d.netf, c.fw, c.f_nlin, f, hf, m);
  else o = netf_entry(P->d.netf, P->W, f, hf, m);
  if (a.mode == MODE_EVAL) return o;
  const float s = c.co.score_scale * o;
  const float z = zraw * m;
  if (a.mode == MODE_SCORE || a.mode == MODE_NORM) {
    s2 += s * s;
    z2 += z * z;
    return s;
  }
  if (a.mode == MODE_CORR) return f + cs * s + cn * z;   // Langevin (solver.py:784-785)
  const float mu = c.co.pa * f + c.co.pb * s;            // predictor (solver.py:283-300, 433-450)
  const float v = mu + c.co.pc * z;
  if (a.write_mean | (a.traj != nullptr)) {
    const size_t g = ((size_t)c.b * c.E + e) * c.K + k;
    if (a.write_mean) a.mean[g] = mu;
    if (a.traj && c.b == 0) a.traj[(size_t)e * c.K + k] = a.denoise ? mu : v;
  }
  return v;
}

template <int FMODE>
__global__ void __launch_bounds__(TA_THREADS, 1) tc_apply_kernel(const DevPlan *__restrict__ P, ApplyArgs a) {
  extern __shared__ uint8_t ta_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, E = d.E, K = d.K, B = d.B, Ep = P->Ep;
  const int ntile = (K + TA_TN - 1) / TA_TN;
  const int nk = (E + 15) >> 4;                 // 16-wide k steps over e'
  const int mtiles = E > 128 ? 2 : 1;
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
                 stage_full = epi_done + 8 * TA_NS, h_ready = stage_full + 8 * TA_NS, tslot = h_ready + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TA_BARS + 64 + 16 * TA_NS + 8);
  float *fcs = reinterpret_cast<float *>(gen + TA_FCS);
  float *red = reinterpret_cast<float *>(gen + TA_RED);
  float *fw = reinterpret_cast<float *>(gen + TA_FW);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(full + 8 * s, TA_LOAD);
      tc::mbar_init(op_empty + 8 * s, 1);
      tc::mbar_init(t_full + 8 * s, 1);
      tc::mbar_init(d_empty + 8 * s, TA_EPI);
    }
    for (int s = 0; s < TA_NS; ++s) {
      tc::mbar_init(epi_done + 8 * s, TA_EPI);
      tc::mbar_init(stage_full + 8 * s, TA_LOAD);
    }
    tc::mbar_init(h_ready, TA_EPI);
    tc::mbar_fence_init();
  }
  if (warp == 20) tc::tmem_alloc(tslot, 512);
  if (FMODE == 2) netf_stage_w8(d.netf, P->W, fw);
  // operand rows that no edge fills stay zero for the whole kernel (their A columns are zero too, but
  // 0 * NaN from uninitialised shared memory would poison the accumulator)
  for (uint32_t o = threadIdx.x * 16u; o < TA_OPER; o += TA_THREADS * 16u)
    *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const int nmine = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ntot = nmine * ntile;

  if (warp < 4) {
    // ===================== loaders =====================
    const bool vec = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.r2) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
    const int cu = threadIdx.x & 7, r0 = threadIdx.x >> 3;   // 16-byte chunk (4 cells) of rows r0 + 16 j
    const bool writes = a.mode != MODE_NORM;
    auto tile_of = [&](int g, int &b, int &k0) {
      const int si = g / ntile;
      b = (int)blockIdx.x + si * (int)gridDim.x;
      k0 = (g - si * ntile) * TA_TN;
    };
    auto issue = [&](int g) {
      int b, k0;
      tile_of(g, b, k0);
      const float *Fb = a.r2 + (size_t)b * E * K;
      const uint32_t st = sStage + (uint32_t)(g % TA_NS) * TA_STAGE;
      const int k = k0 + 4 * cu;
      for (int r = r0; r < E; r += 16) {
        const uint32_t dst = st + (uint32_t)r * 128u + (uint32_t)((cu ^ (r & 7)) << 4);
        const float *src = Fb + (size_t)r * K + k;
        if (vec) {
          const int nb = k + 4 <= K ? 16 : (k < K ? (K - k) * 4 : 0);
          tc::cp_async16(dst, nb ? (const void *)src : (const void *)Fb, (uint32_t)nb);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            tc::cp_async4(dst + 4u * i, (k + i < K) ? (const void *)(src + i) : (const void *)Fb, (k + i < K) ? 4u : 0u);
        }
      }
    };
    auto copy_out = [&](int g) {
      tc::mbar_wait(epi_done + 8 * (g % TA_NS), (uint32_t)(g / TA_NS) & 1u);
      if (!writes) return;
      int b, k0;
      tile_of(g, b, k0);
      float *Ob = a.out + (size_t)b * E * K;
      const uint8_t *st = gen + TA_OPER + (size_t)(g % TA_NS) * TA_STAGE;
      const int k = k0 + 4 * cu;
      if (k >= K) return;
      for (int r = r0; r < E; r += 16) {
        const float4 v = *reinterpret_cast<const float4 *>(st + (size_t)r * 128 + (size_t)((cu ^ (r & 7)) << 4));
        float *dst = Ob + (size_t)r * K + k;
        if (vec) {
          __stcs(reinterpret_cast<float4 *>(dst), v);
        } else {
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k + i < K) __stcs(dst + i, vv[i]);
        }
      }
    };
    for (int t = 0; t < TA_PD; ++t) {
      if (t < ntot) issue(t);
      tc::cp_async_commit();
    }
    int cur_b = -1;
    unsigned long long zm = 0ull;
    for (int g = 0; g < ntot; ++g) {
      if (g + TA_PD < ntot) {
        if (g + TA_PD - TA_NS >= 0) copy_out(g + TA_PD - TA_NS);   // frees the stage tile g + PD lands in
        issue(g + TA_PD);
      }
      tc::cp_async_commit();
      tc::cp_async_wait<TA_PD>();                                   // this thread's chunks of tile g have landed
      int b, k0;
      tile_of(g, b, k0);
      if (b != cur_b) { cur_b = b; zm = zero_mask_of(a.flags + (size_t)b * N, N); }
      if (g >= 2) tc::mbar_wait(op_empty + 8 * (g & 1), (uint32_t)(((g >> 1) & 1) ^ 1));
      uint8_t *st = gen + TA_OPER + (size_t)(g % TA_NS) * TA_STAGE;
      const uint32_t slot4 = (uint32_t)(g & 1) * 4u;
      for (int r = r0; r < E; r += 16) {
        const float4 v = *reinterpret_cast<const float4 *>(st + (size_t)r * 128 + (size_t)((cu ^ (r & 7)) << 4));
        uint2 hi, lo;
        tc::split4(v, hi, lo);
        const uint32_t off = (uint32_t)r * 128u + (((slot4 + (uint32_t)(cu >> 1)) ^ (uint32_t)(r & 7)) << 4) + (uint32_t)(cu & 1) * 8u;
        *reinterpret_cast<uint2 *>(gen + off) = hi;
        *reinterpret_cast<uint2 *>(gen + TA_OPHALF + off) = lo;
      }
      if (threadIdx.x < TA_TN) {
        const int k = k0 + (int)threadIdx.x;
        fcs[(g % TA_NS) * TA_TN + threadIdx.x] = (k < K && !(P->cell_mask[k] & zm)) ? 1.f : 0.f;
      }
      tc::fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
      tc::mbar_arrive(full + 8 * (g & 1));
      tc::mbar_arrive(stage_full + 8 * (g % TA_NS));
    }
    // drain: tiles [ntot - NS, ntot) are still in the ring (the loop copied out tile g + PD - NS)
    for (int g = ntot > TA_NS ? ntot - TA_NS : 0; g < ntot; ++g) copy_out(g);
  } else if (warp == 20) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::make_idesc_bf16(128, TA_TN, /*A from TMEM*/ 0, /*B MN-major*/ 1);
    for (int g = 0; g < ntot; ++g) {
      const int slot = g & 1;
      const int si = g / ntile;
      if (g - si * ntile == 0) tc::mbar_wait(h_ready, (uint32_t)si & 1u);
      tc::mbar_wait(full + 8 * slot, (uint32_t)(g >> 1) & 1u);
      tc::mbar_wait(d_empty + 8 * slot, (uint32_t)(((g >> 1) & 1) ^ 1));
      tc::tc_fence_after_sync();
      if (lane == 0) {
        for (int mt = 0; mt < mtiles; ++mt) {
          const uint32_t dcol = tmem + TA_COL_D + (uint32_t)(slot * 64 + mt * 32);
          const uint32_t a_hi = tmem + (uint32_t)(mt * 192), a_lo = a_hi + 96u;
          for (int k4 = 0; k4 < nk; ++k4) {
            const uint32_t bo = (uint32_t)slot * 64u + (uint32_t)k4 * 2048u;   // 16 e' rows of 128 bytes
            const uint64_t b_hi = tc::make_smem_desc(sOp + bo, 8192, 1024);
            const uint64_t b_lo = tc::make_smem_desc(sOp + TA_OPHALF + bo, 8192, 1024);
            tc::umma_bf16_ts(dcol, a_hi + (uint32_t)k4 * 8u, b_hi, idesc, k4 != 0);
            tc::umma_bf16_ts(dcol, a_hi + (uint32_t)k4 * 8u, b_lo, idesc, 1);
            tc::umma_bf16_ts(dcol, a_lo + (uint32_t)k4 * 8u, b_hi, idesc, 1);
          }
        }
        tc::umma_commit(op_empty + 8 * slot);   // operand slot may be refilled once these MMAs retire
        tc::umma_commit(t_full + 8 * slot);     // accumulators complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 4-19) =====================
    const int ew = warp - 4;             // 0..15
    const int q = ew & 3;                // TMEM lane quarter (= warp % 4)
    const int mt = (ew >> 2) & 1;        // M tile
    const int half = ew >> 3;            // cells [16 half, 16 half + 16) of the tile; e' half for the H fill
    const int et = threadIdx.x - TA_LOAD;
    const int Kg = P->Kp >> 2;
    // the edge row this thread owns: M tile 0 holds edges 0..127 on lanes 0..127; M tile 1 holds edges
    // 128 + 16 q + l on lanes 32 q + l, l < 16 (so that every lane quarter carries the same load)
    int e = -1;
    if (mt == 0) e = q * 32 + lane;
    else if (lane < 16) e = 128 + q * 16 + lane;
    if (e >= E) e = -1;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float cs = 0.f, cn = 0.f;
    if (a.mode == MODE_CORR) { cs = a.coef[4]; cn = a.coef[5]; }
    const int ei = e >= 0 ? P->edge_ij[2 * e] : 0, ej = e >= 0 ? P->edge_ij[2 * e + 1] : 0;
    for (int si = 0; si < nmine; ++si) {
      const int b = (int)blockIdx.x + si * (int)gridDim.x;
      const float *fl = a.flags + (size_t)b * N;
      R2Epi c;
      c.P = P; c.fl = fl; c.fw = fw; c.zm = 0ull;
      c.gs = (unsigned long long)(a.nz.sample_offset + b);
      if (a.mode != MODE_EVAL) c.co = P->sched[a.nz.step * 3 + 2];
      c.b = b; c.E = E; c.K = K; c.Kg = Kg; c.f_nlin = P->f_nlin;
      c.aff0 = d.netf.aff[0]; c.aff1 = d.netf.aff[1]; c.aff2 = d.netf.aff[2];
      const float fe = e >= 0 ? fl[ei] * fl[ej] : 0.f;
      // ---- H of this sample -> TMEM (A operand).  Every MMA of the previous sample has completed: this
      // warp waited on t_full of its last tile. ----
      if (mt < mtiles) {
        const float *Hrow = a.H + ((size_t)b * E + (e >= 0 ? e : 0)) * Ep;
        for (int c16 = 0; c16 < 3; ++c16) {
          const int kb = half * 96 + c16 * 32;          // first e' of this 32-element (16-column) group
          uint32_t hw[16], lw[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = kb + 4 * j4;
            if (e >= 0 && k < Ep) v = __ldg(reinterpret_cast<const float4 *>(Hrow + k));
            if (k + 0 >= E) v.x = 0.f;
            if (k + 1 >= E) v.y = 0.f;
            if (k + 2 >= E) v.z = 0.f;
            if (k + 3 >= E) v.w = 0.f;
            uint2 hi, lo;
            tc::split4(v, hi, lo);
            hw[2 * j4] = hi.x; hw[2 * j4 + 1] = hi.y;
            lw[2 * j4] = lo.x; lw[2 * j4 + 1] = lo.y;
          }
          const uint32_t col = (uint32_t)(mt * 192 + half * 48 + c16 * 16);
          tc::tmem_st16(tmem + lane_base + col, hw);
          tc::tmem_st16(tmem + lane_base + col + 96u, lw);
        }
        tc::tmem_st_wait();
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(h_ready);

      float s2 = 0.f, z2 = 0.f;
      const float *Nb = a.noise ? a.noise + (size_t)b * E * K : nullptr;
      for (int ct = 0; ct < ntile; ++ct) {
        const int g = si * ntile + ct;
        const int slot = g & 1, stg = g % TA_NS;
        tc::mbar_wait(stage_full + 8 * stg, (uint32_t)(g / TA_NS) & 1u);   // staging tile + cell flags visible
        tc::mbar_wait(t_full + 8 * slot, (uint32_t)(g >> 1) & 1u);   // accumulators complete
        tc::tc_fence_after_sync();
        float v[16];
        tc::tmem_ld16(tmem + lane_base + TA_COL_D + (uint32_t)(slot * 64 + mt * 32 + half * 16), v);
        tc::tc_fence_before_sync();
        tc::mbar_arrive(d_empty + 8 * slot);                          // accumulator slot may be overwritten
        if (e >= 0) {
          uint8_t *row = gen + TA_OPER + (size_t)stg * TA_STAGE + (size_t)e * 128;
          const float *fc = fcs + stg * TA_TN + half * 16;
          const int kbase = ct * TA_TN + half * 16;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int k = kbase + 4 * c4;
            float4 *cell = reinterpret_cast<float4 *>(row + (((half * 4 + c4) ^ (e & 7)) << 4));
            const float4 f4 = *cell;
            const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
            float z4[4] = {0.f, 0.f, 0.f, 0.f};
            if (a.mode != MODE_EVAL && k < K) {
              if (Nb) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (k + i < K) z4[i] = Nb[(size_t)e * K + k + i];
              } else {
                normal4(a.nz.seed, c.gs, draw_id(2, a.nz.step, a.slot), (uint32_t)(e * Kg + (k >> 2)), z4);
              }
            }
            float o4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o4[i] = (k + i < K) ? r2_entry<FMODE>(c, a, e, k + i, fv[i], v[4 * c4 + i], z4[i], fe * fc[4 * c4 + i], cs, cn, s2, z2)
                                  : 0.f;
            if (a.mode != MODE_NORM) *cell = make_float4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
        tc::mbar_arrive(epi_done + 8 * stg);                          // (release) tile may be copied out
      }
      if (a.mode == MODE_SCORE || a.mode == MODE_NORM) {
        // per-sample squared norms: reduce over the 16 epilogue warps (named barrier 1, 512 threads)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          z2 += __shfl_xor_sync(0xffffffffu, z2, o);
        }
        if (lane == 0) { red[ew] = s2; red[16 + ew] = z2; }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (et == 0) {
          float ts = 0.f, tz = 0.f;
          for (int w = 0; w < 16; ++w) { ts += red[w]; tz += red[16 + w]; }
          float *np = a.norm_part + ((size_t)(2 * d.B + b) * P->ntile_max) * 2;
          np[0] = ts; np[1] = tz;
          for (int t = 1; t < P->ntile_r2; ++t) { np[2 * t] = 0.f; np[2 * t + 1] = 0.f; }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 20) tc::tmem_dealloc(tmem, 512);
}

static inline int tc_apply_prepare() {
  cudaError_t e = cudaFuncSetAttribute(tc_apply_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TA_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TA_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TA_SMEM);
  return e == cudaSuccess ? 0 : -1;
}

static inline int tc_apply_launch(const DevPlan *dP, const DevPlan &hp, const ApplyArgs &a, void *stream) {
  const int grid = hp.d.B < 148 ? hp.d.B : 148;
  if (hp.f_mode == 1) tc_apply_kernel<1><<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(dP, a);
  else if (hp.f_mode == 2) tc_apply_kernel<2><<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(dP, a);
  else tc_apply_kernel<0><<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(dP, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccsd
