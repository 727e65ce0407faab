// tc_agg.cuh -- the GCN aggregation of the large-graph pipeline on the tensor cores:
//   T[i, o] = d_i (sum_j a_ij Y[j, o] + (1 - a_ii) Y[i, o]) + bias[o]        (DenseGCNConv, layers.py:147-156)
// for one adjacency channel of one graph: a real GEMM (N x N) . (N x 96), N = 361 for grid.  Same operand
// construction as layer 1 of tc_afinal_kernel: rows = 128 nodes i per tile, the (symmetric) plane is an MN-major A
// operand as it lies in memory (element (i, j) at plane[j * Np + i]), Y (node-major [N][YW]) is an MN-major B operand;
// both are converted fp32 -> bf16 hi / lo into the canonical SWIZZLE_128B layout, bf16x3 (hi.hi + hi.lo + lo.hi) with
// fp32 accumulation in tensor memory.  The contraction runs over 64-node chunks through a two-stage operand ring:
// the 8 converter warps fill stage s while the MMAs of the previous chunk read stage s ^ 1, and the global loads of
// chunk k + 1 are issued into registers before chunk k is converted (one chunk of load latency is always hidden).
#pragma once
#include "big_pipe.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TGG_EPI = 256;               // converter / epilogue warps: TMEM lane quarter = warp % 4, column part = warp / 4
constexpr int TGG_THREADS = TGG_EPI + 32;   // + the MMA-issuing warp
constexpr int TGG_MMAW = TGG_EPI / 32;
constexpr int TGG_KC = 64;                 // nodes per contraction chunk
constexpr uint32_t TGG_HALF = 2u * TGG_KC * 128u;        // 16384: hi (or lo) of one operand chunk: [2 blocks of 64][64 k][128 B]
constexpr uint32_t TGG_STAGE = 4u * TGG_HALF;            // A hi, A lo, B hi, B lo
constexpr uint32_t TGG_BARS = 2u * TGG_STAGE;            // 131072
constexpr uint32_t TGG_TAB = TGG_BARS + 64;             // [128] destination offsets (int), [128] biases (float)
constexpr size_t TGG_SMEM = (size_t)TGG_TAB + 1024 + 1024;

static inline int tc_agg_supported(const ccsd_attn_layer_t &ly) {
  const int adp = (ly.attn_dim + 7) / 8 * 8, nhp = (ly.conv_out + 7) / 8 * 8, YW = 2 * adp + nhp;
  return YW % 16 == 0 && YW >= 16 && YW <= 128;
}

__global__ void __launch_bounds__(TGG_THREADS, 1) tc_agg_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  extern __shared__ uint8_t tg_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const int N = d.N, Np = L.big_Np;
  const int ad = ly.attn_dim, nh = ly.conv_out, adp = round_up(ad, 8), nhp = round_up(nh, 8), w2 = 2 * adp, YW = w2 + nhp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *W = P->W;

  const uint32_t raw = tc::smem_u32(tg_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = tg_smem_raw + (base - raw);
  const uint32_t bar0 = base + TGG_BARS, tslot = bar0 + 16;     // bar0 + 8 s: MMAs that read stage s have completed
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TGG_BARS + 16);
  int *soff = reinterpret_cast<int *>(gen + TGG_TAB);
  float *sbias = reinterpret_cast<float *>(gen + TGG_TAB + 512);
  if (threadIdx.x == 0) { tc::mbar_init(bar0, 1); tc::mbar_init(bar0 + 8, 1); tc::mbar_fence_init(); }
  if (warp == TGG_MMAW) tc::tmem_alloc(tslot, 128);
  // the n columns [YW, 128) of the B blocks are never read (N = YW); everything else is rewritten per chunk
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t idesc = tc::make_idesc_bf16(128, YW, /*A MN-major*/ 1, /*B MN-major*/ 1);
  const int ntm = (N + 127) >> 7, nkc = (N + TGG_KC - 1) / TGG_KC;
  const int ntiles = d.B * ly.c_in * ntm;
  uint32_t ph[2] = {0u, 0u};     // parity of the next completion of each stage's barrier
  uint32_t used[2] = {0u, 0u};   // stage has MMAs in flight

  constexpr int NA = TGG_KC * 16 / TGG_EPI;     // A items (8 consecutive rows of one contraction index) per thread and chunk
  constexpr int NB = TGG_KC * 16 / TGG_EPI;     // B items (8 consecutive columns), at most (YW <= 128)
  const int nch = YW >> 3;
  float4 ra[NA][2], rb[NB][2];                  // the next chunk's operands, in flight
  auto load_chunk = [&](const float *pl, const float *y, int m0, int k0) {
#pragma unroll
    for (int u = 0; u < NA; ++u) {
      const int t = threadIdx.x + u * TGG_EPI, k = t >> 4, mc = (t & 15) << 3, j = k0 + k, i = m0 + mc;
      float x[8];
      if (j < N && i + 8 <= Np) {
        ra[u][0] = *reinterpret_cast<const float4 *>(pl + (size_t)j * Np + i);
        ra[u][1] = *reinterpret_cast<const float4 *>(pl + (size_t)j * Np + i + 4);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (j < N && i + q < Np) ? pl[(size_t)j * Np + i + q] : 0.f;
        ra[u][0] = make_float4(x[0], x[1], x[2], x[3]);
        ra[u][1] = make_float4(x[4], x[5], x[6], x[7]);
      }
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int t = threadIdx.x + u * TGG_EPI;
      rb[u][0] = rb[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((t & 15) < nch) {
        const int k = t >> 4, n0 = (t & 15) << 3, j = k0 + k;
        if (j < N) {
          rb[u][0] = __ldg(reinterpret_cast<const float4 *>(y + (size_t)j * YW + n0));
          rb[u][1] = __ldg(reinterpret_cast<const float4 *>(y + (size_t)j * YW + n0 + 4));
        }
      }
    }
  };
  auto store_chunk = [&](uint32_t st, int m0) {
#pragma unroll
    for (int u = 0; u < NA; ++u) {
      const int t = threadIdx.x + u * TGG_EPI, k = t >> 4, mc = (t & 15) << 3, i = m0 + mc;
      float x[8] = {ra[u][0].x, ra[u][0].y, ra[u][0].z, ra[u][0].w, ra[u][1].x, ra[u][1].y, ra[u][1].z, ra[u][1].w};
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (i + q >= N) x[q] = 0.f;
      uint4 hi, lo;
      tc::split8(x, hi, lo);
      const uint32_t off = st + (uint32_t)(mc >> 6) * (TGG_KC * 128u) + (uint32_t)k * 128u + (uint32_t)((((mc & 63) >> 3) ^ (k & 7)) << 4);
      *reinterpret_cast<uint4 *>(gen + off) = hi;
      *reinterpret_cast<uint4 *>(gen + off + TGG_HALF) = lo;
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int t = threadIdx.x + u * TGG_EPI;
      if ((t & 15) < nch) {
        const int k = t >> 4, n0 = (t & 15) << 3;
        const float x[8] = {rb[u][0].x, rb[u][0].y, rb[u][0].z, rb[u][0].w, rb[u][1].x, rb[u][1].y, rb[u][1].z, rb[u][1].w};
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        const uint32_t off = st + 2u * TGG_HALF + (uint32_t)(n0 >> 6) * (TGG_KC * 128u) + (uint32_t)k * 128u + (uint32_t)((((n0 & 63) >> 3) ^ (k & 7)) << 4);
        *reinterpret_cast<uint4 *>(gen + off) = hi;
        *reinterpret_cast<uint4 *>(gen + off + TGG_HALF) = lo;
      }
    }
  };
  auto tile_of = [&](int w, int &b, int &c, int &m0, const float *&pl, const float *&y) {
    b = w / (ly.c_in * ntm);
    const int rem = w - b * (ly.c_in * ntm);
    c = rem / ntm;
    m0 = (rem - c * ntm) << 7;
    pl = big_ptr(P, g, b, L.big_S) + (size_t)(g.ch_in + c) * L.big_PS;
    y = big_ptr(P, g, b, L.big_Y) + (size_t)c * N * YW;
  };

  if (warp < TGG_MMAW && (int)blockIdx.x < ntiles) {
    int b, c, m0; const float *pl, *y;
    tile_of(blockIdx.x, b, c, m0, pl, y);
    load_chunk(pl, y, m0, 0);
  }
  for (int w = blockIdx.x; w < ntiles; w += gridDim.x) {
    int b, c, m0; const float *pl, *y;
    tile_of(w, b, c, m0, pl, y);
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc & 1;
      const uint32_t st = (uint32_t)s * TGG_STAGE;
      if (used[s]) {   // the MMAs that read this stage (two chunks ago) must have completed
        tc::mbar_wait(bar0 + 8 * s, ph[s]);
        ph[s] ^= 1u;
        used[s] = 0u;
        tc::tc_fence_after_sync();
      }
      if (warp < TGG_MMAW) {
        store_chunk(st, m0);
        // issue the next chunk's loads (next chunk of this tile, or the first chunk of this CTA's next tile)
        if (kc + 1 < nkc) load_chunk(pl, y, m0, (kc + 1) * TGG_KC);
        else if (w + (int)gridDim.x < ntiles) {
          int b2, c2, m2; const float *pl2, *y2;
          tile_of(w + gridDim.x, b2, c2, m2, pl2, y2);
          load_chunk(pl2, y2, m2, 0);
        }
        tc::fence_proxy_async_smem();
      }
      tc::tc_fence_before_sync();
      __syncthreads();
      if (warp == TGG_MMAW) {
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t blk = TGG_KC * 128u;
#pragma unroll
          for (int k4 = 0; k4 < TGG_KC / 16; ++k4) {
            const uint64_t a_hi = tc::make_smem_desc(base + st + (uint32_t)k4 * 2048u, blk, 1024);
            const uint64_t a_lo = tc::make_smem_desc(base + st + TGG_HALF + (uint32_t)k4 * 2048u, blk, 1024);
            const uint64_t b_hi = tc::make_smem_desc(base + st + 2u * TGG_HALF + (uint32_t)k4 * 2048u, blk, 1024);
            const uint64_t b_lo = tc::make_smem_desc(base + st + 3u * TGG_HALF + (uint32_t)k4 * 2048u, blk, 1024);
            tc::umma_bf16(tmem_u, a_hi, b_hi, idesc, (kc | k4) != 0);
            tc::umma_bf16(tmem_u, a_hi, b_lo, idesc, 1);
            tc::umma_bf16(tmem_u, a_lo, b_hi, idesc, 1);
          }
          tc::umma_commit(bar0 + 8 * s);
        }
        __syncwarp();
      }
      used[s] = 1u;
    }
    // every MMA of the tile has completed once both stages' last commits have arrived (commits complete in order)
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (used[s]) {
        tc::mbar_wait(bar0 + 8 * s, ph[s]);
        ph[s] ^= 1u;
        used[s] = 0u;
      }
    tc::tc_fence_after_sync();
    // ---- epilogue: T = d_i (acc + (1 - a_ii) y_i) + bias, Q | K rows -> TQK, V rows -> TV (feature-major) ----
    // per-column destination (float offset from the graph's scratch base, -1 = padding column) and bias
    if (threadIdx.x < 128) {
      const int o = threadIdx.x;
      int off = -1;
      float bias = 0.f;
      if (o < w2) {
        const int oo = o < adp ? o : o - adp;
        if (oo < ad) { off = L.big_TQK + (c * w2 + o) * Np; bias = __ldg(W + (o < adp ? ly.q[c].b : ly.k[c].b) + oo); }
      } else if (o < YW && o - w2 < nh) {
        off = L.big_TV + (c * nh + (o - w2)) * Np; bias = __ldg(W + ly.v[c].b + (o - w2));
      }
      soff[o] = off; sbias[o] = bias;
    }
    __syncthreads();
    if (warp < TGG_MMAW) {
      constexpr int NPARTS = TGG_EPI / 128;
      const int lq = warp & 3, cpart = warp >> 2, i = m0 + lq * 32 + lane;
      const int nck = YW >> 4, cper = (nck + NPARTS - 1) / NPARTS;
      const int ck0 = cpart * cper < nck ? cpart * cper : nck, ck1 = ck0 + cper < nck ? ck0 + cper : nck;
      const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
      float *gb = big_ptr(P, g, b, 0);
      const float *dvp = big_ptr(P, g, b, L.big_DV) + c * Np;
      const bool live = i < N;
      const float di = live ? dvp[i] : 0.f, fix = live ? 1.f - pl[(size_t)i * Np + i] : 0.f;
      for (int c0 = ck0 * 16; c0 < ck1 * 16; c0 += 16) {
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)c0, v);
        if (live) {
          const float4 *yp = reinterpret_cast<const float4 *>(y + (size_t)i * YW + c0);
          const float4 y0 = __ldg(yp), y1 = __ldg(yp + 1), y2 = __ldg(yp + 2), y3 = __ldg(yp + 3);
          const float yv[16] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w, y2.x, y2.y, y2.z, y2.w, y3.x, y3.y, y3.z, y3.w};
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int off = soff[c0 + q];
            if (off >= 0) gb[(size_t)off + i] = di * (v[q] + fix * yv[q]) + sbias[c0 + q];
          }
        }
      }
    }
    tc::tc_fence_before_sync();
    __syncthreads();   // the accumulator may be overwritten by the next tile
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TGG_MMAW) tc::tmem_dealloc(tmem, 128);
}

static inline int tc_agg_launch(const DevPlan *dP, const DevPlan &hp, const BigArgs &g, void *stream) {
  static CcsdSmemAttr attr;
  if (ccsd_ensure_smem(tc_agg_kernel, TGG_SMEM, attr)) return -1;
  const ccsd_attn_layer_t &ly = hp.d.neta.layer[g.layer];
  const int ntiles = hp.d.B * ly.c_in * ((hp.d.N + 127) / 128);
  tc_agg_kernel<<<ntiles < 148 ? ntiles : 148, TGG_THREADS, TGG_SMEM, (cudaStream_t)stream>>>(dP, g);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccsd
