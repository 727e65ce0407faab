// tc_afinal.cuh -- final per-edge MLP of ScoreNetworkA (ScoreNetwork_A.py:529-539) as a fused tcgen05 MLP,
// plus the adjacency sampler epilogue.  Replaces afinal_kernel when the MLP is 3 Linears with one output
// (every shipped checkpoint) and fits: fd <= 64 channels, hidden width <= 128.
//
//   rows    = node pairs (i <= j) of one graph, 128 per tile (the channel stack [fd][ldp] of the x/adj pipeline
//             is feature-major, i.e. an MN-major A operand as it lies in memory)
//   layer 1 : D1[128 x H] = X[128 x fd] . W1          A, B MN-major bf16 hi/lo in shared memory
//   layer 2 : D2[128 x H] = elu(D1 + b1) . W2          A in TENSOR MEMORY (tcgen05.mma A-from-TMEM): each row's thread writes
//                                                      elu(D1 + b1) back to its TMEM lane as packed bf16 hi / lo pairs -- no
//                                                      64 KB shared-memory operand, so small networks fit TWO CTAs per SM
//                                                      (the kernel is a chain of dependent phases per tile); B MN-major
//   layer 3 : out = elu(D2 + b2) . w3 + b3             H FMAs per row in the epilogue
// bf16x3 (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM) keeps the 1e-4 parity bar.  W1 / W2 are converted
// once per CTA and stay resident (<= 96 KB); a persistent CTA walks tiles (graph, row block).
#pragma once
#include "xa_pipe.cuh"
#include "tc_common.cuh"

namespace ccsd {

// Column parts NP: the epilogue of a 128-row tile is split over NP warps per TMEM lane quarter (TMEM lane quarter = warp % 4,
// column part = warp / 4), + 1 MMA warp.  NP = 4, one CTA per SM: large networks (grid: 125 KB of shared memory);
// NP = 2, two CTAs per SM: everything that fits 113 KB.
constexpr int TF_NP_MAX = 4;
constexpr uint32_t TF_COL_A2 = 128;   // TMEM columns: D1 / D2 at [0, 128), the layer-2 A operand (bf16 pairs) hi at [128, 192), lo at [192, 256)

struct TcFinLayout {   // byte offsets from the 1024-aligned base; *_half = distance hi -> lo
  int K1p, Hp;         // fd rounded up to 16, hidden width rounded up to 16
  int KC;              // channels of the X tile staged per layer-1 pass (K1p, or K1p / 2 when only that fits two CTAs per SM)
  uint32_t w1, w1_half, w2, w2_half, a1, a1_half, a2, a2_half, vec, bars, total;
};

static inline int tc_afinal_supported(const ccsd_neta_t &A, int fd_have) {
  return A.fin.nl == 3 && A.fin.dout == 1 && fd_have >= 1 && fd_have <= 64 && A.fin.dhid >= 8 && A.fin.dhid <= 128;
}
static inline TcFinLayout tc_afinal_layout(int fd, int dhid, int kc = 0) {
  TcFinLayout L;
  L.K1p = (fd + 15) & ~15;
  L.Hp = (dhid + 15) & ~15;
  L.KC = kc > 0 ? kc : L.K1p;
  uint32_t o = 0;
  L.w1_half = 2u * L.K1p * 128u; L.w1 = o; o += 2 * L.w1_half;     // [2 n-blocks][K1p k-rows][128 B]
  L.w2_half = 2u * L.Hp * 128u;  L.w2 = o; o += 2 * L.w2_half;     // [2 n-blocks][Hp k-rows][128 B]
  L.a1_half = 2u * L.KC * 128u;  L.a1 = o; o += 2 * L.a1_half;     // [2 m-blocks][KC k-rows][128 B]
  L.a2_half = 0; L.a2 = 0;                                        // (the layer-2 A operand lives in tensor memory)
  L.vec = o; o += 3 * 128 * 4 + 256 + 512 * TF_NP_MAX;                 // b1, b2, w3, reduction scratch, [parts][128] partial dot products
  L.bars = o; o += 64;
  L.total = o + 1024;
  return L;
}

struct TcFinArgs {
  XaArgs x;
  TcFinLayout L;
  int fd;              // channels in the stack
  int ntg;             // tiles per graph
  // geometry of the channel stack: per-graph-tile pipeline = triangle rows [fd][ldp]; large-graph pipeline
  // (big_pipe.cuh) = full planes [fd][N][Np], tile = 128 columns of one row, tiles below the diagonal skipped
  const float *gs_base;
  long long gs_stride; // floats per graph
  int ldp, NT;         // plane stride; valid rows of a plane (triangle path)
  int big, Np, nseg;   // large-graph path: row pitch, 128-column segments per row
  int gpt, rpg;        // tiny graphs (NT <= 64): gpt graphs share one 128-row tile, rpg = NT rounded up to 8 rows each (else 1, 128)
};

template <int TF_NP>
__global__ void __launch_bounds__(128 * TF_NP + 32, TF_NP == 2 ? 2 : 1) tc_afinal_kernel(const DevPlan *__restrict__ P, TcFinArgs ta) {
  constexpr int TF_EPI = 128 * TF_NP, TF_THREADS = TF_EPI + 32, TF_MMAW = TF_EPI / 32;
  extern __shared__ uint8_t tf_smem_raw[];
  const XaArgs &a = ta.x;
  const TcFinLayout &TL = ta.L;
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_mlp_t &fin = d.neta.fin;
  const int N = d.N, NP = N * N, NT = ta.NT, ldp = ta.ldp, fd = ta.fd, dh = fin.dhid;
  const int K1p = TL.K1p, Hp = TL.Hp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *W = P->W;

  const uint32_t raw = tc::smem_u32(tf_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = tf_smem_raw + (base - raw);
  const uint32_t bar = base + TL.bars, tslot = bar + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TL.bars + 8);
  float *vb1 = reinterpret_cast<float *>(gen + TL.vec), *vb2 = vb1 + 128, *vw3 = vb2 + 128, *red = vw3 + 128, *part = red + 64;

  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (warp == TF_MMAW) tc::tmem_alloc(tslot, 256);
  // ---- zero the operand buffers, then convert the weights (resident for the whole kernel) ----
  for (uint32_t o = threadIdx.x * 16u; o < TL.vec; o += TF_THREADS * 16u) *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 128; i += TF_THREADS) {
    vb1[i] = i < dh ? __ldg(W + fin.b[0] + i) : 0.f;
    vb2[i] = i < dh ? __ldg(W + fin.b[1] + i) : 0.f;
    vw3[i] = i < dh ? __ldg(W + fin.w[2] + i * 8) : 0.f;   // (in, out_pad = 8), output 0
  }
  __syncthreads();
  {
    // Linear l: (in = k, out_pad) row-major -> MN-major B operand: (n, k) at (n/64)*blk + k*128 + (((n%64)/8) ^ (k%8))*16 + (n%8)*2
    const int opad = round_up(dh, 8);
    for (int l = 0; l < 2; ++l) {
      const int Kin = l == 0 ? fd : dh;
      const uint32_t blk = (uint32_t)(l == 0 ? K1p : Hp) * 128u, dst = l == 0 ? TL.w1 : TL.w2, half = l == 0 ? TL.w1_half : TL.w2_half;
      const int nchunk = opad >> 3;
      for (int t = threadIdx.x; t < Kin * nchunk; t += TF_THREADS) {
        const int k = t / nchunk, nc = t - k * nchunk, n0 = nc << 3;
        const float *src = W + fin.w[l] + (size_t)k * opad + n0;
        float x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = n0 + q < dh ? __ldg(src + q) : 0.f;
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        const uint32_t off = dst + (uint32_t)(n0 >> 6) * blk + (uint32_t)k * 128u + (uint32_t)((((n0 & 63) >> 3) ^ (k & 7)) << 4);
        *reinterpret_cast<uint4 *>(gen + off) = hi;
        *reinterpret_cast<uint4 *>(gen + off + half) = lo;
      }
    }
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t id1 = tc::make_idesc_bf16(128, Hp, /*A MN-major*/ 1, /*B MN-major*/ 1);
  const uint32_t id2 = tc::make_idesc_bf16(128, Hp, /*A K-major*/ 0, /*B MN-major*/ 1);
  const float b3 = __ldg(W + fin.b[2]);
  uint32_t phase = 0;
  const int ntiles = ta.gpt > 1 ? (d.B + ta.gpt - 1) / ta.gpt : d.B * ta.ntg;

  for (int w = blockIdx.x; w < ntiles; w += gridDim.x) {
    const int gpt = ta.gpt, rpg = ta.rpg;
    const int b = gpt > 1 ? w * gpt : w / ta.ntg, tt = gpt > 1 ? 0 : w - b * ta.ntg;   // (first) graph of the tile, tile of the graph
    const int b0t = b;
    int t0 = tt * 128, bi = 0, bj0 = 0, rows = NT - t0;   // rows: valid rows of this tile (of each graph when gpt > 1)
    if (ta.big) {
      bi = tt / ta.nseg; bj0 = (tt - bi * ta.nseg) * 128;
      if (bj0 + 128 <= bi) {   // wholly below the diagonal: the mirrored tile covers it (uniform branch)
        if (a.mode == MODE_SCORE && threadIdx.x == 0) {
          float *np = a.norm_part + ((size_t)(1 * d.B + b) * P->ntile_max + tt) * 2;
          np[0] = 0.f; np[1] = 0.f;
        }
        continue;
      }
      t0 = bi * ta.Np + bj0;
      rows = N - bj0;
    }
    const float *gs = ta.gs_base + (size_t)b * ta.gs_stride;
    // ---- X tile -> A1 (MN-major): chunk = 8 consecutive rows of one channel.  Large networks (grid: 64 channels) stage the
    //      tile in two passes of KC channels, so that the kernel fits two CTAs per SM ----
    const int KC = TL.KC, npass = K1p / KC;
   for (int pass = 0; pass < npass; ++pass) {
    if (warp < TF_MMAW) {
      const int kend = npass == 1 ? fd : KC;   // (two passes: the pad channels of the last pass are written as zeros)
      for (int t = threadIdx.x; t < kend * 16; t += TF_EPI) {
        const int kl = t >> 4, k = pass * KC + kl, mc = t & 15, m0 = mc << 3;
        const float *src = gs + (size_t)k * ldp + t0 + m0;
        float x[8];
        if (k >= fd) {
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = 0.f;
        } else if (gpt > 1) {   // rows [g rpg, g rpg + NT) of the tile = the pairs of graph b + g
          const int g = m0 / rpg, tg = m0 - g * rpg;
          const bool ok = g < gpt && b + g < d.B;
          const float *sg = gs + (size_t)g * ta.gs_stride + (size_t)k * ldp + tg;
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = (ok && tg + q < NT) ? sg[q] : 0.f;
        } else if (t0 + m0 + 8 <= ldp) {
          const float4 v0 = *reinterpret_cast<const float4 *>(src), v1 = *reinterpret_cast<const float4 *>(src + 4);
          x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = t0 + m0 + q < ldp ? src[q] : 0.f;
        }
        if (gpt == 1) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (m0 + q >= rows) x[q] = 0.f;
        }
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        const uint32_t off = TL.a1 + (uint32_t)(m0 >> 6) * ((uint32_t)KC * 128u) + (uint32_t)kl * 128u +
                             (uint32_t)((((m0 & 63) >> 3) ^ (kl & 7)) << 4);
        *reinterpret_cast<uint4 *>(gen + off) = hi;
        *reinterpret_cast<uint4 *>(gen + off + TL.a1_half) = lo;
      }
      tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    // ---- layer 1 MMAs ----
    if (warp == TF_MMAW) {
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t blk = (uint32_t)K1p * 128u, blka = (uint32_t)KC * 128u;
        for (int k4 = 0; k4 < KC / 16; ++k4) {
          const uint32_t kw = (uint32_t)(pass * (KC / 16) + k4);   // k step inside the resident W1
          const uint64_t a_hi = tc::make_smem_desc(base + TL.a1 + (uint32_t)k4 * 2048u, blka, 1024);
          const uint64_t a_lo = tc::make_smem_desc(base + TL.a1 + TL.a1_half + (uint32_t)k4 * 2048u, blka, 1024);
          const uint64_t b_hi = tc::make_smem_desc(base + TL.w1 + kw * 2048u, blk, 1024);
          const uint64_t b_lo = tc::make_smem_desc(base + TL.w1 + TL.w1_half + kw * 2048u, blk, 1024);
          tc::umma_bf16(tmem_u, a_hi, b_hi, id1, (pass | k4) != 0);
          tc::umma_bf16(tmem_u, a_hi, b_lo, id1, 1);
          tc::umma_bf16(tmem_u, a_lo, b_hi, id1, 1);
        }
        tc::umma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);   // (between passes: the MMAs have read A1 before it is overwritten)
    phase ^= 1u;
    tc::tc_fence_after_sync();
   }
    // ---- epilogue 1: elu(D1 + b1) -> the layer-2 A operand in this row's TMEM lane (element k in 32-bit column k / 2) ----
    const int lq = warp & 3, cpart = warp >> 2;               // TMEM lane quarter, column part of this warp
    const int nck = Hp >> 4, cper = (nck + TF_NP - 1) / TF_NP;   // 16-column chunks, chunks per part
    const int ck0 = cpart * cper < nck ? cpart * cper : nck, ck1 = ck0 + cper < nck ? ck0 + cper : nck;
    if (warp < TF_MMAW) {
      const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
      for (int c0 = ck0 * 16; c0 < ck1 * 16; c0 += 16) {
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)c0, v);
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          float x[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + h8 * 8 + q;
            x[q] = c < dh ? fast_elu(v[h8 * 8 + q] + vb1[c]) : 0.f;
          }
          uint4 hi, lo;
          tc::split8(x, hi, lo);
          hw[h8 * 4 + 0] = hi.x; hw[h8 * 4 + 1] = hi.y; hw[h8 * 4 + 2] = hi.z; hw[h8 * 4 + 3] = hi.w;
          lw[h8 * 4 + 0] = lo.x; lw[h8 * 4 + 1] = lo.y; lw[h8 * 4 + 2] = lo.z; lw[h8 * 4 + 3] = lo.w;
        }
        tc::tmem_st8(trow + TF_COL_A2 + (uint32_t)(c0 >> 1), hw);
        tc::tmem_st8(trow + TF_COL_A2 + 64u + (uint32_t)(c0 >> 1), lw);
      }
      tc::tmem_st_wait();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    // ---- layer 2 MMAs ----
    if (warp == TF_MMAW) {
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t blk = (uint32_t)Hp * 128u;
        for (int k4 = 0; k4 < Hp / 16; ++k4) {
          const uint32_t a_hi = tmem_u + TF_COL_A2 + (uint32_t)(k4 * 8), a_lo = a_hi + 64u;   // 16 k values = 8 columns
          const uint64_t b_hi = tc::make_smem_desc(base + TL.w2 + (uint32_t)k4 * 2048u, blk, 1024);
          const uint64_t b_lo = tc::make_smem_desc(base + TL.w2 + TL.w2_half + (uint32_t)k4 * 2048u, blk, 1024);
          tc::umma_bf16_ts(tmem_u, a_hi, b_hi, id2, k4 != 0);   // D2 overwrites D1 (consumed by epilogue 1)
          tc::umma_bf16_ts(tmem_u, a_hi, b_lo, id2, 1);
          tc::umma_bf16_ts(tmem_u, a_lo, b_hi, id2, 1);
        }
        tc::umma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    tc::tc_fence_after_sync();
    // ---- epilogue 2: out = elu(D2 + b2) . w3 + b3, masks, adjacency sampler epilogue ----
    float s2 = 0.f, z2 = 0.f;
    float acc = 0.f;
    if (warp < TF_MMAW) {
      const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
      for (int c0 = ck0 * 16; c0 < ck1 * 16; c0 += 16) {
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
        for (int q = 0; q < 16; ++q) acc += fast_elu(v[q] + vb2[c0 + q]) * vw3[c0 + q];   // vw3 = 0 past dh
      }
      if (cpart) part[(cpart - 1) * 128 + lq * 32 + lane] = acc;
    }
    __syncthreads();   // the second column half's partial dot products
    if (warp < 4) {
      const int r = threadIdx.x;
      const int gg = gpt > 1 ? r / rpg : 0;                 // graph of this row inside the tile
      const int t = gpt > 1 ? r - gg * rpg : t0 + r;
      const int b = b0t + gg;
#pragma unroll
      for (int q = 0; q < TF_NP - 1; ++q) acc += part[q * 128 + r];
      acc += b3;
      int i = 0, j = 0;
      bool live = gpt > 1 ? (gg < gpt && b < d.B && t < NT) : r < rows;
      if (ta.big) { i = bi; j = bj0 + r; live = live && j >= i; }
      else if (live) { const int ij = P->tri_ij[t]; i = ij >> 8; j = ij & 255; }
      if (live) {
        const float fi = a.flags[(size_t)b * N + i], fj = a.flags[(size_t)b * N + j];
        const float o = (i == j) ? 0.f : acc * fi * fj;   // (1 - I) mask and mask_adjs
        const size_t ga = (size_t)b * NP;
        if (a.mode == MODE_EVAL) {
          a.out_adj[ga + i * N + j] = o;
          a.out_adj[ga + j * N + i] = o;
        } else {
          const int stp = nz_step(a.nz);
          const ccsd_objcoef_t ca = P->sched[stp * 3 + 1];
          const unsigned long long gsid = (unsigned long long)(a.nz.sample_offset + b);
          const float s = ca.score_scale * o;
          float z = 0.f;
          if (i != j) {
            const int q = i * N + j;
            z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gsid, draw_id(1, stp, a.slot), q)) * fi * fj;
          }
          if (a.mode == MODE_SCORE) {
            a.out_adj[ga + i * N + j] = s;
            if (i != j) {
              a.out_adj[ga + j * N + i] = s;
              s2 = 2.f * s * s;
              z2 = 2.f * z * z;
            }
          } else {
            const float m = ca.pa * a.adj[ga + i * N + j] + ca.pb * s;
            const float v = m + ca.pc * z;
            a.out_adj[ga + i * N + j] = v;
            a.mean_adj[ga + i * N + j] = m;
            float *tja = a.nz.sd ? a.nz.sd->ta : a.traj_adj;
            if (tja && b == 0) tja[i * N + j] = a.denoise ? m : v;
            if (i != j) {
              a.out_adj[ga + j * N + i] = v;
              a.mean_adj[ga + j * N + i] = m;
              if (tja && b == 0) tja[j * N + i] = a.denoise ? m : v;
            }
          }
        }
      }
    }
    if (a.mode == MODE_SCORE) {
      if (gpt > 1) {
        // several graphs per tile: per-row values to shared memory (each thread only ever touches its own slots of `part`),
        // summed per graph in row order after the barrier
        if (warp < 4) { part[threadIdx.x] = s2; part[128 + threadIdx.x] = z2; }
      } else {
        // per-tile norm partial (fixed order): warps 0-3 reduce, thread 0 sums the four warp values
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          z2 += __shfl_xor_sync(0xffffffffu, z2, o);
        }
        if (lane == 0 && warp < 4) { red[warp] = s2; red[8 + warp] = z2; }
      }
    }
    tc::tc_fence_before_sync();
    __syncthreads();   // every thread is done with D1 / D2 / A1 / A2 of this tile
    if (a.mode == MODE_SCORE) {
      if (gpt > 1) {
        if ((int)threadIdx.x < gpt && b0t + (int)threadIdx.x < d.B) {
          float t2 = 0.f, u2 = 0.f;
          for (int q = 0; q < NT; ++q) { t2 += part[threadIdx.x * rpg + q]; u2 += part[128 + threadIdx.x * rpg + q]; }
          float *np = a.norm_part + ((size_t)(1 * d.B + b0t + threadIdx.x) * P->ntile_max) * 2;
          np[0] = t2; np[1] = u2;
        }
      } else if (threadIdx.x == 0) {
        float *np = a.norm_part + ((size_t)(1 * d.B + b) * P->ntile_max + tt) * 2;
        np[0] = red[0] + red[1] + red[2] + red[3];
        np[1] = red[8] + red[9] + red[10] + red[11];
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TF_MMAW) tc::tmem_dealloc(tmem, 256);
}

static inline int tc_afinal_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, int fd, void *stream) {
  TcFinArgs ta;
  ta.x = a;
  ta.L = tc_afinal_layout(fd, hp.d.neta.fin.dhid);
  if (ta.L.total > 113u * 1024u && (ta.L.K1p / 2) % 16 == 0) {   // two channel passes if that is what fits two CTAs per SM
    const TcFinLayout L2 = tc_afinal_layout(fd, hp.d.neta.fin.dhid, ta.L.K1p / 2);
    if (L2.total <= 113u * 1024u) ta.L = L2;
  }
  ta.fd = fd;
  ta.ntg = (hp.xp.NT + 127) / 128;
  ta.gs_base = a.g_stack; ta.gs_stride = hp.xp.g_stack; ta.ldp = hp.xp.ldp; ta.NT = hp.xp.NT;
  ta.big = 0; ta.Np = 0; ta.nseg = 1;
  ta.gpt = 1; ta.rpg = 128;
  if (!hp.xp.big && hp.xp.NT <= 64) { ta.rpg = (hp.xp.NT + 7) & ~7; ta.gpt = 128 / ta.rpg; }
  if (hp.xp.big) {
    ta.big = 1; ta.Np = hp.xp.big_Np; ta.nseg = (hp.d.N + 127) / 128;
    ta.ntg = hp.d.N * ta.nseg;
    ta.gs_base = a.g_stack;          // the caller passes the large-graph stack base in g_stack
    ta.gs_stride = hp.xp.big_total; ta.ldp = hp.xp.big_PS; ta.NT = hp.xp.big_PS;
  }
  const int ntiles = ta.gpt > 1 ? (hp.d.B + ta.gpt - 1) / ta.gpt : hp.d.B * ta.ntg;
  if (ta.L.total <= 113u * 1024u) {   // two CTAs per SM, 288 threads each
    static CcsdSmemAttr attr2;
    if (ccsd_ensure_smem(tc_afinal_kernel<2>, ta.L.total, attr2)) return -1;
    tc_afinal_kernel<2><<<ntiles < 296 ? ntiles : 296, 128 * 2 + 32, ta.L.total, (cudaStream_t)stream>>>(dP, ta);
  } else {
    static CcsdSmemAttr attr4;
    if (ccsd_ensure_smem(tc_afinal_kernel<4>, ta.L.total, attr4)) return -1;
    tc_afinal_kernel<4><<<ntiles < 148 ? ntiles : 148, 128 * 4 + 32, ta.L.total, (cudaStream_t)stream>>>(dP, ta);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ccsd
