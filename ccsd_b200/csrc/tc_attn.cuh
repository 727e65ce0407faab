// tc_attn.cuh -- Attention.forward (attention.py:84-132) of one adjacency channel for a GROUP of small graphs on
// the 5th-gen tensor cores.  Replaces attn_channel_kernel (fp32 SIMT, one CTA per (graph, channel)) when the
// graphs are small enough that G = floor(128 / N) >= 2 of them fill one 128-row MMA tile.
//
// Work item = (group of G consecutive graphs, channel c); rows r = (graph gl, node i), R = G N <= 128.
//
//   MMA-A  AX[128 x kin] = blockdiag(An_g) . X      A in TENSOR MEMORY (tcgen05.mma A-from-TMEM): the GCN-normalised adjacency
//                                                   channel (layers.py:139-147) of every graph of the group on the diagonal
//                                                   blocks, zeros elsewhere -- each row's thread packs its bf16 hi / lo pairs
//                                                   straight into its TMEM lane, which frees 64 KB of shared memory and lets
//                                                   TWO CTAs share an SM (the kernel is a chain of short dependent phases);
//                                                   Bx MN-major: the node features (one node row per thread)
//   E1     AX -> bf16 hi/lo K-major A operand (TMEM -> registers -> shared memory, one node row per thread)
//   MMA-B  T[128 x N1]   = AX . [Wq | Wk | Wvw]     B1 MN-major: the channel's weights, converted once per CTA
//   E2     Q = T[:, 0:ad] + bq, K = T[:, adq:adq+ad] + bk -> fp32 in shared memory; T[:, 2adq:] + bvw -> g_hmc
//   S      T_h[i, j] = sum_h tanh(q_i.k_j s) for every ordered pair of a graph (fp32 SIMT, Q row in registers, four threads
//          per row; the SFU tanh is its floor), then att[i <= j] = (T_ij + T_ji) / (2 heads) -> g_att
//
// (An X) W = An (X W) (layers.py:144-156); aggregating first makes the K = 128 MMA the narrow one (N = kin).
// Wvw = Wv . W1_c is the value convolution folded with channel c's slice of multi_channel's first Linear
// (attention.py:292: the node MLP is linear in the channel concat, so V itself is never needed); the packer
// computes it in float64 (ccsd_attn_layer_t.vw).  bf16x3 (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM)
// keeps the 1e-4 parity bar.  One CTA serves one channel (its weights are converted once) and walks groups.
#pragma once
#include "xa_pipe.cuh"
#include "tc_common.cuh"

namespace ccsd {

#ifndef TT_PARTS
#define TT_PARTS 2
#endif
constexpr int TT_NP = TT_PARTS;            // column parts per TMEM lane quarter
constexpr int TT_WORK = 128 * TT_NP;       // worker threads: row = threadIdx.x % 128 (TMEM lane quarter = warp % 4)
constexpr int TT_THREADS = TT_WORK + 32;   // + the MMA-issuing warp
constexpr int TT_MMAW = TT_WORK / 32;
constexpr uint32_t TT_COL_A = 128;         // TMEM columns of the An operand: bf16 pairs hi [128, 192), lo [192, 256); accumulators at [0, 128)

struct TcAttnLayout {
  int G, R;          // graphs per group, rows of a full group
  int K1p;           // conv_in rounded up to 16
  int ad, adq;       // attention width; rounded up to 8 (K columns start at adq, Vw columns at 2 adq)
  int o1, N1p;       // width of the folded value columns; total columns rounded up to 16
  int nblk;          // 64-wide n-blocks of the weight operand
  int nk2;           // k steps of the aggregation MMA
  int qld;           // row pitch (floats) of the fp32 Q / K buffers
  int NP;            // row pitch (floats) of the staged adjacency / score matrices
  uint32_t a1, a1_half, b1, b1_half, a2, a2_half, bx, bx_half, qs, ks, adj, tsc, tij, dvec, vec, bars, total;
};

static inline int tc_attn_layout(const ccsd_plan_desc_t &d, const XpLayout &XL, const ccsd_attn_layer_t &ly, TcAttnLayout &T) {
  const int N = d.N;
  if (XL.big || N > 64 || N < 2) return 0;
  const ccsd_mlp_t &mc = ly.multi_channel;
  const int o1 = mc.nl == 1 ? mc.dout : mc.dhid;
  for (int c = 0; c < ly.c_in; ++c)
    if (ly.vw[c].dout != o1 || ly.vw[c].din != ly.conv_in || ly.vw[c].w <= 0) return 0;   // folded value weights not packed
  if (ly.conv_in > 64 || ly.attn_dim < 1 || ly.conv_mlp) return 0;   // conv == "MLP": Q / K are not graph convolutions
  T.G = 128 / N; T.R = T.G * N;
  T.K1p = (ly.conv_in + 15) & ~15;
  T.ad = ly.attn_dim; T.adq = (ly.attn_dim + 7) & ~7;
  T.o1 = o1;
  T.N1p = (2 * T.adq + ((o1 + 7) & ~7) + 15) & ~15;
  if (T.N1p > 128) return 0;
  T.nblk = (T.N1p + 63) / 64;
  T.nk2 = (T.R + 15) / 16;
  T.qld = ((T.ad + 3) & ~3) + 4;   // multiple of 4, >= adq (E2 stores whole 8-column groups), 4 floats of skew
  T.NP = N | 1;                    // odd pitch: column reads of the symmetrisation are conflict free
  uint32_t o = 0;
  T.bx_half = 16384u; T.bx = o; o += 2 * T.bx_half;                              // X     [128 k-rows][128 B]          MN-major
  T.a1 = T.bx; T.a1_half = T.bx_half;                                            // AX    [128 rows][128 B] K-major: overwrites X (dead after MMA-A)
  T.b1_half = (uint32_t)T.nblk * T.K1p * 128u; T.b1 = o; o += 2 * T.b1_half;     // W     [nblk][K1p k-rows][128 B]    MN-major
  T.a2 = 0; T.a2_half = 0;                                                       // (An lives in tensor memory)
  T.tij = o;                                                                     // everything above is zeroed at start
  o += (uint32_t)XL.ldp * 4u;                                                    // (i << 8) | j of the node pairs
  o = (o + 15u) & ~15u;
  T.qs = o; o += 128u * (uint32_t)T.qld * 4u;
  T.ks = o; o += 128u * (uint32_t)T.qld * 4u;
  T.adj = o; o += (uint32_t)T.G * N * T.NP * 4u;                                 // staged adjacency channel, full matrices ...
  T.tsc = T.adj;                                                                 // ... then the head-summed tanh scores (adj is dead after the An rows are built)
  T.dvec = o; o += 128 * 4;
  T.vec = o; o += 128 * 4;                                                       // biases of the N1p columns
  T.bars = o; o += 64;
  T.total = o + 1024;
  return T.total <= 113u * 1024u;   // two CTAs per SM
}

#ifdef TC_ATTN_KERNEL_TU
struct TcAttnArgs {
  XaArgs x;
  TcAttnLayout L;
  int nper;     // CTAs per channel
};

// q[0..32) . k row, head-summed tanh: heads are chunks of DS features (torch.split: the last one may be short)
template <int DS>
__device__ __forceinline__ float tt_score_row(const float (&q)[32], const float *__restrict__ kj, int ad, float scale) {
  float s = 0.f;
#pragma unroll
  for (int h = 0; h < 32 / DS; ++h) {
    if (h * DS < ad) {
      float u = 0.f;
      if (DS >= 4) {
#pragma unroll
        for (int dd = 0; dd < DS; dd += 4) {
          const float4 k4 = ld4(kj + h * DS + dd);   // columns in [ad, adq) hold exact zeros
          u += q[h * DS + dd] * k4.x + q[h * DS + dd + 1] * k4.y + q[h * DS + dd + 2] * k4.z + q[h * DS + dd + 3] * k4.w;
        }
      } else {
#pragma unroll
        for (int dd = 0; dd < DS; ++dd) u += q[h * DS + dd] * kj[h * DS + dd];
      }
      s += fast_tanh(u * scale);
    }
  }
  return s;
}

// tanh(u s) = 1 - 2 / (2^(u c) + 1) with c = 2 s log2(e): two SFU ops, no range fix-ups (2^x saturates to 0 / inf, both exact limits)
__device__ __forceinline__ float tt_tanh_c(float u, float c) {
  float e, rc;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * c));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(e + 1.0f));
  return fmaf(-2.0f, rc, 1.0f);
}

// the common shape -- exactly four heads of DS features (attn_dim = 4 DS): branch free, so the four dot-product / tanh
// chains of a column (and those of the next column) interleave instead of running back to back behind `h * DS < ad` tests
// volatile shared-memory loads: the compiler keeps them together at the top of the column instead of sinking each one next
// to its four FMAs (which serialised load latency x 8 per column)
__device__ __forceinline__ void tt_lds128(uint32_t addr, float &x, float &y, float &z, float &w) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
}
__device__ __forceinline__ void tt_lds64(uint32_t addr, float &x, float &y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
}

// Two heads x two columns per call: 2 x 2 DS K values are loaded first (all shared-memory loads in flight), then four
// independent dot-product chains and four tanh run interleaved.  Only half of the Q row lives in registers per pass --
// with the whole row (32) + a K row (32) the register allocator serialised every load behind its four FMAs.
template <int DS>
__device__ __forceinline__ void tt_score22(const float (&q)[2 * DS], uint32_t kj0, uint32_t kj1, float c, float &s0, float &s1) {
  float k0[2 * DS], k1[2 * DS];
  if (DS >= 4) {
#pragma unroll
    for (int dd = 0; dd < 2 * DS; dd += 4) tt_lds128(kj0 + dd * 4, k0[dd], k0[dd + 1], k0[dd + 2], k0[dd + 3]);
#pragma unroll
    for (int dd = 0; dd < 2 * DS; dd += 4) tt_lds128(kj1 + dd * 4, k1[dd], k1[dd + 1], k1[dd + 2], k1[dd + 3]);
  } else {
#pragma unroll
    for (int dd = 0; dd < 2 * DS; dd += 2) tt_lds64(kj0 + dd * 4, k0[dd], k0[dd + 1]);
#pragma unroll
    for (int dd = 0; dd < 2 * DS; dd += 2) tt_lds64(kj1 + dd * 4, k1[dd], k1[dd + 1]);
  }
  float u[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dd = 0; dd < DS; ++dd) {
    u[0] = fmaf(q[dd], k0[dd], u[0]);
    u[1] = fmaf(q[DS + dd], k0[DS + dd], u[1]);
    u[2] = fmaf(q[dd], k1[dd], u[2]);
    u[3] = fmaf(q[DS + dd], k1[DS + dd], u[3]);
  }
  s0 = tt_tanh_c(u[0], c) + tt_tanh_c(u[1], c);
  s1 = tt_tanh_c(u[2], c) + tt_tanh_c(u[3], c);
}

// one thread's share of the S phase for four whole heads of DS features: columns j = part, part + TT_NP, ...
template <int DS>
__device__ __forceinline__ void tt_scores4(const float *__restrict__ qr, uint32_t kb, int qld, float *__restrict__ trow_s,
                                            int part, int N, float c) {
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float q[2 * DS];
    if (DS >= 4) {
#pragma unroll
      for (int dd = 0; dd < 2 * DS; dd += 4) {
        const float4 t4 = ld4(qr + pass * 2 * DS + dd);
        q[dd] = t4.x; q[dd + 1] = t4.y; q[dd + 2] = t4.z; q[dd + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int dd = 0; dd < 2 * DS; ++dd) q[dd] = qr[pass * 2 * DS + dd];
    }
    const uint32_t kp = kb + (uint32_t)(pass * 2 * DS) * 4u;
    for (int j = part; j < N; j += 2 * TT_NP) {
      const bool two = j + TT_NP < N;
      const uint32_t kj0 = kp + (uint32_t)(j * qld) * 4u, kj1 = two ? kj0 + (uint32_t)(TT_NP * qld) * 4u : kj0;
      float s0, s1;
      tt_score22<DS>(q, kj0, kj1, c, s0, s1);
      if (pass == 0) {
        trow_s[j] = s0;
        if (two) trow_s[j + TT_NP] = s1;
      } else {
        trow_s[j] += s0;
        if (two) trow_s[j + TT_NP] += s1;
      }
    }
  }
}

// Narrow attention (attn_dim <= 12, e.g. QM9: attn_dim 10 in 5 heads of 2): the whole Q row (AP = attn_dim rounded up to 4,
// zero padded) stays in registers, two K rows are loaded per iteration, NH independent head chains per column.
template <int AP, int DS, int NH>
__device__ __forceinline__ void tt_scores_w(const float *__restrict__ qr, uint32_t kb, int qld, float *__restrict__ trow_s,
                                            int part, int N, float c) {
  static_assert(NH * DS <= AP && AP % 4 == 0, "head layout");
  float q[AP];
#pragma unroll
  for (int dd = 0; dd < AP; dd += 4) {
    const float4 t4 = ld4(qr + dd);
    q[dd] = t4.x; q[dd + 1] = t4.y; q[dd + 2] = t4.z; q[dd + 3] = t4.w;
  }
  for (int j = part; j < N; j += 2 * TT_NP) {
    const bool two = j + TT_NP < N;
    const uint32_t kj0 = kb + (uint32_t)(j * qld) * 4u, kj1 = two ? kj0 + (uint32_t)(TT_NP * qld) * 4u : kj0;
    float k0[AP], k1[AP];
#pragma unroll
    for (int dd = 0; dd < AP; dd += 4) tt_lds128(kj0 + dd * 4, k0[dd], k0[dd + 1], k0[dd + 2], k0[dd + 3]);
#pragma unroll
    for (int dd = 0; dd < AP; dd += 4) tt_lds128(kj1 + dd * 4, k1[dd], k1[dd + 1], k1[dd + 2], k1[dd + 3]);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float u0 = 0.f, u1 = 0.f;
#pragma unroll
      for (int dd = 0; dd < DS; ++dd) {
        u0 = fmaf(q[h * DS + dd], k0[h * DS + dd], u0);
        u1 = fmaf(q[h * DS + dd], k1[h * DS + dd], u1);
      }
      s0 += tt_tanh_c(u0, c);
      s1 += tt_tanh_c(u1, c);
    }
    trow_s[j] = s0;
    if (two) trow_s[j + TT_NP] = s1;
  }
}

__device__ __forceinline__ void tt_tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// debug timeline (ccsd_debug_apply_trace): stamp `slot` of this CTA's item `it`, CTA 0 only
#define TT_STAMP(slot_) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0 && it < 32) a.trace[it * 16 + (slot_)] = clock64(); } while (0)

__global__ void __launch_bounds__(TT_THREADS, 2) tc_attn_kernel(const DevPlan *__restrict__ P, TcAttnArgs ta) {
  extern __shared__ uint8_t tt_smem_raw[];
  const XaArgs &a = ta.x;
  const TcAttnLayout &T = ta.L;
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_neta_t &A = d.neta;
  const ccsd_attn_layer_t &ly = A.layer[a.layer];
  const int N = d.N, N4 = L.N4, NT = L.NT, ldp = L.ldp, B = d.B;
  const int G = T.G, R = T.R, K1p = T.K1p, ad = T.ad, adq = T.adq, o1 = T.o1, N1p = T.N1p, qld = T.qld, NP = T.NP;
  const int kin = ly.conv_in;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (int)blockIdx.x % ly.c_in;             // this CTA's channel
  const int g_first = (int)blockIdx.x / ly.c_in;
  const int ngroups = (B + G - 1) / G;
  const float *W = P->W;

  const uint32_t raw = tc::smem_u32(tt_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = tt_smem_raw + (base - raw);
  const uint32_t bar = base + T.bars, tslot = bar + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + T.bars + 8);
  int it = 0;
  TT_STAMP(15);
  float *adj_s = reinterpret_cast<float *>(gen + T.adj), *tsc = reinterpret_cast<float *>(gen + T.tsc);
  float *vbias = reinterpret_cast<float *>(gen + T.vec);
  float *qs = reinterpret_cast<float *>(gen + T.qs), *ks = reinterpret_cast<float *>(gen + T.ks);
  int *tij = reinterpret_cast<int *>(gen + T.tij);

  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (warp == TT_MMAW) tc::tmem_alloc(tslot, 256);
  // ---- zero every operand buffer (pad rows / columns and the off-diagonal blocks of An stay zero), tables, biases ----
  for (uint32_t o = threadIdx.x * 16u; o < T.tij; o += TT_THREADS * 16u) *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  for (int t = threadIdx.x; t < ldp; t += TT_THREADS) tij[t] = t < NT ? P->tri_ij[t] : 0;
  for (int n = threadIdx.x; n < 128; n += TT_THREADS) {
    float v = 0.f;
    if (n < ad) v = __ldg(W + ly.q[c].b + n);
    else if (n >= adq && n < adq + ad) v = __ldg(W + ly.k[c].b + (n - adq));
    else if (n >= 2 * adq && n < 2 * adq + o1) v = __ldg(W + ly.vw[c].b + (n - 2 * adq));
    vbias[n] = v;
  }
  __syncthreads();
  {
    // [Wq | Wk | Wvw] (each (in = k, out_pad) row-major) -> MN-major B operand:
    //   (n, k) at (n/64)*blk + k*128 + (((n%64)/8) ^ (k%8))*16 + (n%8)*2
    const int nchunk = N1p >> 3, o1p = (o1 + 7) & ~7;
    const uint32_t blk = (uint32_t)K1p * 128u;
    for (int t = threadIdx.x; t < kin * nchunk; t += TT_THREADS) {
      const int k = t / nchunk, nc = t - k * nchunk, n0 = nc << 3;
      const float *src = nullptr;
      int lim = 0;   // valid columns of this chunk
      if (n0 < adq) { src = W + ly.q[c].w + (size_t)k * adq + n0; lim = ad - n0; }
      else if (n0 < 2 * adq) { src = W + ly.k[c].w + (size_t)k * adq + (n0 - adq); lim = ad - (n0 - adq); }
      else if (n0 < 2 * adq + o1p) { src = W + ly.vw[c].w + (size_t)k * o1p + (n0 - 2 * adq); lim = o1 - (n0 - 2 * adq); }
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = (src && q < lim) ? __ldg(src + q) : 0.f;
      uint4 hi, lo;
      tc::split8(x, hi, lo);
      const uint32_t off = T.b1 + (uint32_t)(n0 >> 6) * blk + (uint32_t)k * 128u + (uint32_t)((((n0 & 63) >> 3) ^ (k & 7)) << 4);
      *reinterpret_cast<uint4 *>(gen + off) = hi;
      *reinterpret_cast<uint4 *>(gen + off + T.b1_half) = lo;
    }
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  if (warp < 4) {   // the An operand region starts as zeros: a warp only ever rewrites the chunks that hold its rows' diagonal blocks
    const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    for (int cc = 0; cc < 16; ++cc) tt_tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + TT_COL_A + (uint32_t)(cc * 8), z);
    tc::tmem_st_wait();
    tc::tc_fence_before_sync();
  }
  const uint32_t idescA = tc::make_idesc_bf16(128, K1p, /*A K-major*/ 0, /*B MN-major*/ 1);
  const uint32_t idescB = tc::make_idesc_bf16(128, N1p, /*A K-major*/ 0, /*B MN-major*/ 1);
  uint32_t phase = 0;

  // per-thread row geometry (constant over the items): row r = (graph gl of the group, node ni)
  const int r = threadIdx.x & 127, part = (threadIdx.x >> 7);   // part: 0..TT_NP-1 (MMA warp: unused)
  const int gl = r / N, ni = r - gl * N;
  const bool row_in = r < R;
  const int lq = warp & 3;
  const float scale = 1.0f / sqrtf((float)ly.conv_out);   // / sqrt(out_dim)  (attention.py:125)
  const int ds = ad / A.num_heads, nch = (ad + ds - 1) / ds;
  const float inv = 0.5f / (float)nch;
  const int k0 = gl * N;                                    // first column of this row's diagonal block

  for (int gi = g_first; gi < ngroups; gi += ta.nper, ++it) {
    TT_STAMP(0);
    const int b0 = gi * G;
    const int gsz = B - b0 < G ? B - b0 : G;
    const bool live = row_in && gl < gsz;
    // this row's node features (its part's 8-feature chunks): issued first, consumed in L(b) after the adjacency is staged
    constexpr int TT_XR = 64 / 8 / TT_NP * 8;   // conv_in <= 64
    float xr[TT_XR];
    if (warp < TT_MMAW) {
      const float *gx = a.g_xin + (size_t)(b0 + (live ? gl : 0)) * L.g_x + ni;
#pragma unroll
      for (int u = 0; u < TT_XR / 8; ++u) {
        const int q8 = part + u * TT_NP;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int f = q8 * 8 + q;
          xr[u * 8 + q] = (live && f < kin) ? __ldg(gx + (size_t)f * N4) : 0.f;
        }
      }
    }
    if (warp < TT_MMAW) {
      // ---- L(a): the group's adjacency channel (triangle storage in global memory) -> full symmetric matrices ----
      //      (all of a batch's global loads are issued before the first scatter store: one L2 round trip per batch of 8)
      const float *src = a.g_stack + (size_t)b0 * L.g_stack + (size_t)(a.ch_in + c) * ldp;
      const int tot = gsz * NT;
      for (int i0 = (int)threadIdx.x; i0 < tot; i0 += 8 * TT_WORK) {
        float v[8];
        int gt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + u * TT_WORK;
          const int g2 = idx / NT, t = idx - g2 * NT;
          gt[u] = idx < tot ? ((g2 << 16) | t) : -1;
          v[u] = idx < tot ? __ldg(src + (size_t)g2 * L.g_stack + t) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (gt[u] >= 0) {
            const int g2 = gt[u] >> 16, t = gt[u] & 0xffff;
            const int ij = tij[t], i = ij >> 8, j = ij & 255;
            const float vv = i == j ? 1.f : v[u];   // A^ = A with a unit diagonal (layers.py:139-141)
            float *dst = adj_s + g2 * N * NP;
            dst[i * NP + j] = vv;
            dst[j * NP + i] = vv;
          }
        }
      }
    }
    __syncthreads();
    TT_STAMP(1);
    // ---- L(c): GCN degree of the thread's own node, d_i = clamp(rowsum(A^), 1)^-1/2 (layers.py:142-145).  An = D A^ D is
    //      applied as D (A^ (D X)): X's row r is scaled by d_r here, AX's row r by d_r in E1 -- both by the row's own thread ----
    float di = 0.f;
    if (warp < TT_MMAW) {
      const float *__restrict__ row = adj_s + (gl * N + ni) * NP;
      if (live) {
        float s0 = 0.f, s1 = 0.f;
        int j = 0;
        for (; j + 1 < N; j += 2) { s0 += row[j]; s1 += row[j + 1]; }
        if (j < N) s0 += row[j];
        di = rsqrtf(fmaxf(s0 + s1, 1.f));
      }
      // ---- L(b): node features -> Bx (MN-major: k = node row, n = feature), 8 features per 16-byte chunk ----
      {
#pragma unroll
        for (int u = 0; u < TT_XR / 8; ++u) {
          const int q8 = part + u * TT_NP;
          if (q8 >= (K1p >> 3)) break;
          float x[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = di * xr[u * 8 + q];
          uint4 hi, lo;
          tc::split8(x, hi, lo);
          const uint32_t off = T.bx + (uint32_t)r * 128u + (uint32_t)((q8 ^ (r & 7)) << 4);
          *reinterpret_cast<uint4 *>(gen + off) = hi;
          *reinterpret_cast<uint4 *>(gen + off + T.bx_half) = lo;
        }
      }
      TT_STAMP(2);
      // ---- L(d): row r of blockdiag(A^) -> this row's TMEM lane: element k in 32-bit column k / 2 (bf16 pairs), hi then lo.
      //      Chunks of 16 k values (one tcgen05.st of 8 columns each for hi and lo); the chunk range is warp uniform: it covers
      //      the diagonal blocks of the warp's 32 rows (zeros where a lane's own block does not reach) ----
      {
        const int rlo = lq * 32, rhi = rlo + 31 < R - 1 ? rlo + 31 : R - 1;
        const int ck_lo = rlo < R ? ((rlo / N) * N) >> 4 : 0, ck_hi = rlo < R ? ((rhi / N) * N + N + 15) >> 4 : 0;
        const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16) + TT_COL_A;
        for (int ck = ck_lo + part; ck < ck_hi && ck < 8; ck += TT_NP) {
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int j = ck * 16 + h8 * 8 + q - k0;
              float v = 0.f;
              if (live && j >= 0 && j < N) v = row[j];
              x[q] = v;
            }
            uint4 hi, lo;
            tc::split8(x, hi, lo);
            hw[h8 * 4 + 0] = hi.x; hw[h8 * 4 + 1] = hi.y; hw[h8 * 4 + 2] = hi.z; hw[h8 * 4 + 3] = hi.w;
            lw[h8 * 4 + 0] = lo.x; lw[h8 * 4 + 1] = lo.y; lw[h8 * 4 + 2] = lo.z; lw[h8 * 4 + 3] = lo.w;
          }
          tt_tmem_st8(trow + (uint32_t)(ck * 8), hw);
          tt_tmem_st8(trow + 64u + (uint32_t)(ck * 8), lw);
        }
        tc::tmem_st_wait();
      }
      tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    TT_STAMP(3);
    // ---- MMA-A: AX = blockdiag(An) . X ----
    if (warp == TT_MMAW) {
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        for (int k4 = 0; k4 < T.nk2; ++k4) {
          const uint32_t a_hi = tmem_u + TT_COL_A + (uint32_t)(k4 * 8), a_lo = a_hi + 64u;
          const uint64_t b_hi = tc::make_smem_desc(base + T.bx + (uint32_t)k4 * 2048u, 16384, 1024);
          const uint64_t b_lo = tc::make_smem_desc(base + T.bx + T.bx_half + (uint32_t)k4 * 2048u, 16384, 1024);
          tc::umma_bf16_ts(tmem_u, a_hi, b_hi, idescA, k4 != 0);
          tc::umma_bf16_ts(tmem_u, a_hi, b_lo, idescA, 1);
          tc::umma_bf16_ts(tmem_u, a_lo, b_hi, idescA, 1);
        }
        tc::umma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    tc::tc_fence_after_sync();
    TT_STAMP(4);
    // ---- E1: AX (one node row per thread) -> A1 (K-major) ----
    if (warp < TT_MMAW) {
      const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
      for (int ck = part; ck < (K1p >> 4); ck += TT_NP) {
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)(ck * 16), v);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] *= di;
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          uint4 hi, lo;
          tc::split8(v + h8 * 8, hi, lo);
          const int q8 = ck * 2 + h8;
          const uint32_t off = T.a1 + (uint32_t)r * 128u + (uint32_t)((q8 ^ (r & 7)) << 4);
          *reinterpret_cast<uint4 *>(gen + off) = hi;
          *reinterpret_cast<uint4 *>(gen + off + T.a1_half) = lo;
        }
      }
      tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    TT_STAMP(5);
    // ---- MMA-B: T = AX . [Wq | Wk | Wvw] ----
    if (warp == TT_MMAW) {
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t blk = (uint32_t)K1p * 128u;
        for (int k4 = 0; k4 < K1p / 16; ++k4) {
          const uint64_t a_hi = tc::make_smem_desc(base + T.a1 + (uint32_t)k4 * 32u, 0, 1024);
          const uint64_t a_lo = tc::make_smem_desc(base + T.a1 + T.a1_half + (uint32_t)k4 * 32u, 0, 1024);
          const uint64_t b_hi = tc::make_smem_desc(base + T.b1 + (uint32_t)k4 * 2048u, blk, 1024);
          const uint64_t b_lo = tc::make_smem_desc(base + T.b1 + T.b1_half + (uint32_t)k4 * 2048u, blk, 1024);
          tc::umma_bf16(tmem_u, a_hi, b_hi, idescB, k4 != 0);
          tc::umma_bf16(tmem_u, a_hi, b_lo, idescB, 1);
          tc::umma_bf16(tmem_u, a_lo, b_hi, idescB, 1);
        }
        tc::umma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    tc::tc_fence_after_sync();
    TT_STAMP(6);
    // ---- E2: Q, K (+ bias) -> fp32 shared memory; folded value columns (+ bias) -> g_hmc ----
    if (warp < TT_MMAW) {
      const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);   // T overwrote AX in the accumulator columns
      float *gh = a.g_hmc + (size_t)(b0 + (live ? gl : 0)) * L.g_hmc + (size_t)c * L.mc_o1_max * N4 + ni;
      for (int ck = part; ck < (N1p >> 4); ck += TT_NP) {
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)(ck * 16), v);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int n0 = ck * 16 + h8 * 8;
          if (n0 < 2 * adq) {
            float *dst = (n0 < adq ? qs : ks) + r * qld + (n0 < adq ? n0 : n0 - adq);
            const float4 v0 = make_float4(v[h8 * 8 + 0] + vbias[n0 + 0], v[h8 * 8 + 1] + vbias[n0 + 1], v[h8 * 8 + 2] + vbias[n0 + 2], v[h8 * 8 + 3] + vbias[n0 + 3]);
            const float4 v1 = make_float4(v[h8 * 8 + 4] + vbias[n0 + 4], v[h8 * 8 + 5] + vbias[n0 + 5], v[h8 * 8 + 6] + vbias[n0 + 6], v[h8 * 8 + 7] + vbias[n0 + 7]);
            *reinterpret_cast<float4 *>(dst) = v0;       // columns in [ad, adq) get exact zeros (zero weights, zero bias)
            *reinterpret_cast<float4 *>(dst + 4) = v1;
          } else if (live) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int o = n0 + q - 2 * adq;
              if (o < o1) gh[(size_t)o * N4] = v[h8 * 8 + q] + vbias[n0 + q];
            }
          }
        }
      }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    TT_STAMP(7);
    // ---- S: head-summed tanh scores of every ordered node pair of the row's graph: thread (row, part) takes the
    //         columns j = part, part + 4, ... with the Q row in registers ----
    if (warp < TT_MMAW && live) {
      float *__restrict__ trow_s = tsc + (gl * N + ni) * NP;
      const float *__restrict__ kbase = ks + k0 * qld;
      if (ad == 4 * ds && (ds == 8 || ds == 4 || ds == 2)) {
        // four whole heads (every shipped checkpoint): branch-free scores
        const float c2 = scale * 2.885390081777927f;   // 2 log2(e)
        const uint32_t kb = tc::smem_u32(kbase);
        const float *__restrict__ qr = qs + r * qld;
        if (ds == 8) tt_scores4<8>(qr, kb, qld, trow_s, part, N, c2);
        else if (ds == 4) tt_scores4<4>(qr, kb, qld, trow_s, part, N, c2);
        else tt_scores4<2>(qr, kb, qld, trow_s, part, N, c2);
      } else if (ad == 10 && ds == 2) {   // QM9 / QM9_CC: adim 10, 4 "heads" -> torch.split gives 5 chunks of 2
        tt_scores_w<12, 2, 5>(qs + r * qld, tc::smem_u32(kbase), qld, trow_s, part, N, scale * 2.885390081777927f);
      } else if (ad <= 32 && (ds == 8 || ds == 4 || ds == 2)) {
        float q[32];
        const float *__restrict__ qr = qs + r * qld;
#pragma unroll
        for (int dd = 0; dd < 32; dd += 4) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (dd < adq) t4 = ld4(qr + dd);
          q[dd] = t4.x; q[dd + 1] = t4.y; q[dd + 2] = t4.z; q[dd + 3] = t4.w;
        }
        {
          for (int j = part; j < N; j += TT_NP) {
            const float *kj = kbase + j * qld;
            trow_s[j] = ds == 8 ? tt_score_row<8>(q, kj, ad, scale) : (ds == 4 ? tt_score_row<4>(q, kj, ad, scale) : tt_score_row<2>(q, kj, ad, scale));
          }
        }
      } else {
        const float *qr = qs + r * qld;
        for (int j = part; j < N; j += TT_NP) {
          const float *kj = kbase + j * qld;
          float s = 0.f;
          for (int h = 0; h < nch; ++h) {
            const int d0 = h * ds, d1 = (d0 + ds < ad) ? d0 + ds : ad;
            float u = 0.f;
            for (int dd = d0; dd < d1; ++dd) u += qr[dd] * kj[dd];
            s += fast_tanh(u * scale);
          }
          trow_s[j] = s;
        }
      }
    }
    __syncthreads();
    TT_STAMP(8);
    // ---- symmetrise (attention.py:130) and write the attention map in triangle storage ----
    if (warp < TT_MMAW) {
      float *dst = a.g_att + (size_t)b0 * L.g_att + (size_t)c * ldp;
      const int tot = gsz * ldp;
      for (int i0 = (int)threadIdx.x; i0 < tot; i0 += 4 * TT_WORK) {
        float sv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = i0 + u * TT_WORK;
          const int g2 = idx / ldp, t = idx - g2 * ldp;
          sv[u] = 0.f;
          if (idx < tot && t < NT) {
            const float *ts = tsc + g2 * N * NP;
            const int ij = tij[t], i = ij >> 8, j = ij & 255;
            sv[u] = inv * (ts[i * NP + j] + ts[j * NP + i]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = i0 + u * TT_WORK;
          const int g2 = idx / ldp, t = idx - g2 * ldp;
          if (idx < tot) dst[(size_t)g2 * L.g_att + t] = sv[u];
        }
      }
    }
    __syncthreads();   // the score matrices share their buffer with the next item's staged adjacency
    TT_STAMP(9);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TT_MMAW) tc::tmem_dealloc(tmem, 256);
}

#endif  // TC_ATTN_KERNEL_TU

// defined in tc_attn_tu.cu (its own translation unit)
int tc_attn_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, const TcAttnLayout &T, void *stream);

#ifdef TC_ATTN_KERNEL_TU
int tc_attn_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, const TcAttnLayout &T, void *stream) {
  const ccsd_attn_layer_t &ly = hp.d.neta.layer[a.layer];
  TcAttnArgs ta;
  ta.x = a;
  ta.L = T;
  const int ngroups = (hp.d.B + T.G - 1) / T.G;
  int nper = (148 * 2) / ly.c_in;   // two CTAs per SM
  if (nper < 1) nper = 1;
  if (nper > ngroups) nper = ngroups;
  ta.nper = nper;
  static CcsdSmemAttr attr;
  if (ccsd_ensure_smem(tc_attn_kernel, T.total, attr)) return -1;
  tc_attn_kernel<<<nper * ly.c_in, TT_THREADS, T.total, (cudaStream_t)stream>>>(dP, ta);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

#endif  // TC_ATTN_KERNEL_TU

}  // namespace ccsd
