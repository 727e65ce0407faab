// tc_xfin.cuh -- the final node MLP of ScoreNetworkX (ScoreNetwork_X.py:121-131: fdim -> 2 fdim -> 2 fdim -> F over every
// node, 93 % of the network's FLOPs) as a fused tcgen05 MLP, plus the x sampler epilogue.  x_net_kernel keeps the GCN
// stack and hands the concatenated node features [x, h_1 .. h_D] over in global memory (feature-major, L2 resident).
//
//   rows    = the nodes of G = floor(128 / N) consecutive graphs (one graph never straddles a tile)
//   layer 1 : D[128 x H]  = X[128 x fdim] . W1        A: shared memory, K-major bf16 hi/lo (one node row per thread)
//   layer 2 : D[128 x H]  = elu(D + b1)   . W2        A: TENSOR MEMORY -- the epilogue threads (one row = one TMEM lane
//                                                     each) write elu(D + b1) back as packed bf16 hi / lo pairs, so the
//                                                     hidden activations never touch shared memory (tcgen05.mma A-from-TMEM)
//   layer 3 : out[16]     = elu(D + b2)   . W3 + b3   A: tensor memory again (N = 16 MMAs; F <= 16 outputs)
// The weights do not fit shared memory (W2 alone is 229 KB as bf16 hi + lo at H = 224): they are converted ONCE per
// sampler run into their swizzled operand image (tc_xfin_prep_kernel) and STREAMED from L2 by a copy warp with
// cp.async.bulk (the TMA engine) into a two-slot ring, one chunk = (column half of H) x (<= 128 k rows); seven per tile.
// bf16x3 (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM) keeps the 1e-4 parity bar.
//
// Warps: 0-15 workers (row = thread % 128 = TMEM lane, column part = thread / 128), 16 MMA issuer, 17 weight copy.
#pragma once
#include "xa_pipe.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TX_NP = 4;
constexpr int TX_WORK = 128 * TX_NP;
constexpr int TX_THREADS = TX_WORK + 64;
constexpr int TX_MMAW = TX_WORK / 32, TX_CPW = TX_MMAW + 1;

struct TcXfinLayout {
  int G, R;                 // graphs per tile, rows of a full tile
  int fd, K1p;              // input width, rounded up to 16 (<= 128)
  int dh, Hp, NH, NHb;      // hidden width, rounded up to 32 (<= 256); column half; its 64-wide n-blocks
  int KC2;                  // k rows per layer-2 chunk (= Hp / 2)
  int F, nkb1;              // outputs (<= 16); 64-wide k-blocks of the X operand
  uint32_t chunk1, chunk2, chunk3, chunk;   // bytes of a layer-1 / layer-2 / layer-3 weight chunk (hi + lo), ring slot size
  uint32_t a1, a1_half, wb, vec, sc, sq, bars, total;
  long long img_bytes;      // operand image in global memory: 2 layer-1 chunks + 4 layer-2 chunks + 1 layer-3 chunk
};

static inline int tc_xfin_layout(const ccsd_plan_desc_t &d, const XpLayout &XL, TcXfinLayout &T) {
  const ccsd_netx_t &X = d.netx;
  const ccsd_mlp_t &fin = X.fin;
  if (!(d.nets & 1) || XL.big || d.N > 64 || d.N < 2) return 0;
  if (fin.nl != 3 || fin.din != X.fdim || fin.din > 128 || fin.dhid > 256 || fin.dhid < 16 || fin.dout > 16 || fin.dout != d.F) return 0;
  T.G = 128 / d.N; T.R = T.G * d.N;
  T.fd = fin.din; T.K1p = (fin.din + 15) & ~15;
  T.dh = fin.dhid; T.Hp = (fin.dhid + 31) & ~31; T.NH = T.Hp / 2; T.NHb = (T.NH + 63) / 64;
  T.KC2 = T.Hp / 2;
  T.F = fin.dout; T.nkb1 = (T.K1p + 63) / 64;
  T.chunk1 = 2u * T.NHb * T.K1p * 128u;
  T.chunk2 = 2u * T.NHb * T.KC2 * 128u;
  T.chunk3 = 2u * (uint32_t)T.Hp * 128u;             // one 64-wide n-block (16 columns used) x Hp k rows
  T.chunk = T.chunk1 > T.chunk2 ? T.chunk1 : T.chunk2;
  if (T.chunk3 > T.chunk) T.chunk = T.chunk3;
  T.img_bytes = 2ll * T.chunk1 + 4ll * T.chunk2 + T.chunk3;
  uint32_t o = 0;
  T.a1_half = (uint32_t)T.nkb1 * 16384u; T.a1 = o; o += 2 * T.a1_half;
  T.wb = o; o += 2 * T.chunk;
  T.vec = o; o += (256 + 256 + 16) * 4;              // b1, b2, b3
  T.sc = o; o += 128 * 16 * 4;                       // network output of the tile [row][16]
  T.sq = o; o += 2 * 128 * 16 * 4;                   // squared score / noise entries [2][row * F + f] for the fixed-order norms
  T.bars = o; o += 128;
  T.total = o + 1024;
  return T.total <= 227u * 1024u;
}

struct TcXfinArgs {
  XaArgs x;
  TcXfinLayout L;
  const float *hcat;        // [B][fd x N4] node features x, h_1 .. h_D (feature-major), written by x_net_kernel
  int hcat_stride;          // floats per graph
  uint8_t *img;             // weight operand image (global)
};

#ifdef TC_XFIN_KERNEL_TU
// registers -> TMEM: 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tx_tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier
__device__ __forceinline__ void tx_bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Weight operand image: chunk order = the order the MMA warp consumes them: L1 half 0, L1 half 1, for each column half
// of layer 2 its two k chunks, then layer 3.  Inside a chunk: hi then lo, each [n-blocks][KC k-rows][128 B] MN-major
// SWIZZLE_128B: (n, k) at (n/64)*KC*128 + k*128 + (((n%64)/8) ^ (k%8))*16 + (n%8)*2.
CCSD_KERNEL void __launch_bounds__(256) tc_xfin_prep_kernel(const DevPlan *__restrict__ P, TcXfinLayout T, uint8_t *__restrict__ img) {
  const ccsd_mlp_t &fin = P->d.netx.fin;
  const float *W = P->W;
  for (int ch = blockIdx.y; ch < 7; ch += gridDim.y) {
    const int layer = ch < 2 ? 0 : (ch < 6 ? 1 : 2);
    const int nh = layer == 0 ? ch : (layer == 1 ? (ch - 2) >> 1 : 0), kc = layer == 1 ? (ch - 2) & 1 : 0;
    const int KC = layer == 0 ? T.K1p : (layer == 1 ? T.KC2 : T.Hp), Kin = layer == 0 ? T.fd : T.dh;
    const int NW = layer == 2 ? 16 : T.NH, nblk = layer == 2 ? 1 : T.NHb;      // columns of the chunk, its n-blocks
    const int Nout = layer == 2 ? T.F : T.dh, opad = round_up(Nout, 8);
    const uint32_t cbytes = layer == 0 ? T.chunk1 : (layer == 1 ? T.chunk2 : T.chunk3), half = cbytes / 2;
    uint8_t *dst = img + (layer == 0 ? (size_t)ch * T.chunk1
                                     : 2 * (size_t)T.chunk1 + (layer == 1 ? (size_t)(ch - 2) * T.chunk2 : 4 * (size_t)T.chunk2));
    const int nch8 = nblk * 8;   // 8-column groups (whole n-blocks: the pad groups are zero)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < KC * nch8; t += gridDim.x * blockDim.x) {
      const int k = t / nch8, n0 = (t - k * nch8) << 3;
      const int kg = kc * T.KC2 + k, ng = nh * T.NH + n0;     // global k row / first column of this group
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        x[q] = (kg < Kin && n0 + q < NW && ng + q < Nout) ? __ldg(W + fin.w[layer] + (size_t)kg * opad + ng + q) : 0.f;
      uint4 hi, lo;
      tc::split8(x, hi, lo);
      const uint32_t off = (uint32_t)(n0 >> 6) * ((uint32_t)KC * 128u) + (uint32_t)k * 128u + (uint32_t)((((n0 & 63) >> 3) ^ (k & 7)) << 4);
      *reinterpret_cast<uint4 *>(dst + off) = hi;
      *reinterpret_cast<uint4 *>(dst + half + off) = lo;
    }
  }
}

__global__ void __launch_bounds__(TX_THREADS, 1) tc_xfin_kernel(const DevPlan *__restrict__ P, TcXfinArgs ta) {
  extern __shared__ uint8_t tx_smem_raw[];
  const XaArgs &a = ta.x;
  const TcXfinLayout &T = ta.L;
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_mlp_t &fin = d.netx.fin;
  const int N = d.N, N4 = L.N4, F = d.F, B = d.B;
  const int G = T.G, R = T.R, K1p = T.K1p, dh = T.dh, Hp = T.Hp, NH = T.NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *W = P->W;

  const uint32_t raw = tc::smem_u32(tx_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = tx_smem_raw + (base - raw);
  // barriers: wfull[2], wempty[2], dbar (accumulator complete), xready, aready (512 worker arrivals each), TMEM slot
  const uint32_t bars = base + T.bars, wfull = bars, wempty = bars + 16, dbar = bars + 32, xready = bars + 40, aready = bars + 48,
                 tslot = bars + 56;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + T.bars + 56);
  float *vb1 = reinterpret_cast<float *>(gen + T.vec), *vb2 = vb1 + 256, *vb3 = vb2 + 256;
  float *sc = reinterpret_cast<float *>(gen + T.sc), *sq = reinterpret_cast<float *>(gen + T.sq);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { tc::mbar_init(wfull + 8 * s, 1); tc::mbar_init(wempty + 8 * s, 1); }
    tc::mbar_init(dbar, 1);
    tc::mbar_init(xready, TX_WORK);
    tc::mbar_init(aready, TX_WORK);
    tc::mbar_fence_init();
  }
  if (warp == TX_MMAW) tc::tmem_alloc(tslot, 512);
  // zero the X operand (pad rows / k columns stay zero), stage the biases
  for (uint32_t o = threadIdx.x * 16u; o < T.wb; o += TX_THREADS * 16u) *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 256; i += TX_THREADS) {
    vb1[i] = i < dh ? __ldg(W + fin.b[0] + i) : 0.f;
    vb2[i] = i < dh ? __ldg(W + fin.b[1] + i) : 0.f;
    if (i < 16) vb3[i] = i < F ? __ldg(W + fin.b[2] + i) : 0.f;
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const uint32_t TA = 256;                    // TMEM column of the A operand of layers 2 / 3: hi pairs [TA, TA + Hp/2), lo pairs + Hp/2
  const int ntiles = (B + G - 1) / G;
  const int nmine = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == TX_CPW) {
    // ===================== weight copy warp: 7 chunks per tile through the 2-slot ring =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tl = 0; tl < nmine; ++tl)
        for (int ch = 0; ch < 7; ++ch, ++it) {
          const uint32_t s = it & 1u, use = it >> 1;
          if (use > 0) tc::mbar_wait(wempty + 8 * s, (use - 1) & 1u);     // the MMAs that read this slot have retired
          const uint32_t bytes = ch < 2 ? T.chunk1 : (ch < 6 ? T.chunk2 : T.chunk3);
          const uint8_t *src = ta.img + (ch < 2 ? (size_t)ch * T.chunk1
                                                : 2 * (size_t)T.chunk1 + (ch < 6 ? (size_t)(ch - 2) * T.chunk2 : 4 * (size_t)T.chunk2));
          tc::mbar_arrive_expect_tx(wfull + 8 * s, bytes);
          tx_bulk_load(base + T.wb + s * T.chunk, src, bytes, wfull + 8 * s);
        }
    }
  } else if (warp == TX_MMAW) {
    // ===================== MMA issuer =====================
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc1 = tc::make_idesc_bf16(128, NH, /*A K-major (smem) or TMEM*/ 0, /*B MN-major*/ 1);
    const uint32_t idesc3 = tc::make_idesc_bf16(128, 16, 0, 1);
    uint32_t it = 0, aph = 0;
    for (int tl = 0; tl < nmine; ++tl) {
      tc::mbar_wait(xready, (uint32_t)tl & 1u);
      tc::tc_fence_after_sync();
      for (int nh = 0; nh < 2; ++nh, ++it) {       // layer 1: column halves
        const uint32_t s = it & 1u, use = it >> 1;
        tc::mbar_wait(wfull + 8 * s, use & 1u);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t wb = base + T.wb + s * T.chunk, half = T.chunk1 / 2, blk = (uint32_t)K1p * 128u;
          const uint32_t dcol = tmem_u + (uint32_t)(nh * NH);
          for (int k4 = 0; k4 < K1p / 16; ++k4) {
            const uint32_t ao = (uint32_t)(k4 >> 2) * 16384u + (uint32_t)(k4 & 3) * 32u;
            const uint64_t a_hi = tc::make_smem_desc(base + T.a1 + ao, 0, 1024);
            const uint64_t a_lo = tc::make_smem_desc(base + T.a1 + T.a1_half + ao, 0, 1024);
            const uint64_t b_hi = tc::make_smem_desc(wb + (uint32_t)k4 * 2048u, blk, 1024);
            const uint64_t b_lo = tc::make_smem_desc(wb + half + (uint32_t)k4 * 2048u, blk, 1024);
            tc::umma_bf16(dcol, a_hi, b_hi, idesc1, k4 != 0);
            tc::umma_bf16(dcol, a_hi, b_lo, idesc1, 1);
            tc::umma_bf16(dcol, a_lo, b_hi, idesc1, 1);
          }
          tc::umma_commit(wempty + 8 * s);
          if (nh == 1) tc::umma_commit(dbar);
        }
        __syncwarp();
      }
      tc::mbar_wait(aready, aph); aph ^= 1u;         // elu(D + b1) is in tensor memory as the A operand
      tc::tc_fence_after_sync();
      for (int nh = 0; nh < 2; ++nh)
        for (int kc = 0; kc < 2; ++kc, ++it) {       // layer 2: column halves x k chunks
          const uint32_t s = it & 1u, use = it >> 1;
          tc::mbar_wait(wfull + 8 * s, use & 1u);
          tc::tc_fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t wb = base + T.wb + s * T.chunk, half = T.chunk2 / 2, blk = (uint32_t)T.KC2 * 128u;
            const uint32_t dcol = tmem_u + (uint32_t)(nh * NH);
            for (int k4 = 0; k4 < T.KC2 / 16; ++k4) {
              const uint32_t acol = tmem_u + TA + (uint32_t)(kc * (T.KC2 / 2) + k4 * 8);
              const uint64_t b_hi = tc::make_smem_desc(wb + (uint32_t)k4 * 2048u, blk, 1024);
              const uint64_t b_lo = tc::make_smem_desc(wb + half + (uint32_t)k4 * 2048u, blk, 1024);
              tc::umma_bf16_ts(dcol, acol, b_hi, idesc1, (kc | k4) != 0);
              tc::umma_bf16_ts(dcol, acol, b_lo, idesc1, 1);
              tc::umma_bf16_ts(dcol, acol + (uint32_t)(Hp / 2), b_hi, idesc1, 1);
            }
            tc::umma_commit(wempty + 8 * s);
            if (nh == 1 && kc == 1) tc::umma_commit(dbar);
          }
          __syncwarp();
        }
      tc::mbar_wait(aready, aph); aph ^= 1u;         // elu(D + b2) is in tensor memory
      tc::tc_fence_after_sync();
      {                                              // layer 3: 16 output columns, K = Hp
        const uint32_t s = it & 1u, use = it >> 1;
        ++it;
        tc::mbar_wait(wfull + 8 * s, use & 1u);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t wb = base + T.wb + s * T.chunk, half = T.chunk3 / 2;
          for (int k4 = 0; k4 < Hp / 16; ++k4) {
            const uint32_t acol = tmem_u + TA + (uint32_t)(k4 * 8);
            const uint64_t b_hi = tc::make_smem_desc(wb + (uint32_t)k4 * 2048u, (uint32_t)Hp * 128u, 1024);
            const uint64_t b_lo = tc::make_smem_desc(wb + half + (uint32_t)k4 * 2048u, (uint32_t)Hp * 128u, 1024);
            tc::umma_bf16_ts(tmem_u, acol, b_hi, idesc3, k4 != 0);
            tc::umma_bf16_ts(tmem_u, acol, b_lo, idesc3, 1);
            tc::umma_bf16_ts(tmem_u, acol + (uint32_t)(Hp / 2), b_hi, idesc3, 1);
          }
          tc::umma_commit(wempty + 8 * s);
          tc::umma_commit(dbar);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== workers =====================
    const int r = threadIdx.x & 127, part_id = threadIdx.x >> 7, lq = warp & 3;
    const int gl = r / N, ni = r - gl * N;
    const bool row_in = r < R;
    const uint32_t trow = tmem + ((uint32_t)(lq * 32) << 16);
    uint32_t dph = 0;
    for (int tl = 0; tl < nmine; ++tl) {
      const int tile = (int)blockIdx.x + tl * (int)gridDim.x;
      const int b0 = tile * G, gsz = B - b0 < G ? B - b0 : G;
      const bool live = row_in && gl < gsz;
      const int b = b0 + (live ? gl : 0);
      // ---- X rows -> A1 (K-major): 8 features per 16-byte chunk ----
      {
        const float *hx = ta.hcat + (size_t)b * ta.hcat_stride + ni;
        for (int q8 = part_id; q8 < (K1p >> 3); q8 += TX_NP) {
          float x[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int f = q8 * 8 + q;
            x[q] = (live && f < T.fd) ? hx[(size_t)f * N4] : 0.f;
          }
          uint4 hi, lo;
          tc::split8(x, hi, lo);
          const uint32_t off = T.a1 + (uint32_t)(q8 >> 3) * 16384u + (uint32_t)r * 128u + (uint32_t)(((q8 & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4 *>(gen + off) = hi;
          *reinterpret_cast<uint4 *>(gen + off + T.a1_half) = lo;
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(xready);
      }
      // ---- epilogues 1 / 2: elu(D + b) -> TMEM A operand (packed bf16 pairs: element k in column k / 2) ----
      for (int ly = 0; ly < 2; ++ly) {
        const float *vb = ly == 0 ? vb1 : vb2;
        tc::mbar_wait(dbar, dph); dph ^= 1u;
        tc::tc_fence_after_sync();
        for (int ck = part_id; ck < (Hp >> 4); ck += TX_NP) {
          float v[16];
          tc::tmem_ld16(trow + (uint32_t)(ck * 16), v);
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = ck * 16 + h8 * 8 + q;
              x[q] = c < dh ? fast_elu(v[h8 * 8 + q] + vb[c]) : 0.f;
            }
            uint4 hi, lo;
            tc::split8(x, hi, lo);
            hw[h8 * 4 + 0] = hi.x; hw[h8 * 4 + 1] = hi.y; hw[h8 * 4 + 2] = hi.z; hw[h8 * 4 + 3] = hi.w;
            lw[h8 * 4 + 0] = lo.x; lw[h8 * 4 + 1] = lo.y; lw[h8 * 4 + 2] = lo.z; lw[h8 * 4 + 3] = lo.w;
          }
          tx_tmem_st8(trow + TA + (uint32_t)(ck * 8), hw);
          tx_tmem_st8(trow + TA + (uint32_t)(Hp / 2 + ck * 8), lw);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before_sync();
        tc::mbar_arrive(aready);
      }
      // ---- epilogue 3: the 16 output columns of this row -> shared memory ----
      tc::mbar_wait(dbar, dph); dph ^= 1u;
      tc::tc_fence_after_sync();
      if (part_id == 0) {
        float v[16];
        tc::tmem_ld16(trow, v);
#pragma unroll
        for (int f4 = 0; f4 < 4; ++f4)
          *reinterpret_cast<float4 *>(sc + r * 16 + 4 * f4) = make_float4(v[4 * f4] + vb3[4 * f4], v[4 * f4 + 1] + vb3[4 * f4 + 1],
                                                                        v[4 * f4 + 2] + vb3[4 * f4 + 2], v[4 * f4 + 3] + vb3[4 * f4 + 3]);
      }
      tc::tc_fence_before_sync();
      asm volatile("bar.sync 1, 512;" ::: "memory");   // the tile's network output is in `sc`
      // ---- x sampler epilogue: item = (row, feature) of the tile's live rows ----
      const int stp = a.mode == MODE_EVAL ? 0 : nz_step(a.nz);
      const ccsd_objcoef_t cx = a.mode == MODE_EVAL ? ccsd_objcoef_t() : P->sched[stp * 3 + 0];
      const int nitem = gsz * N * F;
      for (int it2 = threadIdx.x; it2 < nitem; it2 += TX_WORK) {
        const int rr = it2 / F, f = it2 - rr * F;          // rr = row of the tile (graph g2, node i)
        const int g2 = rr / N, i = rr - g2 * N, bb = b0 + g2;
        const float fl = a.flags[(size_t)bb * N + i];
        const float o = sc[rr * 16 + f] * fl;               // mask_x
        const size_t gp = ((size_t)bb * N + i) * F + f;
        const int p = i * F + f;
        if (a.mode == MODE_EVAL) { a.out_x[gp] = o; continue; }
        const unsigned long long gsid = (unsigned long long)(a.nz.sample_offset + bb);
        const float s = cx.score_scale * o;
        const float z = (a.noise_x ? a.noise_x[gp] : normal1(a.nz.seed, gsid, draw_id(0, stp, a.slot), p)) * fl;
        if (a.mode == MODE_SCORE) {
          a.out_x[gp] = s;
          sq[it2] = s * s;                 // per-sample norms are summed below in a fixed order (bit-reproducible)
          sq[128 * 16 + it2] = z * z;
        } else {
          const float m = cx.pa * a.x[gp] + cx.pb * s;
          const float v = m + cx.pc * z;
          a.out_x[gp] = v;
          a.mean_x[gp] = m;
          if (bb == 0) { float *tjx = a.nz.sd ? a.nz.sd->tx : a.traj_x; if (tjx) tjx[p] = a.denoise ? m : v; }
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");   // `sc` may be overwritten by the next tile; squared entries complete
      if (a.mode == MODE_SCORE && (int)threadIdx.x < gsz) {
        float ts = 0.f, tz = 0.f;
        const int i0 = (int)threadIdx.x * N * F;
        for (int q = 0; q < N * F; ++q) { ts += sq[i0 + q]; tz += sq[128 * 16 + i0 + q]; }
        float *np = a.norm_part + ((size_t)(0 * d.B + b0 + (int)threadIdx.x) * P->ntile_max) * 2;
        np[0] = ts; np[1] = tz;
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TX_MMAW) tc::tmem_dealloc(tmem, 512);
}

int tc_xfin_prep(const DevPlan *dP, const TcXfinLayout &T, uint8_t *img, void *stream) {
  tc_xfin_prep_kernel<<<dim3(8, 7, 1), 256, 0, (cudaStream_t)stream>>>(dP, T, img);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int tc_xfin_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, const TcXfinLayout &T, const float *hcat, int hcat_stride,
                   uint8_t *img, void *stream) {
  TcXfinArgs ta;
  ta.x = a; ta.L = T; ta.hcat = hcat; ta.hcat_stride = hcat_stride; ta.img = img;
  static CcsdSmemAttr attr;
  if (ccsd_ensure_smem(tc_xfin_kernel, T.total, attr)) return -1;
  const int ntiles = (hp.d.B + T.G - 1) / T.G;
  tc_xfin_kernel<<<ntiles < 148 ? ntiles : 148, TX_THREADS, T.total, (cudaStream_t)stream>>>(dP, ta);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#else
// defined in tc_xfin_tu.cu (its own translation unit)
int tc_xfin_prep(const DevPlan *dP, const TcXfinLayout &T, uint8_t *img, void *stream);
int tc_xfin_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, const TcXfinLayout &T, const float *hcat, int hcat_stride,
                   uint8_t *img, void *stream);
#endif  // TC_XFIN_KERNEL_TU

}  // namespace ccsd
