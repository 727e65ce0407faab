// tc_r2big.cuh -- the two rank-2 contractions of ScoreNetworkF for LARGE complexes (E > 192: grid_small_CC has E = 1176,
// K = 18424, 87 MB of rank-2 state per sample) as K-chunked tcgen05 GEMMs:
//
//   MODE 0 (Gram)   G[b] = F[b] . [F[b] ; Wp]^T     H = G[:, :E] (1 - I)  (cc_utils.py:917-942, 964-969),  P0 = G[:, E:]
//                   output tiles 128 x 256; only tiles that reach the diagonal are computed (H is symmetric), every
//                   element c >= e is written once and mirrored, so H is exactly symmetric and bit-reproducible
//   MODE 1 (H . F)  HF[b] = H[b] . F[b]              (E x E) . (E x K), tiles 128 edges x 256 cells, written to a scratch
//                   tensor that r2_epi_kernel (per-entry network + sampler update) consumes
//
// Both are real GEMMs (2 E^2 K = 51 GFLOP per sample and product at grid_small_CC): persistent CTAs walk (sample, M tile,
// N tile) units; 16 producer warps stream 64-wide k chunks of both operands from global memory (fp32, coalesced), split
// them into bf16 hi / lo (bf16x3: hi.hi + hi.lo + lo.hi with fp32 accumulation keeps the 1e-4 parity bar) and store them
// in the canonical SWIZZLE_128B layouts (A K-major; B K-major for the Gram product, MN-major -- the state's own layout --
// for H . F); one elected thread issues the 12 MMAs of a chunk (M = 128, N = 256, K = 16); the accumulators are double
// buffered in tensor memory (2 x 256 columns) so the epilogue of one unit overlaps the MMAs of the next.
#pragma once
#include "r2_kernels.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TB_PROD_WARPS = 16;
constexpr int TB_PROD = TB_PROD_WARPS * 32;
constexpr int TB_THREADS = TB_PROD + 128 + 32;     // producers, 4 epilogue warps, 1 MMA warp
constexpr int TB_BK = 64;
constexpr uint32_t TB_A = 128u * 128u;             // A part of a stage half: 128 rows x 128 B
constexpr uint32_t TB_B = 256u * 128u;             // B part: 256 rows (K-major) or 4 n-blocks x 64 k rows (MN-major)
constexpr uint32_t TB_HALF = TB_A + TB_B;          // hi (or lo) half of a stage
constexpr uint32_t TB_STAGE = 2u * TB_HALF;        // 96 KB
constexpr int TB_STAGES = 2;
constexpr size_t TB_SMEM = (size_t)TB_STAGES * TB_STAGE + 1024 + 256;
constexpr int TB_MAX_TILES = 192;

struct TcBigArgs {
  const float *r2;        // MODE 0: the matrix X whose Gram product is taken, [B][E][ldx] (the rank-2 state F, or H for H . H);
                          // MODE 1: F [B,E,K]
  int Kx, ldx;            // MODE 0: contraction length and row pitch of X
  int PRx;                // MODE 0: projection rows of the weight blob appended as Gram columns (0 for H . H)
  int mask_diag;          // MODE 0: zero the diagonal of the output
  float *H;               // [B,E,Ep]      (MODE 0: written; MODE 1: read)
  float *P0;              // [B,E,PRx]     (MODE 0)
  float *Dg, *Rs;         // MODE 0, optional [B,E]: diag(X X^T) before the mask; row sums X 1 (one more all-ones Gram column)
  float *hf;              // [B,E,K]       (MODE 1: written)
  int ntile;              // (M tile, N tile) pairs per sample
  int mt;                 // MODE 1: M tiles per sample; unit tl -> (mi = tl % mt, nj = tl / mt), so that consecutive units
                          // share the same 256 cells of F in L2
  unsigned short tile[TB_MAX_TILES];   // MODE 0: (mi << 8) | nj of the tiles that reach the diagonal
};

#ifdef TC_R2BIG_KERNEL_TU
template <int MODE>
__global__ void __launch_bounds__(TB_THREADS, 1) tc_r2big_kernel(const DevPlan *__restrict__ P, TcBigArgs a) {
  extern __shared__ uint8_t tb_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const int E = d.E, K = MODE == 0 ? a.Kx : d.K, PR0 = MODE == 0 ? a.PRx : 0, Kw = P->Kp, Ep = P->Ep, B = d.B;
  const int ldx = MODE == 0 ? a.ldx : d.K;                    // row pitch of the MODE 0 input
  const int Ec0 = (E + 7) & ~7;                               // first projection column of the Gram product
  const int rsum = MODE == 0 && a.Rs != nullptr;              // all-ones operand row at Gram column Ec0 + PR0
  const int nkb = MODE == 0 ? (K + TB_BK - 1) / TB_BK : (E + TB_BK - 1) / TB_BK;
  const int nunits = B * a.ntile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t raw = tc::smem_u32(tb_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = tb_smem_raw + (base - raw);
  const uint32_t bars = base + TB_STAGES * TB_STAGE;
  const uint32_t full0 = bars, empty0 = bars + 8 * TB_STAGES, tfull0 = bars + 16 * TB_STAGES, tempty0 = tfull0 + 16, tslot = tempty0 + 16;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TB_STAGES * TB_STAGE + 16 * TB_STAGES + 32);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TB_STAGES; ++s) { tc::mbar_init(full0 + 8 * s, TB_PROD); tc::mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(tfull0 + 8 * s, 1); tc::mbar_init(tempty0 + 8 * s, 128); }
    tc::mbar_fence_init();
  }
  if (warp == TB_PROD_WARPS + 4) tc::tmem_alloc(tslot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const int nmine = ((int)blockIdx.x < nunits) ? (nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp < TB_PROD_WARPS) {
    // ===================== producers =====================
    const float *Wp = P->W + d.neta.proj_w;
    const bool vecK = MODE == 0 ? ((ldx & 3) == 0) : ((K & 3) == 0);
    const int t = threadIdx.x;
    long it = 0;
    for (int ul = 0; ul < nmine; ++ul) {
      const int u = (int)blockIdx.x + ul * (int)gridDim.x;
      const int b = u / a.ntile, tl = u - b * a.ntile;
      const int mi = MODE == 0 ? a.tile[tl] >> 8 : tl % a.mt, nj = MODE == 0 ? a.tile[tl] & 255 : tl / a.mt;
      const float *Fb = a.r2 + (size_t)b * E * ldx;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        float x[6][8];
        if (MODE == 0) {
          // rows r0 + 64 j: j < 2 -> A rows (edges mi*128 + r), j >= 2 -> B rows (Gram columns nj*256 + r - 128)
          const int c = t & 7, r0 = t >> 3;
          const int k = kb * TB_BK + c * 8;
          const bool fast = vecK && (k + 8 <= K) && ((reinterpret_cast<uintptr_t>(Fb) & 15) == 0);
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int rr = r0 + 64 * j;
            const float *src = nullptr;
            int klim = K;
            bool ones = false;
            if (j < 2) { const int e = mi * 128 + rr; if (e < E) src = Fb + (size_t)e * ldx; }
            else {
              const int cc = nj * 256 + rr - 128;
              if (cc < E) src = Fb + (size_t)cc * ldx;
              else if (cc >= Ec0 && cc - Ec0 < PR0) { src = Wp + (size_t)(cc - Ec0) * Kw; klim = Kw < K ? Kw : K; }
              else if (rsum && cc == Ec0 + PR0) ones = true;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) x[j][q] = (ones && k + q < K) ? 1.f : 0.f;
            if (src) {
              if (fast) {
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src + k)), v1 = __ldg(reinterpret_cast<const float4 *>(src + k + 4));
                x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w; x[j][4] = v1.x; x[j][5] = v1.y; x[j][6] = v1.z; x[j][7] = v1.w;
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) x[j][q] = (k + q < klim) ? __ldg(src + k + q) : 0.f;
              }
            }
          }
        } else {
          // A: H rows (edges mi*128 + r0 + 64 j, j < 2), 8 columns e' per item; B: F rows e' = kb*64 + kr, 8 cells per item
          {
            const int c = t & 7, r0 = t >> 3;
            const int k = kb * TB_BK + c * 8;          // e' column
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int e = mi * 128 + r0 + 64 * j;
#pragma unroll
              for (int q = 0; q < 8; ++q) x[j][q] = 0.f;
              if (e < E && k < E) {
                const float *src = a.H + ((size_t)b * E + e) * Ep + k;
                if (k + 8 <= Ep) {
                  const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src)), v1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
                  x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w; x[j][4] = v1.x; x[j][5] = v1.y; x[j][6] = v1.z; x[j][7] = v1.w;
                } else {
#pragma unroll
                  for (int q = 0; q < 8; ++q) x[j][q] = (k + q < Ep) ? __ldg(src + q) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (k + q >= E) x[j][q] = 0.f;
              }
            }
          }
          {
            const int cn = t & 31, kr0 = t >> 5;
            const int cell = nj * 256 + cn * 8;
            const bool fast = vecK && (cell + 8 <= K);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ep = kb * TB_BK + kr0 + 16 * j;   // e' row of F
#pragma unroll
              for (int q = 0; q < 8; ++q) x[2 + j][q] = 0.f;
              if (ep < E && cell < K) {
                const float *src = Fb + (size_t)ep * K + cell;
                if (fast) {
                  const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src)), v1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
                  x[2 + j][0] = v0.x; x[2 + j][1] = v0.y; x[2 + j][2] = v0.z; x[2 + j][3] = v0.w;
                  x[2 + j][4] = v1.x; x[2 + j][5] = v1.y; x[2 + j][6] = v1.z; x[2 + j][7] = v1.w;
                } else {
#pragma unroll
                  for (int q = 0; q < 8; ++q) x[2 + j][q] = (cell + q < K) ? __ldg(src + q) : 0.f;
                }
              }
            }
          }
        }
        const int s = (int)(it % TB_STAGES);
        const uint32_t ph = (uint32_t)(it / TB_STAGES) & 1u;
        tc::mbar_wait(empty0 + 8 * s, ph ^ 1u);
        uint8_t *st = gen + (size_t)s * TB_STAGE;
        if (MODE == 0) {
          const int c = t & 7, r0 = t >> 3;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int rr = j < 2 ? r0 + 64 * j : r0 + 64 * j - 128;      // row inside the A / B part
            const uint32_t part = j < 2 ? 0u : TB_A;
            const uint32_t off = part + (uint32_t)(rr >> 3) * 1024u + (uint32_t)(rr & 7) * 128u + (uint32_t)((c ^ (rr & 7)) << 4);
            uint4 hi, lo;
            tc::split8(x[j], hi, lo);
            *reinterpret_cast<uint4 *>(st + off) = hi;
            *reinterpret_cast<uint4 *>(st + TB_HALF + off) = lo;
          }
        } else {
          {
            const int c = t & 7, r0 = t >> 3;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int rr = r0 + 64 * j;
              const uint32_t off = (uint32_t)(rr >> 3) * 1024u + (uint32_t)(rr & 7) * 128u + (uint32_t)((c ^ (rr & 7)) << 4);
              uint4 hi, lo;
              tc::split8(x[j], hi, lo);
              *reinterpret_cast<uint4 *>(st + off) = hi;
              *reinterpret_cast<uint4 *>(st + TB_HALF + off) = lo;
            }
          }
          {
            const int cn = t & 31, kr0 = t >> 5, n0 = cn * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int kr = kr0 + 16 * j;
              // MN-major B: (n, k) at (n/64)*8192 + k*128 + (((n%64)/8) ^ (k%8))*16
              const uint32_t off = TB_A + (uint32_t)(n0 >> 6) * 8192u + (uint32_t)kr * 128u + (uint32_t)((((n0 & 63) >> 3) ^ (kr & 7)) << 4);
              uint4 hi, lo;
              tc::split8(x[2 + j], hi, lo);
              *reinterpret_cast<uint4 *>(st + off) = hi;
              *reinterpret_cast<uint4 *>(st + TB_HALF + off) = lo;
            }
          }
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(full0 + 8 * s);
      }
    }
  } else if (warp == TB_PROD_WARPS + 4) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::make_idesc_bf16(128, 256, 0, MODE == 0 ? 0 : 1);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t it = 0;
    for (int ul = 0; ul < nmine; ++ul) {
      const uint32_t ab = (uint32_t)ul & 1u, use = (uint32_t)ul >> 1;
      tc::mbar_wait(tempty0 + 8 * ab, (use & 1u) ^ 1u);          // the epilogue has drained this accumulator buffer
      tc::tc_fence_after_sync();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % TB_STAGES;
        const uint32_t ph = (it / TB_STAGES) & 1u;
        tc::mbar_wait(full0 + 8 * s, ph);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t sb = base + (uint32_t)s * TB_STAGE;
          const uint32_t dcol = tmem_u + ab * 256u;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t a_hi = tc::make_smem_desc(sb + (uint32_t)k4 * 32u, 0, 1024);
            const uint64_t a_lo = tc::make_smem_desc(sb + TB_HALF + (uint32_t)k4 * 32u, 0, 1024);
            uint64_t b_hi, b_lo;
            if (MODE == 0) {
              b_hi = tc::make_smem_desc(sb + TB_A + (uint32_t)k4 * 32u, 0, 1024);
              b_lo = tc::make_smem_desc(sb + TB_HALF + TB_A + (uint32_t)k4 * 32u, 0, 1024);
            } else {
              b_hi = tc::make_smem_desc(sb + TB_A + (uint32_t)k4 * 2048u, 8192, 1024);
              b_lo = tc::make_smem_desc(sb + TB_HALF + TB_A + (uint32_t)k4 * 2048u, 8192, 1024);
            }
            tc::umma_bf16(dcol, a_hi, b_hi, idesc, (kb | k4) != 0);
            tc::umma_bf16(dcol, a_hi, b_lo, idesc, 1);
            tc::umma_bf16(dcol, a_lo, b_hi, idesc, 1);
          }
          tc::umma_commit(empty0 + 8 * s);
          if (kb == nkb - 1) tc::umma_commit(tfull0 + 8 * ab);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 16-19 -> TMEM lane quarters 0-3) =====================
    const int q = warp & 3;
    const int mask_diag = MODE == 0 ? a.mask_diag : 0;
    const bool vecK = (K & 3) == 0;
    for (int ul = 0; ul < nmine; ++ul) {
      const int u = (int)blockIdx.x + ul * (int)gridDim.x;
      const int b = u / a.ntile, tl = u - b * a.ntile;
      const int mi = MODE == 0 ? a.tile[tl] >> 8 : tl % a.mt, nj = MODE == 0 ? a.tile[tl] & 255 : tl / a.mt;
      const uint32_t ab = (uint32_t)ul & 1u, use = (uint32_t)ul >> 1;
      tc::mbar_wait(tfull0 + 8 * ab, use & 1u);
      tc::tc_fence_after_sync();
      const int e = mi * 128 + q * 32 + lane;           // this thread's edge row
      const int e_lo = mi * 128 + q * 32;               // first row of the warp
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + ab * 256u;
      for (int c0 = 0; c0 < 256; c0 += 16) {
        const int cb = nj * 256 + c0;
        if (MODE == 0) {
          // warp-uniform skips: columns entirely below the warp's rows (their mirrors are written by another tile),
          // or past everything
          if ((cb + 15 < e_lo && cb + 16 <= Ec0) || cb >= Ec0 + PR0 + rsum) continue;
        } else {
          if (cb >= K) continue;
        }
        float v[16];
        tc::tmem_ld16(trow + (uint32_t)c0, v);
        if (e >= E) continue;
        if (MODE == 0) {
          float *Hb = a.H + (size_t)b * E * Ep;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = cb + j;
            if (c < E) {
              if (c >= e) {
                const float val = (mask_diag && c == e) ? 0.f : v[j];
                Hb[(size_t)e * Ep + c] = val;
                if (c > e) Hb[(size_t)c * Ep + e] = val;
                if (c == e && a.Dg) a.Dg[(size_t)b * E + e] = v[j];
              }
            } else if (c >= Ec0 && c - Ec0 < PR0) {
              a.P0[((size_t)b * E + e) * PR0 + (c - Ec0)] = v[j];
            } else if (rsum && c == Ec0 + PR0) {
              a.Rs[(size_t)b * E + e] = v[j];
            }
          }
        } else {
          float *dst = a.hf + ((size_t)b * E + e) * K + cb;
          if (vecK && cb + 16 <= K) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) *reinterpret_cast<float4 *>(dst + 4 * j4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cb + j < K) dst[j] = v[j];
          }
        }
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(tempty0 + 8 * ab);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TB_PROD_WARPS + 4) tc::tmem_dealloc(tmem, 512);
}

template <int MODE>
static int tc_r2big_launch_m(const DevPlan *dP, const DevPlan &hp, TcBigArgs &a, void *stream) {
  static CcsdSmemAttr attr;
  if (ccsd_ensure_smem(tc_r2big_kernel<MODE>, TB_SMEM, attr)) return -1;
  const int nunits = hp.d.B * a.ntile;
  tc_r2big_kernel<MODE><<<nunits < 148 ? nunits : 148, TB_THREADS, TB_SMEM, (cudaStream_t)stream>>>(dP, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Gram product of X [B][E][ldx] (contraction length Kx): out = X X^T (diagonal zeroed when mask_diag), optionally PRx
// projection columns (P0), the diagonal (Dg) and the row sums (Rs)
int tc_r2big_gram_x(const DevPlan *dP, const DevPlan &hp, const float *X, int Kx, int ldx, int PRx, int mask_diag, float *out,
                    float *P0, float *Dg, float *Rs, void *stream) {
  TcBigArgs a;
  memset(&a, 0, sizeof a);
  a.r2 = X; a.Kx = Kx; a.ldx = ldx; a.PRx = PRx; a.mask_diag = mask_diag; a.H = out; a.P0 = P0; a.Dg = Dg; a.Rs = Rs;
  const int E = hp.d.E, Ec0 = (E + 7) & ~7, ctot = Ec0 + PRx + (Rs ? 1 : 0);
  const int mt = (E + 127) / 128, nt = (ctot + 255) / 256;
  int n = 0;
  for (int mi = 0; mi < mt; ++mi)
    for (int nj = 0; nj < nt; ++nj)
      if (nj * 256 + 255 >= mi * 128 || nj * 256 + 256 > Ec0) {   // reaches the diagonal, or holds projection columns
        if (n >= TB_MAX_TILES) return -1;
        a.tile[n++] = (unsigned short)((mi << 8) | nj);
      }
  a.ntile = n;
  return tc_r2big_launch_m<0>(dP, hp, a, stream);
}

int tc_r2big_gram(const DevPlan *dP, const DevPlan &hp, const float *r2, float *H, float *P0, float *Dg, float *Rs, void *stream) {
  return tc_r2big_gram_x(dP, hp, r2, hp.d.K, hp.d.K, hp.PR0, hp.d.netf.use_hodge_mask, H, P0, Dg, Rs, stream);
}

int tc_r2big_hf(const DevPlan *dP, const DevPlan &hp, const float *r2, const float *H, float *hf, void *stream) {
  TcBigArgs a;
  memset(&a, 0, sizeof a);
  a.r2 = r2; a.H = const_cast<float *>(H); a.hf = hf;
  a.mt = (hp.d.E + 127) / 128;
  a.ntile = a.mt * ((hp.d.K + 255) / 256);
  return tc_r2big_launch_m<1>(dP, hp, a, stream);
}
#else
int tc_r2big_gram_x(const DevPlan *dP, const DevPlan &hp, const float *X, int Kx, int ldx, int PRx, int mask_diag, float *out,
                    float *P0, float *Dg, float *Rs, void *stream);
int tc_r2big_gram(const DevPlan *dP, const DevPlan &hp, const float *r2, float *H, float *P0, float *Dg, float *Rs, void *stream);
int tc_r2big_hf(const DevPlan *dP, const DevPlan &hp, const float *r2, const float *H, float *hf, void *stream);
#endif

static inline int tc_r2big_supported(const ccsd_plan_desc_t &d, int PR0) {
  const int mt = (d.E + 127) / 128, nt = (((d.E + 7) & ~7) + PR0 + 1 + 255) / 256;
  return d.is_cc && d.E > 192 && mt * nt <= TB_MAX_TILES && mt <= 255 && (d.K + 255) / 256 <= 65535 && PR0 <= 64;
}

}  // namespace ccsd
