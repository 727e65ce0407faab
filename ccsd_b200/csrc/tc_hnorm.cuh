// tc_hnorm.cuh -- the Langevin norms of the rank-2 object WITHOUT a pass over the rank-2 state.
//
// LangevinCorrector (solver.py:763-767) needs ||score||_2 per sample before any entry can be corrected, which cost a
// whole read of the state (the NORM pass of tc_apply: H.F on the tensor cores, the network, and the noise, only to be
// reduced to two numbers).  When ScoreNetworkF is affine (every shipped CC checkpoint but the Base_CC ablations:
// score = sc m (a f + b (H f) + c), packer.py) the norm follows from E x E Gram quantities of the sample:
//   sum f^2       = sum_e D_e                       D = diag(F F^T)                (tc_gram, before the (1 - I) mask)
//   sum f (H f)   = ||H||_F^2 = sum_e (H^2)_ee      H = (F F^T)(1 - I), symmetric
//   sum (H f)^2   = tr(H G H) = sum_ee' (H^2)_ee' H_ee' + sum_e D_e (H^2)_ee      G = H + diag(D)
//   sum f         = sum_e r_e                       r = F 1  (one more Gram column, tc_gram)
//   sum (H f)     = sum_e (H r)_e
//   sum m         = C(n, 2) * sum_d C(n, d)         n = active nodes (f and H f vanish outside the mask)
//   ||score||^2   = sc^2 (a^2 sum f^2 + b^2 sum (H f)^2 + 2ab sum f (H f) + 2ac sum f + 2bc sum (H f) + c^2 sum m)
// The only O(E^3) piece is H^2 = H . H: one tcgen05 product per sample (bf16x3; its term is a few per cent of the norm, so
// the 2^-16 product error is ~1e-6 of the result), A and B both read the SAME K-major operand (H is symmetric).  Small
// complexes are stacked G per work unit as a block-diagonal operand (blockdiag(H)^2 = blockdiag(H^2)).
// ||z||^2 of the corrector's noise is a Philox-only reduction (znorm_kernel, r2_kernels.cuh).
#pragma once
#include "r2_kernels.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TH_WORK = 256;                 // 8 worker warps: warps 0-3 own M tile 0 (rows 0-127), 4-7 M tile 1 (128-255)
constexpr int TH_THREADS = TH_WORK + 32;     // + the MMA-issuing warp
constexpr int TH_MMAW = TH_WORK / 32;
constexpr uint32_t TH_KB = 256u * 128u;      // bytes of one 64-wide k-block of the operand: 256 rows x 128 B
constexpr uint32_t TH_HALF = 3u * TH_KB;     // hi (or lo) half: 3 k-blocks (K <= 192)
constexpr uint32_t TH_PART = 2u * TH_HALF;   // per-row partials [256][4]
constexpr uint32_t TH_RS = TH_PART + 256 * 4 * 4;     // row sums r of the unit [256]
constexpr uint32_t TH_DG = TH_RS + 256 * 4;           // diag D of the unit [256]
constexpr uint32_t TH_BARS = TH_DG + 256 * 4;
constexpr size_t TH_SMEM = (size_t)TH_BARS + 64 + 1024;

struct TcHnormArgs {
  const float *H;        // [B][E][Ep]
  const float *Dg, *Rs;  // [B][E]
  const float *flags;    // [B][N]
  float *norm_part;      // [3][B][ntile_max][2]: writes slot 0 of the rank-2 object's ||score||^2, zeroes the other score slots
  int step;
  int G;                 // samples per work unit
};

#ifdef TC_HNORM_KERNEL_TU
__global__ void __launch_bounds__(TH_THREADS, 1) tc_hnorm_kernel(const DevPlan *__restrict__ P, TcHnormArgs a) {
  extern __shared__ uint8_t th_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const int E = d.E, Ep = P->Ep, B = d.B, N = d.N, G = a.G;
  const int EB = G * E;                        // rows of a full unit (<= 192)
  const int NC = (EB + 15) & ~15;              // MMA N (and K extent)
  const int nk = NC >> 4;
  const int mtiles = EB > 128 ? 2 : 1;
  const int nunits = (B + G - 1) / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t raw = tc::smem_u32(th_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = th_smem_raw + (base - raw);
  const uint32_t bar = base + TH_BARS, tslot = bar + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TH_BARS + 8);
  float *part = reinterpret_cast<float *>(gen + TH_PART), *rs = reinterpret_cast<float *>(gen + TH_RS), *dg = reinterpret_cast<float *>(gen + TH_DG);

  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (warp == TH_MMAW) tc::tmem_alloc(tslot, 512);
  for (uint32_t o = threadIdx.x * 16u; o < TH_PART; o += TH_THREADS * 16u) *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t idesc = tc::make_idesc_bf16(128, NC, 0, 0);
  const ccsd_netf_t &Fn = d.netf;
  const float fa = Fn.aff[0], fb = Fn.aff[1], fc = Fn.aff[2];
  const float sc = P->sched[a.step * 3 + 2].score_scale;
  uint32_t phase = 0;

  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int b0 = u * G, gsz = B - b0 < G ? B - b0 : G;
    if (warp < TH_MMAW) {
      // ---- operand: row (g, e) holds H_g[e, :] in columns [g E, g E + E), zeros elsewhere; 8 columns per item ----
      if (G == 1) {
        // one sample: thread t owns 16-byte chunk c = t & 7 of every k-block of rows (t >> 3) + 32 j; all 18 row loads
        // (two 16-byte loads each, rows are 16-byte aligned: Ep % 4 == 0) are independent and issued before the splits
        const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;
        const float *Hb = a.H + (size_t)b0 * E * Ep;
#pragma unroll 1
        for (int kb = 0; kb < 3; ++kb) {
          const int col0 = kb * 64 + c * 8;
          if (col0 >= NC) break;
          float x[6][8];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int row = r0 + 32 * j;
#pragma unroll
            for (int q = 0; q < 8; ++q) x[j][q] = 0.f;
            if (row < E) {
              const float *Hr = Hb + (size_t)row * Ep + col0;
              if (col0 + 8 <= Ep) {
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(Hr)), v1 = __ldg(reinterpret_cast<const float4 *>(Hr + 4));
                x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w; x[j][4] = v1.x; x[j][5] = v1.y; x[j][6] = v1.z; x[j][7] = v1.w;
              } else if (col0 + 4 <= Ep) {
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(Hr));
                x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w;
              }
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (col0 + q >= E) x[j][q] = 0.f;       // pitch padding of the H buffer
            }
          }
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int row = r0 + 32 * j;
            if (row < EB) {
              uint4 hi, lo;
              tc::split8(x[j], hi, lo);
              const uint32_t off = (uint32_t)kb * TH_KB + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
              *reinterpret_cast<uint4 *>(gen + off) = hi;
              *reinterpret_cast<uint4 *>(gen + TH_HALF + off) = lo;
            }
          }
        }
      } else {
        const int c8n = NC >> 3;                                // 16-byte chunks per row
        for (int it = threadIdx.x; it < EB * c8n; it += TH_WORK) {
          const int row = it / c8n, c8 = it - row * c8n;
          const int g = row / E, e = row - g * E, k0 = g * E;
          const int col0 = c8 * 8;
          if (col0 + 8 <= k0 || col0 >= k0 + E) continue;       // outside the diagonal block: stays zero
          float x[8];
          const float *Hr = a.H + ((size_t)(b0 + (g < gsz ? g : 0)) * E + e) * Ep;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = col0 + q - k0;
            x[q] = (g < gsz && j >= 0 && j < E) ? __ldg(Hr + j) : 0.f;
          }
          uint4 hi, lo;
          tc::split8(x, hi, lo);
          const uint32_t off = (uint32_t)(col0 >> 6) * TH_KB + (uint32_t)row * 128u + (uint32_t)((((col0 & 63) >> 3) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4 *>(gen + off) = hi;
          *reinterpret_cast<uint4 *>(gen + TH_HALF + off) = lo;
        }
      }
      for (int row = threadIdx.x; row < 256; row += TH_WORK) {
        const int g = row / E, e = row - g * E;
        const bool on = row < EB && g < gsz;
        rs[row] = on ? a.Rs[(size_t)(b0 + g) * E + e] : 0.f;
        dg[row] = on ? a.Dg[(size_t)(b0 + g) * E + e] : 0.f;
      }
      tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    // ---- H^2 = H . H (both operands K-major, the same buffer) ----
    if (warp == TH_MMAW) {
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        for (int mt = 0; mt < mtiles; ++mt)
          for (int k4 = 0; k4 < nk; ++k4) {
            const uint32_t ko = (uint32_t)(k4 >> 2) * TH_KB + (uint32_t)(k4 & 3) * 32u;
            const uint64_t a_hi = tc::make_smem_desc(base + ko + (uint32_t)mt * 16384u, 0, 1024);
            const uint64_t a_lo = tc::make_smem_desc(base + TH_HALF + ko + (uint32_t)mt * 16384u, 0, 1024);
            const uint64_t b_hi = tc::make_smem_desc(base + ko, 0, 1024);
            const uint64_t b_lo = tc::make_smem_desc(base + TH_HALF + ko, 0, 1024);
            const uint32_t dcol = tmem_u + (uint32_t)(mt * 192);
            tc::umma_bf16(dcol, a_hi, b_hi, idesc, k4 != 0);
            tc::umma_bf16(dcol, a_hi, b_lo, idesc, 1);
            tc::umma_bf16(dcol, a_lo, b_hi, idesc, 1);
          }
        tc::umma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    tc::tc_fence_after_sync();
    // ---- per-row sums: t3 = sum_e' (H^2)_ee' H_ee', d2 = (H^2)_ee, hr = (H r)_e ----
    if (warp < TH_MMAW) {
      const int mt = warp >> 2, row = mt * 128 + (warp & 3) * 32 + lane;
      float t3 = 0.f, d2 = 0.f, hr = 0.f;
      if (mt < mtiles) {
        const int g = row / E, e = row - g * E, k0 = g * E;
        const bool on = row < EB && g < gsz;
        // warp-uniform range of 16-column chunks that hold a diagonal block of the warp's 32 rows
        const int rlo = mt * 128 + (warp & 3) * 32, rhi = rlo + 31 < EB - 1 ? rlo + 31 : EB - 1;
        const int clo = rlo < EB ? ((rlo / E) * E) >> 4 : 0, chi = rlo < EB ? ((rhi / E) * E + E + 15) >> 4 : 0;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * 192);
        (void)e;
        for (int ck = clo; ck < chi && ck < nk; ++ck) {
          float v[16];
          tc::tmem_ld16(trow + (uint32_t)(ck * 16), v);
          if (on) {
            // H[row][16 ck .. 16 ck + 15] back from the operand: hi + lo reproduces it to 2^-17 (its own row: conflict free)
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int col0 = ck * 16 + h8 * 8;
              const uint32_t off = (uint32_t)(col0 >> 6) * TH_KB + (uint32_t)row * 128u + (uint32_t)((((col0 & 63) >> 3) ^ (row & 7)) << 4);
              const uint4 hi = *reinterpret_cast<const uint4 *>(gen + off), lo = *reinterpret_cast<const uint4 *>(gen + TH_HALF + off);
              const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const uint32_t hb = (q & 1) ? (hw[q >> 1] & 0xFFFF0000u) : (hw[q >> 1] << 16);
                const uint32_t lb = (q & 1) ? (lw[q >> 1] & 0xFFFF0000u) : (lw[q >> 1] << 16);
                const float h = __uint_as_float(hb) + __uint_as_float(lb);       // zero outside the row's diagonal block
                const int col = col0 + q;
                t3 += v[h8 * 8 + q] * h;
                hr += h * rs[col < 256 ? col : 0];
                if (col == row) d2 = v[h8 * 8 + q];
              }
            }
          }
        }
        (void)k0;
      }
      float4 *pp = reinterpret_cast<float4 *>(part + (size_t)(mt * 128 + (warp & 3) * 32 + lane) * 4);
      *pp = make_float4(t3, d2, hr, 0.f);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    // ---- per-sample reduction in a fixed order, the norm, and the partial slots coef_kernel sums ----
    if ((int)threadIdx.x < gsz) {
      const int g = threadIdx.x, b = b0 + g;
      float Sff = 0.f, Sfh = 0.f, Shh = 0.f, Sf = 0.f, Sh = 0.f;
      for (int e = 0; e < E; ++e) {
        const int row = g * E + e;
        const float4 p4 = *reinterpret_cast<const float4 *>(part + (size_t)row * 4);
        Sff += dg[row]; Sfh += p4.y; Shh += p4.x + dg[row] * p4.y; Sf += rs[row]; Sh += p4.z;
      }
      int n = 0;
      for (int i = 0; i < N; ++i) n += a.flags[(size_t)b * N + i] != 0.f;
      float nkc = 0.f;                                   // active cells: sum_d C(n, d)
      for (int dd = d.d_min; dd <= d.d_max; ++dd) {
        float cmb = dd <= n ? 1.f : 0.f;
        for (int q = 0; q < dd && dd <= n; ++q) cmb = cmb * (float)(n - q) / (float)(q + 1);
        nkc += cmb;
      }
      const float ne = 0.5f * (float)n * (float)(n - 1);
      const float s2 = sc * sc * (fa * fa * Sff + fb * fb * Shh + 2.f * fa * fb * Sfh + 2.f * fa * fc * Sf + 2.f * fb * fc * Sh + fc * fc * ne * nkc);
      float *np = a.norm_part + ((size_t)(2 * B + b) * P->ntile_max) * 2;
      np[0] = s2 > 0.f ? s2 : 0.f;
      for (int t = 1; t < P->ntile_r2; ++t) np[2 * t] = 0.f;
    }
    // the next unit's operand stores are behind the __syncthreads above; `part` / `rs` / `dg` are rewritten after it too,
    // but only by threads that have passed this point -- order them against the readers:
    __syncthreads();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TH_MMAW) tc::tmem_dealloc(tmem, 512);
}

int tc_hnorm_launch(const DevPlan *dP, const DevPlan &hp, const TcHnormArgs &a, void *stream) {
  static CcsdSmemAttr attr;
  if (ccsd_ensure_smem(tc_hnorm_kernel, TH_SMEM, attr)) return -1;
  const int nunits = (hp.d.B + a.G - 1) / a.G;
  tc_hnorm_kernel<<<nunits < 148 ? nunits : 148, TH_THREADS, TH_SMEM, (cudaStream_t)stream>>>(dP, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#else
int tc_hnorm_launch(const DevPlan *dP, const DevPlan &hp, const TcHnormArgs &a, void *stream);
#endif

static inline int tc_hnorm_supported(const ccsd_plan_desc_t &d, int f_mode) {
  return d.is_cc && (d.nets & 4) && f_mode == 1 && d.netf.use_hodge_mask && d.E >= 8 && d.E <= 192;
}

}  // namespace ccsd
