// tc_apply.cuh -- ScoreNetworkF second contraction H . F on the tensor cores, fused with the per-entry
// network, masks, score scaling, Langevin norms / predictor update and Philox noise.
//
// Computed transposed so that the accumulator rows are CELLS:
//     D[m = cell k, n = edge e] = sum_{e'} F[e', k] * H[e, e']          (= (H F)[e, k])
//   A = F^T tile  (128 cells x 64 e' per stage): MN-major operand -- the cells are the contiguous
//       dimension of the state in HBM, so the fp32 rows are read coalesced, split to bf16 hi/lo and
//       stored in the canonical MN-major SWIZZLE_128B layout
//       (cell m, row e': (m/64)*8192 + e'*128 + (((m%64)/8) ^ (e'%8))*16 + (m%8)*2, LBO 8192, SBO 1024);
//   B = H (192 x 192, K-major SWIZZLE_128B, hi/lo), converted once per sample and resident (144 KB);
//   D in TMEM: 128 lanes x 192 fp32 columns, double buffered (384 of 512 columns), so the epilogue of
//       cell tile t overlaps the MMAs of tile t+1.
// With cells on the TMEM lanes every epilogue load/store of the state (F[e][k], noise, new state) is a
// fully coalesced 128-byte access per warp.
//
// Warp roles (13 warps): 0-3 producers, 4-11 epilogue (two warps per TMEM lane quarter, each taking
// half of the edge columns), 12 MMA issuer.  bf16x3 split as in tc_gram.cuh.
#pragma once
#include "r2_kernels.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TA_PROD = 128;
constexpr int TA_EPI = 256;
constexpr int TA_THREADS = TA_PROD + TA_EPI + 32;
constexpr int TA_NE = 192;                       // padded edge count (N of the MMA, K extent of H)
constexpr uint32_t TA_BHALF = 3u * 192u * 128u;  // 73728: hi (or lo) half of resident H (3 k-blocks)
constexpr uint32_t TA_B = 2u * TA_BHALF;         // 147456
constexpr uint32_t TA_AHALF = 2u * 8192u;        // 16384: hi (or lo) half of one A stage
constexpr uint32_t TA_ASTAGE = 2u * TA_AHALF;    // 32768
constexpr int TA_STAGES = 2;
constexpr size_t TA_SMEM = (size_t)TA_B + TA_STAGES * TA_ASTAGE + 1024 + 1024 + 6144 /*staged F-net weights*/;

static inline int tc_apply_supported(int E, int K) { return E >= 8 && E <= TA_NE && K >= 8; }

// one entry (edge e, cell k): network -> score -> mode-specific output
template <int FMODE>
__device__ __forceinline__ void r2_epilogue1(const R2Epi &c, const ApplyArgs &a, int e, int k, float f, float hf,
                                             float zraw, float fe, float fc, float &s2, float &z2) {
  const DevPlan *P = c.P;
  const float m = fe * fc;
  float o;
  if (FMODE == 1) o = m * (c.aff0 * f + c.aff1 * hf + c.aff2);
  else if (FMODE == 2) o = netf_entry_w8(P->d.netf, c.fw, c.f_nlin, f, hf, m);
  else o = netf_entry(P->d.netf, P->W, f, hf, m);
  const size_t g = ((size_t)c.b * c.E + e) * c.K + k;
  if (a.mode == MODE_EVAL) {
    a.out[g] = o;
    return;
  }
  const float s = c.co.score_scale * o;
  const float z = zraw * m;
  if (a.mode == MODE_SCORE) {
    a.out[g] = s;
    s2 += s * s;
    z2 += z * z;
  } else {
    const float mu = c.co.pa * f + c.co.pb * s;
    const float v = mu + c.co.pc * z;
    a.out[g] = v;
    if (a.write_mean) a.mean[g] = mu;
    if (a.traj && c.b == 0) a.traj[(size_t)e * c.K + k] = a.denoise ? mu : v;
  }
}

template <int FMODE>
__global__ void __launch_bounds__(TA_THREADS, 1) tc_apply_kernel(const DevPlan *__restrict__ P, ApplyArgs a) {
  extern __shared__ uint8_t ta_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, E = d.E, K = d.K, B = d.B;
  const int ntile = (K + 127) / 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t raw = tc::smem_u32(ta_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = ta_smem_raw + (base - raw);
  const uint32_t sB = base, sA = base + TA_B;
  const uint32_t bars = sA + TA_STAGES * TA_ASTAGE;
  // barriers: a_full[2] a_empty[2] b_full b_empty t_full[2] t_empty[2]  (8 bytes each), then tmem slot
  const uint32_t a_full = bars, a_empty = bars + 16, b_full = bars + 32, b_empty = bars + 40, t_full = bars + 48,
                 t_empty = bars + 64, tslot = bars + 80;
  uint8_t *gen_bars = gen + TA_B + TA_STAGES * TA_ASTAGE;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen_bars + 80);
  float *red = reinterpret_cast<float *>(gen_bars + 128);   // [2][8] warp partials
  float *fes = reinterpret_cast<float *>(gen_bars + 256);   // [192] per-edge flag products of the sample
  float *fw = reinterpret_cast<float *>(gen_bars + 1024);   // staged ScoreNetworkF weights (FMODE 2)

  if (threadIdx.x == 0) {
    for (int s = 0; s < TA_STAGES; ++s) {
      tc::mbar_init(a_full + 8 * s, TA_PROD);
      tc::mbar_init(a_empty + 8 * s, 1);
      tc::mbar_init(t_full + 8 * s, 1);
      tc::mbar_init(t_empty + 8 * s, TA_EPI);
    }
    tc::mbar_init(b_full, TA_PROD);
    tc::mbar_init(b_empty, 1);
    tc::mbar_fence_init();
  }
  if (warp == 12) tc::tmem_alloc(tslot, 512);
  if (FMODE == 2) netf_stage_w8(d.netf, P->W, fw);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const int nmine = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < 4) {
    // ===================== producers =====================
    const bool vec = (K & 3) == 0;
    uint32_t ait = 0;
    for (int sm_i = 0; sm_i < nmine; ++sm_i) {
      const int b = (int)blockIdx.x + sm_i * (int)gridDim.x;
      const float *Hb = a.H + (size_t)b * E * E;
      const float *Fb = a.r2 + (size_t)b * E * K;
      // ---- H -> resident B operand (K-major): row n, k-block kb, 16-byte chunk c ----
      tc::mbar_wait(b_empty, (sm_i & 1) ^ 1);
      for (int t = threadIdx.x; t < TA_NE * 24; t += TA_PROD) {
        const int n = t / 24, cc = t - n * 24;        // cc: chunk of 8 e' within the 192-wide row
        const int e0 = cc * 8;
        float x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (n < E && e0 + q < E) ? __ldg(Hb + (size_t)n * E + e0 + q) : 0.f;
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        const int kb = cc >> 3, c8 = cc & 7;
        const uint32_t off = (uint32_t)kb * 24576u + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                             (uint32_t)((c8 ^ (n & 7)) << 4);
        *reinterpret_cast<uint4 *>(gen + off) = hi;
        *reinterpret_cast<uint4 *>(gen + TA_BHALF + off) = lo;
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(b_full);
      // ---- F^T tiles (MN-major A operand): stage = 128 cells x 64 e' ----
      for (int ct = 0; ct < ntile; ++ct) {
        const int k0 = ct * 128;
        for (int kb = 0; kb < 3; ++kb, ++ait) {
          const int s = ait % TA_STAGES;
          const uint32_t ph = (ait / TA_STAGES) & 1;
          float x[8][8];
          // task j: e' row = (t >> 4) + 8 j, 16-byte chunk (8 cells) cq = t & 15 of the 128-cell tile
          const int cq = threadIdx.x & 15, er0 = threadIdx.x >> 4;
          const int kc = k0 + cq * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int er = kb * 64 + er0 + 8 * j;
            const float *src = Fb + (size_t)er * K + kc;
            if (er < E && vec && kc + 8 <= K) {
              const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src));
              const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
              x[j][0] = v0.x; x[j][1] = v0.y; x[j][2] = v0.z; x[j][3] = v0.w;
              x[j][4] = v1.x; x[j][5] = v1.y; x[j][6] = v1.z; x[j][7] = v1.w;
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) x[j][q] = (er < E && kc + q < K) ? __ldg(src + q) : 0.f;
            }
          }
          tc::mbar_wait(a_empty + 8 * s, ph ^ 1);
          uint8_t *st = gen + TA_B + (size_t)s * TA_ASTAGE;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int el = er0 + 8 * j;               // e' row within the stage (0..63)
            uint4 hi, lo;
            tc::split8(x[j], hi, lo);
            const uint32_t off = (uint32_t)(cq >> 3) * 8192u + (uint32_t)el * 128u + (uint32_t)(((cq & 7) ^ (el & 7)) << 4);
            *reinterpret_cast<uint4 *>(st + off) = hi;
            *reinterpret_cast<uint4 *>(st + TA_AHALF + off) = lo;
          }
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(a_full + 8 * s);
        }
      }
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::make_idesc_bf16(128, TA_NE, /*A MN-major*/ 1, /*B K-major*/ 0);
    uint32_t ait = 0, tit = 0;
    for (int sm_i = 0; sm_i < nmine; ++sm_i) {
      tc::mbar_wait(b_full, sm_i & 1);
      tc::tc_fence_after_sync();
      for (int ct = 0; ct < ntile; ++ct, ++tit) {
        const int tb = tit & 1;
        tc::mbar_wait(t_empty + 8 * tb, ((tit >> 1) & 1) ^ 1);
        tc::tc_fence_after_sync();
        const uint32_t dcol = tmem + (uint32_t)(tb * TA_NE);
        for (int kb = 0; kb < 3; ++kb, ++ait) {
          const int s = ait % TA_STAGES;
          tc::mbar_wait(a_full + 8 * s, (ait / TA_STAGES) & 1);
          tc::tc_fence_after_sync();
          if (lane == 0) {
            const uint32_t sa = sA + (uint32_t)s * TA_ASTAGE;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              // A: 16 e' rows = two 8-row groups (SBO 1024), two 64-cell blocks (LBO 8192)
              const uint64_t a_hi = tc::make_smem_desc(sa + (uint32_t)k4 * 2048u, 8192, 1024);
              const uint64_t a_lo = tc::make_smem_desc(sa + TA_AHALF + (uint32_t)k4 * 2048u, 8192, 1024);
              const uint32_t bo = (uint32_t)kb * 24576u + (uint32_t)k4 * 32u;
              const uint64_t b_hi = tc::make_smem_desc(sB + bo, 0, 1024);
              const uint64_t b_lo = tc::make_smem_desc(sB + TA_BHALF + bo, 0, 1024);
              tc::umma_bf16(dcol, a_hi, b_hi, idesc, (kb | k4) != 0);
              tc::umma_bf16(dcol, a_hi, b_lo, idesc, 1);
              tc::umma_bf16(dcol, a_lo, b_hi, idesc, 1);
            }
            tc::umma_commit(a_empty + 8 * s);
            if (kb == 2) {
              tc::umma_commit(t_full + 8 * tb);
              if (ct == ntile - 1) tc::umma_commit(b_empty);   // H may be replaced by the next sample's
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (warps 4-11) =====================
    const int ew = warp - 4;            // 0..7
    const int q = warp & 3;             // TMEM lane quarter of this warp (warp id % 4)
    const int half = ew >> 2;           // which half of the edge columns
    const int et = threadIdx.x - TA_PROD;   // 0..255
    const int Kg = P->Kp >> 2;
    uint32_t tit = 0;
    for (int sm_i = 0; sm_i < nmine; ++sm_i) {
      const int b = (int)blockIdx.x + sm_i * (int)gridDim.x;
      const float *fl = a.flags + (size_t)b * N;
      R2Epi c;
      c.P = P; c.fl = fl; c.fw = fw; c.zm = zero_mask_of(fl, N);
      c.gs = (unsigned long long)(a.nz.sample_offset + b);
      if (a.mode != MODE_EVAL) c.co = P->sched[a.nz.step * 3 + 2];
      c.b = b; c.E = E; c.K = K; c.Kg = Kg; c.f_nlin = P->f_nlin;
      c.aff0 = d.netf.aff[0]; c.aff1 = d.netf.aff[1]; c.aff2 = d.netf.aff[2];
      const float *Fb = a.r2 + (size_t)b * E * K;
      const float *Nb = a.noise ? a.noise + (size_t)b * E * K : nullptr;
      // per-edge node-mask products for this sample (shared by the 8 epilogue warps)
      asm volatile("bar.sync 1, 256;" ::: "memory");   // previous sample's readers are done
      for (int e = et; e < TA_NE; e += TA_EPI)
        fes[e] = e < E ? fl[P->edge_ij[2 * e]] * fl[P->edge_ij[2 * e + 1]] : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float s2 = 0.f, z2 = 0.f;
      for (int ct = 0; ct < ntile; ++ct, ++tit) {
        const int tb = tit & 1;
        const int k = ct * 128 + q * 32 + lane;
        const bool kval = k < K;
        const int kk = kval ? k : 0;
        const float fc = (kval && !(P->cell_mask[kk] & c.zm)) ? 1.f : 0.f;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * TA_NE);
        bool waited = false;
        for (int ch = 0; ch < 6; ++ch) {
          const int e0 = half * 96 + ch * 16;
          if (e0 < E) {
            // every global load of the chunk is issued before the accumulator is touched
            float f16[16], n16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int e = (e0 + j < E) ? e0 + j : E - 1;
              f16[j] = Fb[(size_t)e * K + kk];
              n16[j] = (Nb && a.mode != MODE_EVAL) ? Nb[(size_t)e * K + kk] : 0.f;
            }
            if (!waited) { tc::mbar_wait(t_full + 8 * tb, (tit >> 1) & 1); tc::tc_fence_after_sync(); waited = true; }
            float v[16];
            tc::tmem_ld16(trow + (uint32_t)e0, v);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              // Philox: lane i of each aligned 4-lane group draws the normals of edge e0+4*g4+i for the
              // group's 4 cells; a 4x4 exchange hands every lane its own cell's value for the 4 edges.
              float zz[4] = {n16[4 * g4], n16[4 * g4 + 1], n16[4 * g4 + 2], n16[4 * g4 + 3]};
              if (a.mode != MODE_EVAL && !Nb) {
                float z4[4];
                const int em = e0 + 4 * g4 + (lane & 3);
                normal4(a.nz.seed, c.gs, draw_id(2, a.nz.step, a.slot), (uint32_t)(em * Kg + (k >> 2)), z4);
                const int i = lane & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int srcl = (lane & ~3) | j;
                  const float t0 = __shfl_sync(0xffffffffu, z4[0], srcl), t1 = __shfl_sync(0xffffffffu, z4[1], srcl);
                  const float t2 = __shfl_sync(0xffffffffu, z4[2], srcl), t3 = __shfl_sync(0xffffffffu, z4[3], srcl);
                  zz[j] = i == 0 ? t0 : (i == 1 ? t1 : (i == 2 ? t2 : t3));
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = e0 + 4 * g4 + j;
                if (e < E && kval)
                  r2_epilogue1<FMODE>(c, a, e, k, f16[4 * g4 + j], v[4 * g4 + j], zz[j], fes[e], fc, s2, z2);
              }
            }
          }
        }
        if (!waited) { tc::mbar_wait(t_full + 8 * tb, (tit >> 1) & 1); tc::tc_fence_after_sync(); }
        tc::tc_fence_before_sync();
        tc::mbar_arrive(t_empty + 8 * tb);
      }
      if (a.mode == MODE_SCORE) {
        // per-sample squared norms: reduce over the 8 epilogue warps (named barrier 1, 256 threads)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          z2 += __shfl_xor_sync(0xffffffffu, z2, o);
        }
        if (lane == 0) { red[ew] = s2; red[8 + ew] = z2; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          float ts = 0.f, tz = 0.f;
          for (int w = 0; w < 8; ++w) { ts += red[w]; tz += red[8 + w]; }
          float *np = a.norm_part + ((size_t)(2 * d.B + b) * P->ntile_max) * 2;
          np[0] = ts; np[1] = tz;
          for (int t = 1; t < P->ntile_r2; ++t) { np[2 * t] = 0.f; np[2 * t + 1] = 0.f; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) tc::tmem_dealloc(tmem, 512);
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
