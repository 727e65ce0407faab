// tc_edge.cuh -- the per-edge MLP of AttentionLayer (attention.py:295-302) on the 5th-gen tensor cores:
//   M = MLP_nl(cat[A_1 .. A_c, adj_1 .. adj_c]) per node pair, adj_out = mask_adjs(M + M^T) = 2 M mask  (M is symmetric:
//   every input plane is).  Rows = node pairs i <= j in triangle storage -- the channel planes of the x / adj pipeline are
//   feature-major, so a channel's 128 consecutive pairs are one coalesced read.
// The widths are tiny (2 c_in <= 16 -> hidden <= 16 -> c_out <= 8), which is exactly why the fp32 register-tile path was
// slow here (its per-item overhead dominates at K = 16: 59 % of the ZINC250k step, 13-40 % of the QM9 / community_small
// steps): one tile of 128 pairs is three N = 16 MMAs per layer, the hidden activations go back to tensor memory as the
// next layer's A operand (tcgen05.mma A-from-TMEM), and a CTA needs ~50 KB of shared memory and 32 TMEM columns, so
// several of them share an SM and hide each other's latencies.  bf16x3, fp32 accumulation.
#pragma once
#include "xa_pipe.cuh"
#include "tc_common.cuh"

namespace ccsd {

constexpr int TE_WORK = 128;                 // one pair row per thread (= TMEM lane): loader and epilogue
constexpr int TE_THREADS = TE_WORK + 32;     // + the MMA-issuing warp
constexpr int TE_MMAW = TE_WORK / 32;
constexpr uint32_t TE_A1 = 0, TE_A1_HALF = 128u * 128u;             // pair features [128 rows][128 B] K-major, hi | lo
constexpr uint32_t TE_W = 2u * TE_A1_HALF;                           // 4 layers x {hi, lo} x [16 k-rows][128 B] MN-major
constexpr uint32_t TE_WL = 2u * 2048u;                               // bytes per layer (hi + lo)
constexpr uint32_t TE_VEC = TE_W + 4u * TE_WL;                       // biases [4][16]
constexpr uint32_t TE_BARS = TE_VEC + 4 * 16 * 4;
constexpr size_t TE_SMEM = (size_t)TE_BARS + 64 + 1024;

static inline int tc_edge_supported(const ccsd_plan_desc_t &d, const XpLayout &XL, const ccsd_attn_layer_t &ly) {
  const ccsd_mlp_t &m = ly.mlp;
  if (XL.big || d.N > 64 || d.N < 2) return 0;
  if (m.nl < 2 || m.nl > 4 || m.din != 2 * ly.c_in || m.din > 16 || m.dhid > 16 || m.dout > 16 || m.dout != ly.c_out) return 0;
  return 1;
}

#ifdef TC_EDGE_KERNEL_TU
__device__ __forceinline__ void te_tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __launch_bounds__(TE_THREADS, 4) tc_edge_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  extern __shared__ uint8_t te_smem_raw[];
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_attn_layer_t &ly = d.neta.layer[a.layer];
  const ccsd_mlp_t &m = ly.mlp;
  const int N = d.N, NT = L.NT, ldp = L.ldp, B = d.B, cin = ly.c_in, cout = ly.c_out, nl = m.nl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *W = P->W;

  const uint32_t raw = tc::smem_u32(te_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen = te_smem_raw + (base - raw);
  const uint32_t bar = base + TE_BARS, tslot = bar + 8;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + TE_BARS + 8);
  float *vb = reinterpret_cast<float *>(gen + TE_VEC);

  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (warp == TE_MMAW) tc::tmem_alloc(tslot, 32);
  for (uint32_t o = threadIdx.x * 16u; o < TE_VEC; o += TE_THREADS * 16u) *reinterpret_cast<uint4 *>(gen + o) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // weights: Linear l stored (in = k, out_pad) row-major -> MN-major B operand, one 64-wide n-block, 16 k rows:
  //   (n, k) at k*128 + (((n/8) ^ (k%8)) * 16) + (n%8)*2          (n < 16, k < 16)
  for (int t = threadIdx.x; t < nl * 16 * 2; t += TE_THREADS) {
    const int l = t >> 5, k = (t >> 1) & 15, n0 = (t & 1) << 3;
    const int Kin = l == 0 ? m.din : m.dhid, O = l == nl - 1 ? m.dout : m.dhid, opad = round_up(O, 8);
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = (k < Kin && n0 + q < O) ? __ldg(W + m.w[l] + (size_t)k * opad + n0 + q) : 0.f;
    uint4 hi, lo;
    tc::split8(x, hi, lo);
    const uint32_t off = TE_W + (uint32_t)l * TE_WL + (uint32_t)k * 128u + (uint32_t)(((n0 >> 3) ^ (k & 7)) << 4);
    *reinterpret_cast<uint4 *>(gen + off) = hi;
    *reinterpret_cast<uint4 *>(gen + off + 2048u) = lo;
  }
  for (int t = threadIdx.x; t < nl * 16; t += TE_THREADS) {
    const int l = t >> 4, n = t & 15, O = l == nl - 1 ? m.dout : m.dhid;
    vb[t] = n < O ? __ldg(W + m.b[l] + n) : 0.f;
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
  const uint32_t idesc = tc::make_idesc_bf16(128, 16, 0, 1);
  uint32_t phase = 0;

  // tiles: a graph's pairs in chunks of 128 (NT >= 128), or Gt = 128 / NT whole graphs per tile (small graphs)
  const int Gt = NT >= 128 ? 1 : 128 / NT;
  const int ntg = NT >= 128 ? (NT + 127) / 128 : 1;           // tiles per graph (group)
  const int ngrp = (B + Gt - 1) / Gt;
  const int ntiles = ngrp * ntg;
  const int rows_g = NT >= 128 ? 128 : NT;                    // rows of one graph inside a tile

  for (int w = blockIdx.x; w < ntiles; w += gridDim.x) {
    const int grp = w / ntg, tt = w - grp * ntg;
    const int b0 = grp * Gt, gsz = B - b0 < Gt ? B - b0 : Gt;
    const int t0 = tt * 128;
    if (warp < TE_MMAW) {
      // ---- pair features -> A1 (K-major): one row per thread, the 2 c_in channels as two 16-byte chunks ----
      const int r = threadIdx.x;
      const int gl = r / rows_g, t = t0 + (r - gl * rows_g);
      const bool live = gl < gsz && t < NT && gl < Gt;
      const size_t gb = (size_t)(b0 + (live ? gl : 0));
      const float *ga = a.g_att + gb * L.g_att + t;
      const float *gs = a.g_stack + gb * L.g_stack + (size_t)a.ch_in * ldp + t;
      float x[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float v = 0.f;
        if (live && k < 2 * cin) v = k < cin ? ga[(size_t)k * ldp] : gs[(size_t)(k - cin) * ldp];
        x[k] = v;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 hi, lo;
        tc::split8(x + half * 8, hi, lo);
        const uint32_t off = TE_A1 + (uint32_t)r * 128u + (uint32_t)((half ^ (r & 7)) << 4);
        *reinterpret_cast<uint4 *>(gen + off) = hi;
        *reinterpret_cast<uint4 *>(gen + off + TE_A1_HALF) = lo;
      }
      tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    for (int l = 0; l < nl; ++l) {
      if (warp == TE_MMAW) {
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t wb = base + TE_W + (uint32_t)l * TE_WL;
          const uint64_t b_hi = tc::make_smem_desc(wb, 2048, 1024), b_lo = tc::make_smem_desc(wb + 2048u, 2048, 1024);
          if (l == 0) {
            const uint64_t a_hi = tc::make_smem_desc(base + TE_A1, 0, 1024), a_lo = tc::make_smem_desc(base + TE_A1 + TE_A1_HALF, 0, 1024);
            tc::umma_bf16(tmem_u, a_hi, b_hi, idesc, 0);
            tc::umma_bf16(tmem_u, a_hi, b_lo, idesc, 1);
            tc::umma_bf16(tmem_u, a_lo, b_hi, idesc, 1);
          } else {
            tc::umma_bf16_ts(tmem_u, tmem_u + 16u, b_hi, idesc, 0);
            tc::umma_bf16_ts(tmem_u, tmem_u + 16u, b_lo, idesc, 1);
            tc::umma_bf16_ts(tmem_u, tmem_u + 24u, b_hi, idesc, 1);
          }
          tc::umma_commit(bar);
        }
        __syncwarp();
      }
      tc::mbar_wait(bar, phase);
      phase ^= 1u;
      tc::tc_fence_after_sync();
      if (warp < 4) {
        const int r = threadIdx.x;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        float v[16];
        tc::tmem_ld16(trow, v);
        if (l < nl - 1) {
          // hidden activations -> TMEM A operand of the next layer (element k in column k / 2; hi at +16, lo at +24)
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int c = h8 * 8 + q;
              x[q] = c < m.dhid ? fast_elu(v[c] + vb[l * 16 + c]) : 0.f;
            }
            uint4 hi, lo;
            tc::split8(x, hi, lo);
            hw[h8 * 4 + 0] = hi.x; hw[h8 * 4 + 1] = hi.y; hw[h8 * 4 + 2] = hi.z; hw[h8 * 4 + 3] = hi.w;
            lw[h8 * 4 + 0] = lo.x; lw[h8 * 4 + 1] = lo.y; lw[h8 * 4 + 2] = lo.z; lw[h8 * 4 + 3] = lo.w;
          }
          te_tmem_st8(trow + 16u, hw);
          te_tmem_st8(trow + 24u, lw);
          tc::tmem_st_wait();
        } else {
          // adj_out = 2 M mask: one plane per output channel, 128 consecutive pairs per warp store
          const int gl = r / rows_g, t = t0 + (r - gl * rows_g);
          if (gl < gsz && gl < Gt && t < ldp && (NT >= 128 || r - gl * rows_g < NT)) {
            const size_t gb = (size_t)(b0 + gl);
            float fm = 0.f;
            if (t < NT) {
              const int ij = P->tri_ij[t];
              fm = 2.0f * a.flags[gb * N + (ij >> 8)] * a.flags[gb * N + (ij & 255)];
            }
            float *dst = a.g_stack + gb * L.g_stack + (size_t)a.ch_out * ldp + t;
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < cout) dst[(size_t)c * ldp] = (v[c] + vb[l * 16 + c]) * fm;
          }
        }
      }
      tc::tc_fence_before_sync();
      __syncthreads();
    }
  }
  // pad entries [NT, ldp) of the output planes (read as don't-care rows by the next layer's 4-row groups): zero them
  if (NT < ldp)
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * cout * (ldp - NT); idx += gridDim.x * blockDim.x) {
      const int b = idx / (cout * (ldp - NT)), rem = idx - b * cout * (ldp - NT), c = rem / (ldp - NT), t = NT + rem - c * (ldp - NT);
      a.g_stack[(size_t)b * L.g_stack + (size_t)(a.ch_out + c) * ldp + t] = 0.f;
    }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TE_MMAW) tc::tmem_dealloc(tmem, 32);
}

int tc_edge_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, void *stream) {
  static CcsdSmemAttr attr;
  if (ccsd_ensure_smem(tc_edge_kernel, TE_SMEM, attr)) return -1;
  const int NT = hp.xp.NT;
  const int Gt = NT >= 128 ? 1 : 128 / NT, ntg = NT >= 128 ? (NT + 127) / 128 : 1;
  const int ntiles = ((hp.d.B + Gt - 1) / Gt) * ntg;
  const int grid = ntiles < 148 * 4 ? ntiles : 148 * 4;   // up to four CTAs per SM (50 KB of shared memory, 160 threads each)
  tc_edge_kernel<<<grid, TE_THREADS, TE_SMEM, (cudaStream_t)stream>>>(dP, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#else
int tc_edge_launch(const DevPlan *dP, const DevPlan &hp, const XaArgs &a, void *stream);
#endif

}  // namespace ccsd
