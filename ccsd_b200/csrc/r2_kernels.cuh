// r2_kernels.cuh -- the rank-2 (incidence tensor, B x E x K) side of the score evaluation, fp32 FMA
// path, plus the element-wise sampler kernels (prior, Langevin coefficients, corrector / S4 update).
//
//   gram_kernel   G = F [F ; Wp]^T  ->  H = (F F^T) (1 - I)   (cc_utils.py:917-979, hodge_laplacian +
//                 default_mask) and P0 = F Wp^T, the hodge-attention projections rank2 @ W_{q,k}
//                 (hodge_layers.py:185) as extra Gram columns.
//   proj1_kernel  P1 = rank2' Wp1^T for a second hodge layer; rank2' is the first layer's value
//                 output, an element-wise function of F because the layer-0 Hodge dual is diagonal
//                 (hodge_attention.py:107, 322-323).
//   apply_kernel  H F, the per-entry channel MLPs (hodge_layers.py:70-92), final MLP, masks
//                 (ScoreNetwork_F.py:175-217) and the sampler epilogue.
#pragma once
#include "prims.cuh"

namespace ccsd {

__device__ __forceinline__ unsigned long long zero_mask_of(const float *flags, int N) {
  unsigned long long m = 0ull;
  for (int n = 0; n < N; ++n)
    if (flags[n] == 0.f) m |= (1ull << n);
  return m;
}

// ---------------------------------------------------------------------------------------------
struct GramArgs {
  const float *r2;   // [B,E,K]
  float *H;          // [B,E,E]
  float *P0;         // [B,E,PR0]
};

CCSD_KERNEL void __launch_bounds__(256) gram_kernel(const DevPlan *__restrict__ P, GramArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const int E = d.E, K = d.K, PR0 = P->PR0, Kw = P->Kp, Ep = P->Ep;
  const int b = blockIdx.z, m0 = blockIdx.y * GRAM_BM, n0 = blockIdx.x * GRAM_BN;
  const float *Fb = a.r2 + (size_t)b * E * K;
  const float *Wp = P->W + d.neta.proj_w;  // [PR_total x Kw], rows 0..PR0-1 = hodge layer 0
  constexpr int LDT = GRAM_BM + 4;
  float *As = sm;                      // [BK][LDT]
  float *Bs = sm + GRAM_BK * LDT;      // [BK][LDT]
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#ifdef CCSD_EMU
  const int nthr = 256;  // emulate the 256 logical threads of the tile serially
  for (int k0 = 0; k0 < K; k0 += GRAM_BK) {
    for (int t = 0; t < nthr; ++t) {
      const int r = t >> 2, kq = (t & 3) << 2;
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + kq + q;
        const int m = m0 + r, n = n0 + r;
        As[(kq + q) * LDT + r] = (m < E && k < K) ? Fb[(size_t)m * K + k] : 0.f;
        float bv = 0.f;
        if (k < K) {
          if (n < E) bv = Fb[(size_t)n * K + k];
          else if (n - E < PR0) bv = Wp[(size_t)(n - E) * Kw + k];
        }
        Bs[(kq + q) * LDT + r] = bv;
      }
    }
    // serial tile product written straight to the outputs via a static accumulator tile
    static thread_local float tile[GRAM_BM][GRAM_BN];
    if (k0 == 0) memset(tile, 0, sizeof(tile));
    for (int kk = 0; kk < GRAM_BK; ++kk)
      for (int i = 0; i < GRAM_BM; ++i)
        for (int j = 0; j < GRAM_BN; ++j) tile[i][j] += As[kk * LDT + i] * Bs[kk * LDT + j];
    if (k0 + GRAM_BK >= K) {
      for (int i = 0; i < GRAM_BM; ++i)
        for (int j = 0; j < GRAM_BN; ++j) {
          const int m = m0 + i, n = n0 + j;
          if (m >= E) continue;
          if (n < E) a.H[((size_t)b * E + m) * Ep + n] = (d.netf.use_hodge_mask && m == n) ? 0.f : tile[i][j];
          else if (n - E < PR0) a.P0[((size_t)b * E + m) * PR0 + (n - E)] = tile[i][j];
        }
    }
  }
  (void)acc;
#else
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int r = t >> 2, kq = (t & 3) << 2;
  for (int k0 = 0; k0 < K; k0 += GRAM_BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + kq + q;
      const int m = m0 + r, n = n0 + r;
      As[(kq + q) * LDT + r] = (m < E && k < K) ? __ldg(Fb + (size_t)m * K + k) : 0.f;
      float bv = 0.f;
      if (k < K) {
        if (n < E) bv = __ldg(Fb + (size_t)n * K + k);
        else if (n - E < PR0) bv = __ldg(Wp + (size_t)(n - E) * Kw + k);
      }
      Bs[(kq + q) * LDT + r] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GRAM_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4 *>(As + kk * LDT + ty * 4);
      const float4 bv = *reinterpret_cast<const float4 *>(Bs + kk * LDT + tx * 4);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += aa[i] * bb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m >= E) continue;
      if (n < E) a.H[((size_t)b * E + m) * Ep + n] = (d.netf.use_hodge_mask && m == n) ? 0.f : acc[i][j];
      else if (n - E < PR0) a.P0[((size_t)b * E + m) * PR0 + (n - E)] = acc[i][j];
    }
#endif
}

// ---------------------------------------------------------------------------------------------
struct Proj1Args {
  const float *r2, *flags;        // [B,E,K] [B,N]
  const float *g_stack;           // channel stack of the x/adj pipeline: channels [0, c_init) = adjacency powers (tri storage)
  int g_stack_stride, ldp;
  float *P1;                      // [B,E,PR1]
  int epc;                        // edges per CTA
};

// One CTA per (block of `epc` edges, sample): the value output rank2' of hodge layer 0 for those edge rows
// (an element-wise function of F, see the file header) goes to shared memory, then every warp takes
// projection rows r and accumulates the epc dot products with ONE pass over the weight row, so the weights
// (PR1 x K floats, shared by all CTAs through L2) are read once per edge block instead of once per edge.
CCSD_KERNEL void __launch_bounds__(128) proj1_kernel(const DevPlan *__restrict__ P, Proj1Args a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const ccsd_neta_t &A = d.neta;
  const int N = d.N, E = d.E, K = d.K, PR0 = P->PR0h, PR1 = P->PR1, Kw = P->Kp;
  const int e0 = blockIdx.x * a.epc, b = blockIdx.y;
  const int ne = (E - e0 < a.epc) ? E - e0 : a.epc;
  const float *fl = a.flags + (size_t)b * N;
  const float *stack = a.g_stack + (size_t)b * a.g_stack_stride;
  float *r2v = sm;                 // [epc][Kw]
  const int c0 = A.c_init;
  const unsigned long long zm = zero_mask_of(fl, N);
  const ccsd_hodge_layer_t &h0 = A.hodge[0];
  // value MLP weights in shared memory as zero-padded 8-wide rows: layer l at wsm + l*72 (W[8][8], b[8])
  float *wsm = sm + (size_t)a.epc * Kw;
  const ccsd_mlp_t &mv = h0.mlp_value;
  const bool narrow = mv.din <= 8 && mv.dhid <= 8 && mv.nl <= 2;
  if (narrow)
    for (int p = threadIdx.x; p < mv.nl * 72; p += blockDim.x) {
      const int l = p / 72, q = p - l * 72;
      const int din = l == 0 ? mv.din : mv.dhid;
      wsm[p] = q < 64 ? ((q >> 3) < din ? __ldg(P->W + mv.w[l] + q) : 0.f) : __ldg(P->W + mv.b[l] + (q - 64));
    }
  __syncthreads();
  const int Kq = Kw >> 2;
  for (int p = threadIdx.x; p < ne * Kq; p += blockDim.x) {
    const int le = p / Kq, k0 = (p - le * Kq) << 2, e = e0 + le;
    const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
    const int t = i * N - (i * (i - 1)) / 2 + (j - i);   // tri index, i < j
    const float fe = fl[i] * fl[j];
    float fv[4], ov[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) fv[q] = k0 + q < K ? a.r2[((size_t)b * E + e) * K + k0 + q] : 0.f;
    if (narrow) {
      // V_c = diag(a_c) rank2 (hodge_attention.py:107) through the 1- or 2-Linear value MLP, four cells at a time
      float h[8][4];
#pragma unroll
      for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int q = 0; q < 4; ++q) h[o][q] = wsm[64 + o];
      for (int c = 0; c < c0; ++c) {
        const float ac = stack[c * a.ldp + t];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float w = wsm[c * 8 + o] * ac;
#pragma unroll
          for (int q = 0; q < 4; ++q) h[o][q] += w * fv[q];
        }
      }
      if (mv.nl == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) ov[q] = h[0][q];
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) ov[q] = wsm[72 + 64];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float w = wsm[72 + o * 8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float hv = fast_elu(h[o][q]);
            ov[q] += w * hv;
          }
        }
      }
    } else {
      for (int q = 0; q < 4; ++q) {
        float in[CCSD_MAX_CH], out[SMALL_MAX];
        for (int c = 0; c < c0; ++c) in[c] = stack[c * a.ldp + t] * fv[q];
        small_mlp(mv, P->W, in, out, ACT_ELU);
        ov[q] = out[0];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + q;
      const float fc = (k < K && !(P->cell_mask[k < K ? k : 0] & zm)) ? 1.f : 0.f;
      r2v[le * Kw + k] = ov[q] * fe * fc;   // mask_rank2 (hodge_attention.py:323)
    }
  }
  __syncthreads();
#ifdef CCSD_EMU
  const int lane = 0, nlane = 1, warp = 0, nwarp = 1;
#else
  const int lane = threadIdx.x & 31, nlane = 32, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#endif
  const float *Wp = P->W + A.proj_w + (size_t)PR0 * Kw;  // rows of hodge layer 1
  for (int r = warp; r < PR1; r += nwarp) {
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K; k += nlane) {
      const float w = __ldg(Wp + (size_t)r * Kw + k);
#pragma unroll
      for (int le = 0; le < 8; ++le)
        if (le < ne) s[le] += r2v[le * Kw + k] * w;
    }
#pragma unroll
    for (int le = 0; le < 8; ++le) {
#ifndef CCSD_EMU
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[le] += __shfl_xor_sync(0xffffffffu, s[le], o);
#endif
      if (lane == 0 && le < ne) a.P1[((size_t)b * E + e0 + le) * PR1 + r] = s[le];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ScoreNetworkF on one entry: channels [f, (H f)], HodgeNetworkLayer MLPs with mask, final MLP.
__device__ __forceinline__ float netf_entry(const ccsd_netf_t &Fn, const float *__restrict__ W, float f, float hf,
                                            float m) {
  float lst[SMALL_MAX], cur[SMALL_MAX], nxt[SMALL_MAX];
  lst[0] = f; lst[1] = hf;
  cur[0] = f; cur[1] = hf;
  int nl = 2;
  for (int l = 0; l < Fn.num_layers; ++l) {
    small_mlp(Fn.layer[l], W, cur, nxt, ACT_ELU);
    const int O = Fn.layer[l].dout;
    for (int o = 0; o < O; ++o) { cur[o] = nxt[o] * m; lst[nl + o] = cur[o]; }
    nl += O;
  }
  float out[SMALL_MAX];
  small_mlp(Fn.fin, W, lst, out, ACT_ELU);
  return out[0] * m;
}

// Fast path for networks whose every layer is at most 8 wide and whose final MLP is one Linear:
// weights staged in shared memory as zero-padded 8x8 blocks (72 floats per Linear: W[8][8], b[8]),
// then the final weights (40 floats) and bias; everything unrolled into registers.
__device__ __forceinline__ float netf_entry_w8(const ccsd_netf_t &Fn, const float *fw, int nlin, float f, float hf,
                                               float m) {
  const float *wf = fw + nlin * 72;
  float cur[8] = {f, hf, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float acc = wf[40] + wf[0] * f + wf[1] * hf;
  int off = 2, li = 0;
  for (int l = 0; l < Fn.num_layers; ++l) {
    const int nl = Fn.layer[l].nl;
    for (int i = 0; i < nl; ++i, ++li) {
      const float *Wm = fw + li * 72;
      float t[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) t[o] = Wm[64 + o];
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int o = 0; o < 8; ++o) t[o] += cur[k] * Wm[k * 8 + o];
      if (i == nl - 1) {
#pragma unroll
        for (int o = 0; o < 8; ++o) cur[o] = t[o] * m;
      } else {
#pragma unroll
        for (int o = 0; o < 8; ++o) cur[o] = t[o] > 0.f ? t[o] : expm1f(t[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) acc += wf[off + o] * cur[o];
    off += Fn.layer[l].dout;
  }
  return acc * m;
}

// Same network on FOUR entries at once when every layer is at most 4 wide (the staged 8x8 blocks are read
// as their top-left 4x4): each weight is loaded once per four entries, so the per-entry cost is ~20 shared
// loads and ~70 FMAs instead of ~290 loads.  (ENZYMES_small_CC: 2 -> 4 -> 4, 4 -> 4 -> 2, final 8 -> 1.)
__device__ __forceinline__ void netf_entry_w4x4(const ccsd_netf_t &Fn, const float *fw, int nlin, const float f[4],
                                                const float hf[4], const float m[4], float out[4]) {
  const float *wf = fw + nlin * 72;
  float cur[4][4], acc[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    cur[0][e] = f[e]; cur[1][e] = hf[e]; cur[2][e] = 0.f; cur[3][e] = 0.f;
    acc[e] = wf[40] + wf[0] * f[e] + wf[1] * hf[e];
  }
  int off = 2, li = 0;
  for (int l = 0; l < Fn.num_layers; ++l) {
    const int nl = Fn.layer[l].nl;
    for (int i = 0; i < nl; ++i, ++li) {
      const float *Wm = fw + li * 72;
      float t[4][4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const float bv = Wm[64 + o];
#pragma unroll
        for (int e = 0; e < 4; ++e) t[o][e] = bv;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 wr = *reinterpret_cast<const float4 *>(Wm + k * 8);
        const float wv[4] = {wr.x, wr.y, wr.z, wr.w};
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int e = 0; e < 4; ++e) t[o][e] += cur[k][e] * wv[o];
      }
      if (i == nl - 1) {
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int e = 0; e < 4; ++e) cur[o][e] = t[o][e] * m[e];
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int e = 0; e < 4; ++e) cur[o][e] = fast_elu(t[o][e]);
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float wv = wf[off + o];   // columns past dout hold zeros in cur, so a neighbouring weight is harmless
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += wv * cur[o][e];
    }
    off += Fn.layer[l].dout;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = acc[e] * m[e];
}

__device__ __forceinline__ void netf_stage_w8(const ccsd_netf_t &Fn, const float *__restrict__ W, float *fw) {
  int li = 0;
  for (int l = 0; l < Fn.num_layers; ++l) {
    const ccsd_mlp_t &M = Fn.layer[l];
    for (int i = 0; i < M.nl; ++i, ++li) {
      const int din = (i == 0) ? M.din : M.dhid;
      for (int p = threadIdx.x; p < 72; p += blockDim.x) {
        float v;
        if (p < 64) v = (p / 8 < din) ? __ldg(W + M.w[i] + p) : 0.f;   // rows are 8 floats (out_pad = 8)
        else v = __ldg(W + M.b[i] + (p - 64));
        fw[li * 72 + p] = v;
      }
    }
  }
  // final Linear: (fd, out_pad=8) with out = 1 -> column 0
  for (int p = threadIdx.x; p < 41; p += blockDim.x) {
    float v = 0.f;
    if (p < 40) v = (p < Fn.fin.din) ? __ldg(W + Fn.fin.w[0] + p * 8) : 0.f;
    else v = __ldg(W + Fn.fin.b[0]);
    fw[li * 72 + p] = v;
  }
}

// f_mode 4: layers at most 8 wide (staged 8x8 blocks as above, at most 3 of them) and a TWO-Linear final MLP
// (num_layers_mlp == 2: the Base_CC checkpoints).  The final MLP's first Linear is staged per hidden unit h
// as a row of NETF_F2_LD floats over the statically indexed register inputs:
//   [0] f, [1] H f, [2 + 8 l + o] output o of layer l (zero weight past dout), [26] b1[h], [27] w2[h]
// followed by b2; every thread of a warp reads the same row (shared-memory broadcast).
constexpr int NETF_F2_LD = 28;
constexpr int NETF_F2_MAXL = 3;
__device__ __forceinline__ float netf_entry_w8f2(const ccsd_netf_t &Fn, const float *fw, int nlin, float f, float hf,
                                                 float m) {
  float lay[NETF_F2_MAXL][8];
  float cur[8] = {f, hf, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int li = 0;
#pragma unroll
  for (int l = 0; l < NETF_F2_MAXL; ++l) {
    if (l < Fn.num_layers) {
      const int nl = Fn.layer[l].nl;
      for (int i = 0; i < nl; ++i, ++li) {
        const float *Wm = fw + li * 72;
        float t[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) t[o] = Wm[64 + o];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int o = 0; o < 8; ++o) t[o] += cur[k] * Wm[k * 8 + o];
        if (i == nl - 1) {
#pragma unroll
          for (int o = 0; o < 8; ++o) cur[o] = t[o] * m;
        } else {
#pragma unroll
          for (int o = 0; o < 8; ++o) cur[o] = fast_elu(t[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) lay[l][o] = (l < Fn.num_layers) ? cur[o] : 0.f;
  }
  const float *wf = fw + nlin * 72;
  const int dh = Fn.fin.dhid;
  float acc = wf[dh * NETF_F2_LD];
  for (int h = 0; h < dh; ++h) {
    const float *r = wf + h * NETF_F2_LD;
    float t = r[26] + r[0] * f + r[1] * hf;
#pragma unroll
    for (int l = 0; l < NETF_F2_MAXL; ++l)
#pragma unroll
      for (int o = 0; o < 8; ++o) t += r[2 + 8 * l + o] * lay[l][o];
    acc += r[27] * fast_elu(t);
  }
  return acc * m;
}

// staged size in floats of the f_mode 2/3 (final = 1 Linear) and f_mode 4 layouts
static inline int netf_stage_floats(const ccsd_netf_t &Fn, int nlin, int f_mode) {
  return nlin * 72 + (f_mode == 4 ? Fn.fin.dhid * NETF_F2_LD + 4 : 44);
}

__device__ __forceinline__ void netf_stage_w8f2(const ccsd_netf_t &Fn, const float *__restrict__ W, float *fw) {
  int li = 0;
  for (int l = 0; l < Fn.num_layers; ++l) {
    const ccsd_mlp_t &M = Fn.layer[l];
    for (int i = 0; i < M.nl; ++i, ++li) {
      const int din = (i == 0) ? M.din : M.dhid;
      for (int p = threadIdx.x; p < 72; p += blockDim.x) {
        float v;
        if (p < 64) v = (p / 8 < din) ? __ldg(W + M.w[i] + p) : 0.f;
        else v = __ldg(W + M.b[i] + (p - 64));
        fw[li * 72 + p] = v;
      }
    }
  }
  float *wf = fw + li * 72;
  const ccsd_mlp_t &Fm = Fn.fin;
  const int dh = Fm.dhid, hp = (dh + 7) / 8 * 8;    // first Linear stored (fdim, hp); second (dh, 8) with out = 1
  for (int p = threadIdx.x; p < dh * NETF_F2_LD; p += blockDim.x) {
    const int h = p / NETF_F2_LD, c = p - h * NETF_F2_LD;
    float v = 0.f;
    if (c == 26) v = __ldg(W + Fm.b[0] + h);
    else if (c == 27) v = __ldg(W + Fm.w[1] + h * 8);
    else if (c < 2) v = __ldg(W + Fm.w[0] + c * hp + h);
    else {
      // input row of lst = [f, hf, layer 0 outputs, layer 1 outputs, ...]
      const int l = (c - 2) >> 3, o = (c - 2) & 7;
      if (l < Fn.num_layers && o < Fn.layer[l].dout) {
        int row = 2;
        for (int q = 0; q < l; ++q) row += Fn.layer[q].dout;
        v = __ldg(W + Fm.w[0] + (size_t)(row + o) * hp + h);
      }
    }
    wf[p] = v;
  }
  if (threadIdx.x == 0) wf[dh * NETF_F2_LD] = __ldg(W + Fm.b[1]);
}

struct ApplyArgs {
  const float *r2, *H, *flags;   // [B,E,K] [B,E,E] [B,N]
  int mode;
  float *out;                    // EVAL: raw output; SCORE: scaled score; PRED: new state (may alias r2)
  float *mean;                   // PRED: means (written when write_mean)
  float *norm_part;
  const float *coef;             // CORR: [3][2] Langevin step sizes from coef_kernel
  const float *noise;            // raw normals [B,E,K] for this draw or nullptr
  float *traj;                   // PRED: sample 0 destination or nullptr
  int slot, denoise, write_mean;
  NoiseCtx nz;
  long long *trace;              // debug: per-tile clock64 stamps of CTA 0 ([tile][16]) or nullptr
  const unsigned long long *zmask;   // [B] zero-flag node masks of the samples (zmask_kernel)
};

// Per-CTA epilogue context (everything the per-entry work needs), shared by the fp32 and the
// tensor-core apply kernels.
struct R2Epi {
  const DevPlan *P;
  const float *fl;       // flags of this sample
  const float *fw;       // staged ScoreNetworkF weights (f_mode 2) in shared memory
  unsigned long long zm; // zero-flag node mask
  unsigned long long gs; // global sample index
  ccsd_objcoef_t co;
  int b, E, K, Kg, f_nlin;
  int step;              // diffusion step (nz_step: argument or device-resident), read once per CTA
  float aff0, aff1, aff2;
};

// 4 consecutive cells k0..k0+3 of edge row e: network -> score -> mode-specific output.
// f4: state entries, hf4: (H F) entries.  Accumulates the squared norms for MODE_SCORE.
template <int FMODE>
__device__ __forceinline__ void r2_epilogue4(const R2Epi &c, const ApplyArgs &a, int e, int k0, const float f4[4],
                                             const float hf4[4], float &s2, float &z2) {
  const DevPlan *P = c.P;
  const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
  const float fe = c.fl[i] * c.fl[j];
  float z4[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t g0 = ((size_t)c.b * c.E + e) * c.K + k0;
  if (a.mode != MODE_EVAL) {
    if (a.noise) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (k0 + q < c.K) z4[q] = a.noise[g0 + q];
    } else {
      normal4(a.nz.seed, c.gs, draw_id(2, c.step, a.slot), (uint32_t)(e * c.Kg + (k0 >> 2)), z4);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int k = k0 + q;
    if (k < c.K) {
      const float fc = (P->cell_mask[k] & c.zm) ? 0.f : 1.f;
      const float m = fe * fc;
      const float f = f4[q];
      float o;
      if (FMODE == 1) o = m * (c.aff0 * f + c.aff1 * hf4[q] + c.aff2);
      else if (FMODE == 2) o = netf_entry_w8(P->d.netf, c.fw, c.f_nlin, f, hf4[q], m);
      else if (FMODE == 4) o = netf_entry_w8f2(P->d.netf, c.fw, c.f_nlin, f, hf4[q], m);
      else o = netf_entry(P->d.netf, P->W, f, hf4[q], m);
      const size_t g = g0 + q;
      if (a.mode == MODE_EVAL) {
        a.out[g] = o;
      } else {
        const float s = c.co.score_scale * o;
        const float z = z4[q] * m;
        if (a.mode == MODE_SCORE || a.mode == MODE_NORM) {
          if (a.mode == MODE_SCORE) a.out[g] = s;
          s2 += s * s;
          z2 += z * z;
        } else if (a.mode == MODE_CORR) {
          a.out[g] = f + a.coef[4] * s + a.coef[5] * z;   // Langevin (solver.py:784-785)
        } else {
          const float mu = c.co.pa * f + c.co.pb * s;
          const float v = mu + c.co.pc * z;
          a.out[g] = v;
          if (a.write_mean) a.mean[g] = mu;
          if (a.traj && c.b == 0) a.traj[(size_t)e * c.K + k] = a.denoise ? mu : v;
        }
      }
    }
  }
}

template <int FMODE>
__global__ void __launch_bounds__(256) apply_kernel(const DevPlan *__restrict__ P, ApplyArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, E = d.E, K = d.K;
  const int b = blockIdx.y, n0 = blockIdx.x * APPLY_TN;
  const float *Fb = a.r2 + (size_t)b * E * K;
  const float *Hb = a.H + (size_t)b * E * P->Ep;
  const float *fl = a.flags + (size_t)b * N;
  constexpr int LDF = APPLY_TN + 4, LDH = 64 + 4;
  float *Fs = sm;                         // [E][LDF]
  float *Hs = sm + (size_t)E * LDF;       // [16][LDH]
  float *red = Hs + 16 * LDH;             // [40]
  float *fw = red + 40;                   // staged ScoreNetworkF weights (f_mode 2)
  if (FMODE == 2) netf_stage_w8(d.netf, P->W, fw);
  if (FMODE == 4) netf_stage_w8f2(d.netf, P->W, fw);
  R2Epi c;
  c.P = P; c.fl = fl; c.fw = fw; c.zm = zero_mask_of(fl, N);
  c.gs = (unsigned long long)(a.nz.sample_offset + b);
  c.step = a.mode != MODE_EVAL ? nz_step(a.nz) : 0;
  if (a.mode != MODE_EVAL) c.co = P->sched[c.step * 3 + 2];
  c.b = b; c.E = E; c.K = K; c.Kg = P->Kp >> 2; c.f_nlin = P->f_nlin;
  c.aff0 = d.netf.aff[0]; c.aff1 = d.netf.aff[1]; c.aff2 = d.netf.aff[2];

  for (int p = threadIdx.x; p < E * APPLY_TN; p += blockDim.x) {
    const int e = p / APPLY_TN, cc = p - e * APPLY_TN;
    const int k = n0 + cc;
    Fs[e * LDF + cc] = (k < K) ? Fb[(size_t)e * K + k] : 0.f;
  }
  __syncthreads();

  float s2 = 0.f, z2 = 0.f;
#ifdef CCSD_EMU
  (void)Hs;
  for (int e = 0; e < E; ++e)
    for (int cq = 0; cq < APPLY_TN; cq += 4) {
      if (n0 + cq >= K) continue;
      float hf4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int e2 = 0; e2 < E; ++e2) {
        const float h = Hb[(size_t)e * P->Ep + e2];
        for (int q = 0; q < 4; ++q) hf4[q] += h * Fs[e2 * LDF + cq + q];
      }
      r2_epilogue4<FMODE>(c, a, e, n0 + cq, Fs + e * LDF + cq, hf4, s2, z2);
    }
#else
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lr = t >> 2, lq = (t & 3) << 2;
  for (int m0 = 0; m0 < E; m0 += 64) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int e0 = 0; e0 < E; e0 += 16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + lr, e2 = e0 + lq + q;
        Hs[(lq + q) * LDH + lr] = (m < E && e2 < E) ? __ldg(Hb + (size_t)m * P->Ep + e2) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const int e2 = e0 + kk;
        if (e2 < E) {
          const float4 av = *reinterpret_cast<const float4 *>(Hs + kk * LDH + ty * 4);
          const float4 bv = *reinterpret_cast<const float4 *>(Fs + e2 * LDF + tx * 4);
          acc[0][0] += av.x * bv.x; acc[0][1] += av.x * bv.y; acc[0][2] += av.x * bv.z; acc[0][3] += av.x * bv.w;
          acc[1][0] += av.y * bv.x; acc[1][1] += av.y * bv.y; acc[1][2] += av.y * bv.z; acc[1][3] += av.y * bv.w;
          acc[2][0] += av.z * bv.x; acc[2][1] += av.z * bv.y; acc[2][2] += av.z * bv.z; acc[2][3] += av.z * bv.w;
          acc[3][0] += av.w * bv.x; acc[3][1] += av.w * bv.y; acc[3][2] += av.w * bv.z; acc[3][3] += av.w * bv.w;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = m0 + ty * 4 + i;
      if (e < E && n0 + tx * 4 < K) {
        const float4 fv = *reinterpret_cast<const float4 *>(Fs + e * LDF + tx * 4);
        const float f4[4] = {fv.x, fv.y, fv.z, fv.w};
        r2_epilogue4<FMODE>(c, a, e, n0 + tx * 4, f4, acc[i], s2, z2);
      }
    }
  }
#endif
  if (a.mode == MODE_SCORE || a.mode == MODE_NORM) {
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(2 * d.B + b) * P->ntile_max + blockIdx.x) * 2;
      np[0] = s2; np[1] = z2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Large-E fallback of the apply pass (E > 192 and an F column block [E][64] that no longer fits shared memory:
// grid_small_CC, E = 1176, K = 18424): HF = H F as a plain tiled fp32 GEMM into a scratch tensor, then an
// element-wise kernel that runs the same per-entry epilogue (r2_epilogue4) as apply_kernel.
struct HfArgs {
  const float *r2, *H;   // [B,E,K] [B,E,Ep]
  float *hf;             // [B,E,K]
};

CCSD_KERNEL void __launch_bounds__(256) hf_gemm_kernel(const DevPlan *__restrict__ P, HfArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const int E = d.E, K = d.K, Ep = P->Ep;
  const int b = blockIdx.z, m0 = blockIdx.y * GRAM_BM, n0 = blockIdx.x * GRAM_BN;
  const float *Fb = a.r2 + (size_t)b * E * K;
  const float *Hb = a.H + (size_t)b * E * Ep;
  float *Ob = a.hf + (size_t)b * E * K;
  constexpr int LDT = GRAM_BM + 4;
  float *As = sm;                      // [BK][LDT]   As[kk][m] = H[m0 + m][e0 + kk]
  float *Bs = sm + GRAM_BK * LDT;      // [BK][LDT]   Bs[kk][n] = F[e0 + kk][n0 + n]
#ifdef CCSD_EMU
  (void)As; (void)Bs;
  for (int i = 0; i < GRAM_BM; ++i)
    for (int j = 0; j < GRAM_BN; ++j) {
      const int m = m0 + i, n = n0 + j;
      if (m >= E || n >= K) continue;
      float s = 0.f;
      for (int e2 = 0; e2 < E; ++e2) s += Hb[(size_t)m * Ep + e2] * Fb[(size_t)e2 * K + n];
      Ob[(size_t)m * K + n] = s;
    }
#else
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ar = t >> 2, aq = (t & 3) << 2;      // A loader: row ar, 4 consecutive k
  const int bk = t >> 4, bn = (t & 15) << 2;     // B loader: k row bk, 4 consecutive n
  for (int e0 = 0; e0 < E; e0 += GRAM_BK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = m0 + ar, e2 = e0 + aq + q;
      As[(aq + q) * LDT + ar] = (m < E && e2 < E) ? __ldg(Hb + (size_t)m * Ep + e2) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e2 = e0 + bk, n = n0 + bn + q;
      Bs[bk * LDT + bn + q] = (e2 < E && n < K) ? __ldg(Fb + (size_t)e2 * K + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GRAM_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4 *>(As + kk * LDT + ty * 4);
      const float4 bv = *reinterpret_cast<const float4 *>(Bs + kk * LDT + tx * 4);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += aa[i] * bb[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < E && n < K) Ob[(size_t)m * K + n] = acc[i][j];
    }
#endif
}

constexpr int R2EPI_CHUNKS = 64;   // CTAs (= norm partial slots) per sample of r2_epi_kernel

// a.out may alias hf (MODE_SCORE) or a.r2 (CORR / PRED): every entry is read before it is written by its own thread
template <int FMODE>
__global__ void __launch_bounds__(256) r2_epi_kernel(const DevPlan *__restrict__ P, ApplyArgs a, const float *hf) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, E = d.E, K = d.K, Kg = P->Kp >> 2;
  const int b = blockIdx.y;
  const float *Fb = a.r2 + (size_t)b * E * K;
  const float *HFb = hf + (size_t)b * E * K;
  const float *fl = a.flags + (size_t)b * N;
  float *red = sm, *fw = sm + 40;
  if (FMODE == 2) netf_stage_w8(d.netf, P->W, fw);
  if (FMODE == 4) netf_stage_w8f2(d.netf, P->W, fw);
  R2Epi c;
  c.P = P; c.fl = fl; c.fw = fw; c.zm = zero_mask_of(fl, N);
  c.gs = (unsigned long long)(a.nz.sample_offset + b);
  c.step = a.mode != MODE_EVAL ? nz_step(a.nz) : 0;
  if (a.mode != MODE_EVAL) c.co = P->sched[c.step * 3 + 2];
  c.b = b; c.E = E; c.K = K; c.Kg = Kg; c.f_nlin = P->f_nlin;
  c.aff0 = d.netf.aff[0]; c.aff1 = d.netf.aff[1]; c.aff2 = d.netf.aff[2];
  __syncthreads();
  float s2 = 0.f, z2 = 0.f;
  const long long groups = (long long)E * Kg;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(g / Kg), k0 = (int)(g - (long long)e * Kg) << 2;
    float f4[4], h4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool in = k0 + q < K;
      f4[q] = in ? Fb[(size_t)e * K + k0 + q] : 0.f;
      h4[q] = in ? HFb[(size_t)e * K + k0 + q] : 0.f;
    }
    r2_epilogue4<FMODE>(c, a, e, k0, f4, h4, s2, z2);
  }
  if (a.mode == MODE_SCORE || a.mode == MODE_NORM) {
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(2 * d.B + b) * P->ntile_max + blockIdx.x) * 2;
      np[0] = s2; np[1] = z2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// prior: state = mask(raw normal)  (solver.py:963-968, 1111-1118; sde.py:436, 448-449)
struct InitArgs {
  const float *flags;
  const float *px, *padj, *pr2;  // raw normals or nullptr
  float *x, *adj, *r2;
  NoiseCtx nz;
};

CCSD_KERNEL void __launch_bounds__(256) init_kernel(const DevPlan *__restrict__ P, InitArgs a) {
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, F = d.F, E = d.E, K = d.K, Kg = P->Kp >> 2;
  const int obj = blockIdx.y;
  const size_t stride = (size_t)gridDim.x * blockDim.x, start = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (obj == 0) {
    for (size_t g = start; g < (size_t)d.B * N * F; g += stride) {
      const int b = (int)(g / (N * F)), p = (int)(g - (size_t)b * N * F), i = p / F;
      const float z = a.px ? a.px[g] : normal1(a.nz.seed, a.nz.sample_offset + b, draw_id(0, -1, 0), p);
      a.x[g] = z * a.flags[(size_t)b * N + i];
    }
  } else if (obj == 1) {
    for (size_t g = start; g < (size_t)d.B * N * N; g += stride) {
      const int b = (int)(g / (N * N)), p = (int)(g - (size_t)b * N * N), i = p / N, j = p - i * N;
      float z = 0.f;
      if (i != j) {
        const int q = (i < j) ? i * N + j : j * N + i;
        z = a.padj ? a.padj[(size_t)b * N * N + q] : normal1(a.nz.seed, a.nz.sample_offset + b, draw_id(1, -1, 0), q);
      }
      a.adj[g] = z * a.flags[(size_t)b * N + i] * a.flags[(size_t)b * N + j];
    }
  } else if (d.is_cc) {
    for (size_t u = start; u < (size_t)d.B * E * Kg; u += stride) {
      const int b = (int)(u / ((size_t)E * Kg));
      const int rem = (int)(u - (size_t)b * E * Kg), e = rem / Kg, kg = rem - e * Kg;
      const float *fl = a.flags + (size_t)b * N;
      const unsigned long long zm = zero_mask_of(fl, N);
      const float fe = fl[P->edge_ij[2 * e]] * fl[P->edge_ij[2 * e + 1]];
      float z4[4] = {0.f, 0.f, 0.f, 0.f};
      if (!a.pr2) normal4(a.nz.seed, a.nz.sample_offset + b, draw_id(2, -1, 0), (uint32_t)(e * Kg + kg), z4);
      for (int q = 0; q < 4; ++q) {
        const int k = kg * 4 + q;
        if (k >= K) break;
        const size_t g = ((size_t)b * E + e) * K + k;
        const float z = a.pr2 ? a.pr2[g] : z4[q];
        a.r2[g] = z * fe * ((P->cell_mask[k] & zm) ? 0.f : 1.f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Langevin / S4 step sizes from the batch means of the per-sample norms
// (solver.py:695-701, 763-769, 1300-1316):  cs = (snr*mean||z||/mean||g||)^2 * 2*alpha, cn = sqrt(2 cs)*eps
struct CoefArgs {
  const float *norm_part;  // [3][B][ntile_max][2]
  float *coef;             // [3][2]
  int step, s4;
  int obj_mask;            // bit k: compute object k's step sizes (the others keep their previous values)
  const StepDev *sd;       // graph replay: the step index lives on the device (plan_dev.h)
};

// graph replay (ccsd_plan_run): set / advance the device-resident step state
CCSD_KERNEL void step_set_kernel(StepDev *sd, int step, float *tx, float *ta) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { sd->step = step; sd->pad = 0; sd->tx = tx; sd->ta = ta; }
}
CCSD_KERNEL void step_advance_kernel(StepDev *sd, int stride_x, int stride_adj) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    sd->step += 1;
    if (sd->tx) sd->tx += stride_x;
    if (sd->ta) sd->ta += stride_adj;
  }
}

// One CTA per object (x, adj, rank2), up to 1024 threads: the batch mean is a fixed-order reduction (thread-strided partial
// sums, then block_sum), so a run is bit reproducible; as one 256-thread CTA for all objects it was 0.11 ms per step at B = 10000.
CCSD_KERNEL void __launch_bounds__(1024) coef_kernel(const DevPlan *__restrict__ P, CoefArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  {
    const int obj = blockIdx.x;
    if (!((a.obj_mask >> obj) & 1)) return;   // uniform per CTA
    const int nt = (obj == 2) ? P->ntile_r2 : (obj == 1 ? P->ntile_adj : P->ntile_x);
    float gs = 0.f, zs = 0.f;
    for (int b = threadIdx.x; b < d.B; b += blockDim.x) {
      const float *np = a.norm_part + ((size_t)(obj * d.B + b) * P->ntile_max) * 2;
      float g2 = 0.f, z2 = 0.f;
      for (int t = 0; t < nt; ++t) { g2 += np[2 * t]; z2 += np[2 * t + 1]; }
      gs += sqrtf(g2);
      zs += sqrtf(z2);
    }
    gs = block_sum(gs, sm);
    zs = block_sum(zs, sm);
    if (threadIdx.x == 0) {
      const ccsd_objcoef_t co = P->sched[(a.sd ? a.sd->step : a.step) * 3 + obj];
      const float alpha = a.s4 ? co.s4_alpha : co.lg_alpha;
      const float gm = gs / (float)d.B, zm = zs / (float)d.B;
      const float r = d.snr * zm / gm;
      const float cs = r * r * 2.f * alpha;
      a.coef[obj * 2 + 0] = cs;
      a.coef[obj * 2 + 1] = sqrtf(cs * 2.f) * d.scale_eps;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Corrector update (PC) or the whole S4 chain after the scores are known.
struct UpdateArgs {
  const float *flags;
  float *x, *adj, *r2;                 // state, updated in place
  const float *sx, *sadj, *sr2;        // scaled scores
  const float *coef;                   // [3][2] from coef_kernel
  float *mx, *madj, *mr2;              // means (S4)
  const float *nx, *nadj, *nr2;        // injected raw normals [n_draws][B][...] or nullptr
  float *tx, *tadj, *tr2;              // traj destinations (S4) or nullptr
  int s4, denoise, write_mean_r2;
  int obj0;                            // first object of this launch (object = obj0 + blockIdx.y)
  int slot0;                           // draw slot of the (first) noise draw: Langevin inner step (PC), 0 (S4)
  NoiseCtx nz;
};

__device__ __forceinline__ float upd_elem(const ccsd_objcoef_t &co, float cs, float cn, int s4, float v, float s,
                                          const float z[3], float *mean) {
  float o = v + cs * s + cn * z[0];  // Langevin (solver.py:700-701) / S4 correction (:1312-1316)
  if (!s4) { *mean = v + cs * s; return o; }
  o = co.s4_m1 * o + co.s4_s1 * z[1];       // transition(t, dt/2)      (solver.py:1339-1342)
  o = o + co.s4_sd * s;                     // + Sdrift * dt            (:1344-1345)
  const float mu = co.s4_m2 * o;            // transition(t+dt/2, dt/2) (:1347-1350)
  *mean = mu;
  return mu + co.s4_s2 * z[2];
}

CCSD_KERNEL void __launch_bounds__(256) update_kernel(const DevPlan *__restrict__ P, UpdateArgs a) {
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, F = d.F, E = d.E, K = d.K, Kg = P->Kp >> 2;
  const int obj = a.obj0 + blockIdx.y;
  const int ndraw = a.s4 ? 3 : 1;
  const int stp = nz_step(a.nz);
  const ccsd_objcoef_t co = P->sched[stp * 3 + obj];
  const float cs = a.coef[obj * 2], cn = a.coef[obj * 2 + 1];
  const size_t stride = (size_t)gridDim.x * blockDim.x, start = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (obj == 0) {
    const size_t tot = (size_t)d.B * N * F;
    for (size_t g = start; g < tot; g += stride) {
      const int b = (int)(g / (N * F)), p = (int)(g - (size_t)b * N * F), i = p / F;
      const float f = a.flags[(size_t)b * N + i];
      float z[3] = {0.f, 0.f, 0.f};
      for (int s = 0; s < ndraw; ++s)
        z[s] = f * (a.nx ? a.nx[(size_t)(a.slot0 + s) * tot + g]
                         : normal1(a.nz.seed, a.nz.sample_offset + b, draw_id(0, stp, a.slot0 + s), p));
      float mean;
      const float v = upd_elem(co, cs, cn, a.s4, a.x[g], a.sx[g], z, &mean);
      a.x[g] = v;
      if (a.s4) {
        a.mx[g] = mean;
        if (b == 0) { float *tjx = a.nz.sd ? a.nz.sd->tx : a.tx; if (tjx) tjx[p] = a.denoise ? mean : v; }
      }
    }
  } else if (obj == 1) {
    const size_t tot = (size_t)d.B * N * N;
    for (size_t g = start; g < tot; g += stride) {
      const int b = (int)(g / (N * N)), p = (int)(g - (size_t)b * N * N), i = p / N, j = p - i * N;
      const float f = a.flags[(size_t)b * N + i] * a.flags[(size_t)b * N + j];
      float z[3] = {0.f, 0.f, 0.f};
      if (i != j) {
        const int q = (i < j) ? i * N + j : j * N + i;
        for (int s = 0; s < ndraw; ++s)
          z[s] = f * (a.nadj ? a.nadj[(size_t)(a.slot0 + s) * tot + (size_t)b * N * N + q]
                             : normal1(a.nz.seed, a.nz.sample_offset + b, draw_id(1, stp, a.slot0 + s), q));
      }
      float mean;
      const float v = upd_elem(co, cs, cn, a.s4, a.adj[g], a.sadj[g], z, &mean);
      a.adj[g] = v;
      if (a.s4) {
        a.madj[g] = mean;
        if (b == 0) { float *tja = a.nz.sd ? a.nz.sd->ta : a.tadj; if (tja) tja[p] = a.denoise ? mean : v; }
      }
    }
  } else if (d.is_cc) {
    const size_t tot = (size_t)d.B * E * K;
    for (size_t u = start; u < (size_t)d.B * E * Kg; u += stride) {
      const int b = (int)(u / ((size_t)E * Kg));
      const int rem = (int)(u - (size_t)b * E * Kg), e = rem / Kg, kg = rem - e * Kg;
      const float *fl = a.flags + (size_t)b * N;
      const unsigned long long zm = zero_mask_of(fl, N);
      const float fe = fl[P->edge_ij[2 * e]] * fl[P->edge_ij[2 * e + 1]];
      float zz[3][4];
      for (int s = 0; s < 3; ++s)
        for (int q = 0; q < 4; ++q) zz[s][q] = 0.f;
      if (!a.nr2)
        for (int s = 0; s < ndraw; ++s)
          normal4(a.nz.seed, a.nz.sample_offset + b, draw_id(2, stp, a.slot0 + s), (uint32_t)(e * Kg + kg), zz[s]);
      for (int q = 0; q < 4; ++q) {
        const int k = kg * 4 + q;
        if (k >= K) break;
        const size_t g = ((size_t)b * E + e) * K + k;
        const float m = fe * ((P->cell_mask[k] & zm) ? 0.f : 1.f);
        float z[3] = {0.f, 0.f, 0.f};
        for (int s = 0; s < ndraw; ++s) z[s] = m * (a.nr2 ? a.nr2[(size_t)(a.slot0 + s) * tot + g] : zz[s][q]);
        float mean;
        const float v = upd_elem(co, cs, cn, a.s4, a.r2[g], a.sr2[g], z, &mean);
        a.r2[g] = v;
        if (a.s4) {
          if (a.write_mean_r2) a.mr2[g] = mean;
          if (a.tr2 && b == 0) a.tr2[(size_t)e * K + k] = a.denoise ? mean : v;
        }
      }
    }
  }
}

// zero-flag node mask of every sample (bit n set <=> flags[b][n] == 0): a cell is masked iff it shares a bit
CCSD_KERNEL void __launch_bounds__(256) zmask_kernel(const float *__restrict__ flags, unsigned long long *__restrict__ zmask,
                                                    int B, int N) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x)
    zmask[b] = zero_mask_of(flags + (size_t)b * N, N);
}

// quantize / quantize_mol (graph_utils.py:181-213)
CCSD_KERNEL void __launch_bounds__(256) quantize_kernel(const float *__restrict__ in, uint8_t *__restrict__ out,
                                                        size_t n, float thr, int mol) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
    const float v = in[g];
    uint8_t q;
    if (mol) q = v >= 2.5f ? 3 : (v >= 1.5f ? 2 : (v >= 0.5f ? 1 : 0));
    else q = v < thr ? 0 : 1;
    out[g] = q;
  }
}

// Molecule post-processing of Sampler_mol.sample (sampler.py:814-825) in one pass over the sampler outputs:
//   adj_out[b][c][i][j] = one_hot((quantize_mol(adj) - 1) mod 4)   (classes: 0 single, 1 double, 2 triple, 3 no bond), int64
//   x_out[b][i][f] = x > 0.5 (f < F),  x_out[b][i][F] = 1 - sum_f x_out[b][i][f]   (the "no atom" column), int64
CCSD_KERNEL void __launch_bounds__(256) mol_onehot_kernel(const float *__restrict__ x, const float *__restrict__ adj,
                                                          long long *__restrict__ x_out, long long *__restrict__ adj_out,
                                                          int B, int N, int F) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, start = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t na = (size_t)B * N * N, nn = (size_t)N * N;
  for (size_t g = start; g < na; g += stride) {
    const size_t b = g / nn, p = g - b * nn;
    const float v = adj[g];
    const int q = v >= 2.5f ? 3 : (v >= 1.5f ? 2 : (v >= 0.5f ? 1 : 0));
    const int cls = q == 0 ? 3 : q - 1;
#pragma unroll
    for (int c = 0; c < 4; ++c) adj_out[(b * 4 + c) * nn + p] = (c == cls) ? 1 : 0;
  }
  const size_t nx = (size_t)B * N;
  for (size_t g = start; g < nx; g += stride) {
    long long s = 0;
    for (int f = 0; f < F; ++f) {
      const long long o = x[g * F + f] > 0.5f ? 1 : 0;
      x_out[g * (F + 1) + f] = o;
      s += o;
    }
    x_out[g * (F + 1) + F] = 1 - s;
  }
}

// ---------------------------------------------------------------------------------------------
// Batched core of cc_from_incidence (cc_utils.py:156-265): for every sample and every candidate rank-2 cell k
//   present[b, k] = any_e (rank2[b, e, k] != 0)                            (:244 incidence_matrix[:, i].any())
//   row[b, k]     = argmax_e |rank2[b, e, k]|  (first maximum, as torch)   (:245)
//   label[b, k]   = rank2[b, row, k]                                       (:247)
// The reference walks the K columns of every sample in Python with three .item() synchronisations per column;
// here one thread owns a column (lanes run along k: coalesced) and walks the E rows once.
CCSD_KERNEL void __launch_bounds__(256) cc_cells_kernel(const float *__restrict__ r2, uint8_t *__restrict__ present,
                                                       int *__restrict__ row, float *__restrict__ label, int B, int E, int K) {
  const size_t tot = (size_t)B * K;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < tot; g += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(g / K), k = (int)(g - (size_t)b * K);
    const float *col = r2 + (size_t)b * E * K + k;
    float best = -1.f, val = 0.f;
    int arg = 0;
    bool any = false;
    for (int e = 0; e < E; ++e) {
      const float v = col[(size_t)e * K];
      any = any || (v != 0.f);
      const float a = fabsf(v);
      if (a > best) { best = a; arg = e; val = v; }   // strict: keeps the FIRST maximum
    }
    present[g] = any ? 1 : 0;
    row[g] = arg;
    label[g] = val;
  }
}

// ---------------------------------------------------------------------------------------------
// ||z||^2 of the rank-2 corrector noise of every sample WITHOUT touching the state (solver.py:765-767): the draw is
// a pure function of (seed, sample, draw id, entry) -- or the injected stream -- and the mask of the flags.  One CTA =
// (APPLY_TN-cell tile t, sample b): partial sum -> norm_part slot t of the rank-2 object (fixed-order block reduction, so the
// Langevin step size is bit-reproducible).  Fully masked 4-cell groups / edge rows cost nothing.
struct ZnormArgs {
  const float *flags;
  const float *noise;                  // raw normals [B,E,K] of this draw or nullptr (Philox)
  const unsigned long long *zmask;     // [B]
  float *norm_part;
  int slot;
  NoiseCtx nz;
};

CCSD_KERNEL void __launch_bounds__(256) znorm_kernel(const DevPlan *__restrict__ P, ZnormArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const int N = d.N, E = d.E, K = d.K, Kg = P->Kp >> 2;
  constexpr int GPT = APPLY_TN / 4;   // 4-cell groups per tile
  const int b = blockIdx.y;
  const int ntiles = (K + APPLY_TN - 1) / APPLY_TN;      // CTA t takes the tiles t, t + gridDim.x, ... (gridDim.x = norm slots)
  const float *fl = a.flags + (size_t)b * N;
  const unsigned long long zm = a.zmask[b];
  const unsigned long long gs = (unsigned long long)(a.nz.sample_offset + b);
  const uint32_t did = draw_id(2, nz_step(a.nz), a.slot);
  // staged once per CTA: edge flags [E]; per tile: the live-cell bits of its 4-cell groups [GPT]
  float *fes = sm + 40;
  unsigned *gm = reinterpret_cast<unsigned *>(fes + ((E + 3) & ~3));
  for (int e = threadIdx.x; e < E; e += blockDim.x) fes[e] = fl[P->edge_ij[2 * e]] * fl[P->edge_ij[2 * e + 1]];
  float z2 = 0.f;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    __syncthreads();   // fes ready / the previous tile's gm no longer read
    for (int g = threadIdx.x; g < GPT; g += blockDim.x) {
      unsigned m = 0;
      for (int q = 0; q < 4; ++q) {
        const int k = (t * GPT + g) * 4 + q;
        if (k < K && !(P->cell_mask[k] & zm)) m |= 1u << q;
      }
      gm[g] = m;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < E * GPT; it += blockDim.x) {
      const int e = it / GPT, g = it - e * GPT;
      const unsigned m = gm[g];
      if (m == 0u || fes[e] == 0.f) continue;
      const int kg = t * GPT + g, k = kg * 4;
      float z4[4];
      if (a.noise) {
#pragma unroll
        for (int q = 0; q < 4; ++q) z4[q] = k + q < K ? a.noise[((size_t)b * E + e) * K + k + q] : 0.f;
      } else {
        normal4(a.nz.seed, gs, did, (uint32_t)(e * Kg + kg), z4);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if ((m >> q) & 1u) z2 += z4[q] * z4[q];
    }
  }
  z2 = block_sum(z2, sm);
  if (threadIdx.x == 0) a.norm_part[((size_t)(2 * d.B + b) * P->ntile_max + blockIdx.x) * 2 + 1] = z2;
}

// ---------------------------------------------------------------------------------------------
// Large complexes (E > 192): the Gram-quantity norm of tc_hnorm.cuh with H . H taken by the K-chunked GEMM
// (tc_r2big.cuh, H2 in a scratch matrix): the E^2 reductions and the norm.  One CTA per sample, fixed summation order.
struct HnormBigArgs {
  const float *H, *H2, *Dg, *Rs, *flags;
  float *norm_part;
  int step;
};

CCSD_KERNEL void __launch_bounds__(256) hnorm_big_kernel(const DevPlan *__restrict__ P, HnormBigArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const int E = d.E, Ep = P->Ep, N = d.N, b = blockIdx.x;
  const float *H = a.H + (size_t)b * E * Ep, *H2 = a.H2 + (size_t)b * E * Ep, *D = a.Dg + (size_t)b * E, *R = a.Rs + (size_t)b * E;
  float t3 = 0.f, sfh = 0.f, sdd = 0.f, sh = 0.f, sff = 0.f, sf = 0.f;
  for (int e = threadIdx.x >> 5; e < E; e += blockDim.x >> 5) {      // one warp per row, lanes along the columns
#ifdef CCSD_EMU
    const int lane = 0, nl = 1;
#else
    const int lane = threadIdx.x & 31, nl = 32;
#endif
    for (int c = lane; c < E; c += nl) {
      const float h = H[(size_t)e * Ep + c], h2 = H2[(size_t)e * Ep + c];
      t3 += h2 * h;
      sh += h * R[c];
      if (c == e) { sfh += h2; sdd += D[e] * h2; sff += D[e]; sf += R[e]; }
    }
  }
  t3 = block_sum(t3, sm); sfh = block_sum(sfh, sm); sdd = block_sum(sdd, sm); sh = block_sum(sh, sm);
  sff = block_sum(sff, sm); sf = block_sum(sf, sm);
  if (threadIdx.x == 0) {
    const ccsd_netf_t &Fn = d.netf;
    const float fa = Fn.aff[0], fb = Fn.aff[1], fc = Fn.aff[2];
    const float sc = P->sched[a.step * 3 + 2].score_scale;
    int n = 0;
    for (int i = 0; i < N; ++i) n += a.flags[(size_t)b * N + i] != 0.f;
    float nkc = 0.f;
    for (int dd = d.d_min; dd <= d.d_max; ++dd) {
      float cmb = dd <= n ? 1.f : 0.f;
      for (int q = 0; q < dd && dd <= n; ++q) cmb = cmb * (float)(n - q) / (float)(q + 1);
      nkc += cmb;
    }
    const float ne = 0.5f * (float)n * (float)(n - 1);
    const float shh = t3 + sdd;
    const float s2 = sc * sc * (fa * fa * sff + fb * fb * shh + 2.f * fa * fb * sfh + 2.f * fa * fc * sf + 2.f * fb * fc * sh + fc * fc * ne * nkc);
    float *np = a.norm_part + ((size_t)(2 * d.B + b) * P->ntile_max) * 2;
    np[0] = s2 > 0.f ? s2 : 0.f;
    for (int t = 1; t < P->ntile_r2; ++t) np[2 * t] = 0.f;
  }
}

}  // namespace ccsd
