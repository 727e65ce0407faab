// big_pipe.cuh -- ScoreNetworkX and ScoreNetworkA for LARGE graphs (max_node_num > 64: grid N = 361,
// ENZYMES N = 125), graph-only configs.  Same reference functions as xa_pipe.cuh (ScoreNetwork_X.py:102-133,
// ScoreNetwork_A.py:505-541, attention.py:84-132,270-304, layers.py:115-158, graph_utils.py:274-292), other
// shape regime: one graph no longer fits a CTA's shared memory (54 adjacency channels x 361^2 = 28 MB), so
// every phase is tiled over (row chunk | channel | graph) CTAs and the channel stack lives in HBM / L2 as
// full N x Np planes (Np = N rounded up to 8).
//
// All the contractions are expressed through the feature-major register-tile primitive dense_fm (prims.cuh):
//   out(r, o) = sum_k in(k, r) W[k, o]
//   * A^c = A^(c-1) A        in(k, r) = A^(c-1)[k][r] (symmetric), W = A            (pow_tensor)
//   * x W_{q|k|v}            in = node features, feature-major [kin][Np]            (DenseGCNConv transform)
//   * A^ (x W)               in(k, r) = adj_c[k][r] (symmetric), W = the scaled transform Y, node-major
//                            (DenseGCNConv aggregation; the unit diagonal of A^ is a rank-one fix-up)
//   * the per-edge MLPs      in = channel planes (plane stride = N Np), rows = 64-pair segments of one row
// Symmetry: every adjacency-shaped tensor is symmetric, so the per-pair kernels only visit the 64-column
// segments of row i that reach the diagonal (j0 + 63 >= i) and write (i, j) and (j, i).
//
// Launch sequence of one evaluation (host: launch_xa_big):
//   prep -> pow (c_init - 1) -> deg -> [X net: depth x (xw, agg) -> xfin]
//        -> L x (xw, agg, attn, node, edge, deg) -> final
#pragma once
#include "xa_pipe.cuh"

namespace ccsd {

constexpr int BIG_RC = 32;     // node rows per CTA of the row-chunk kernels
constexpr int BIG_RCX = 8;     // node rows per CTA of the ScoreNetworkX final MLP (few rows per graph: many small CTAs)
constexpr int BIG_SEG = 64;    // node pairs (columns of one row) per CTA of the per-pair kernels

struct BigArgs {
  XaArgs a;
  float *base;        // scratch [B][big.g_total]
  int layer;          // attention layer
  int ch_in, ch_out;  // first input / output plane of the stack
  const float *xin;   // node features read  (feature-major [kin][Np], per-graph stride big.g_xf)
  float *xout;        // node features written
  int c;              // pow: plane to produce
  int xmode;          // 0: attention layer `layer`, 1: GCN layer `gk` of ScoreNetworkX
  int gk, in_row, out_row;   // X net: GCN layer, first input / output row of the feature concat HC
  int nch;            // deg: number of planes starting at ch_in
};

__device__ __forceinline__ float *big_ptr(const DevPlan *P, const BigArgs &g, int b, int off) {
  return g.base + (size_t)b * P->xp.big_total + off;
}

// plane 0 of the stack <- adj; node features (feature-major) <- x, also rows [0, F) of the X net's concat
CCSD_KERNEL void __launch_bounds__(256) big_prep_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.y, N = d.N, F = d.F, Np = L.big_Np;
  float *S = big_ptr(P, g, b, L.big_S), *xf = big_ptr(P, g, b, L.big_XF0), *hc = big_ptr(P, g, b, L.big_HC);
  const int tot = N * Np + F * Np;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < tot; p += gridDim.x * blockDim.x) {
    if (p < N * Np) {
      const int i = p / Np, j = p - i * Np;
      S[p] = j < N ? g.a.adj[((size_t)b * N + i) * N + j] : 0.f;
    } else {
      const int q = p - N * Np, f = q / Np, i = q - f * Np;
      const float v = i < N ? g.a.x[((size_t)b * N + i) * F + f] : 0.f;
      xf[q] = v;
      hc[q] = v;
    }
  }
}

// pow_tensor (graph_utils.py:274-292): plane c = plane (c-1) . plane 0
CCSD_KERNEL void __launch_bounds__(128) big_pow_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, N = P->d.N, Np = L.big_Np, PS = L.big_PS;
  const int i0 = blockIdx.x * BIG_RC, R = (N - i0 < BIG_RC) ? N - i0 : BIG_RC;
  float *S = big_ptr(P, g, b, L.big_S);
  dense_fm(S + (size_t)(g.c - 1) * PS + i0, Np, N, nullptr, 0, 0, S, nullptr, N, S + (size_t)g.c * PS + (size_t)i0 * Np, Np, 1, R,
           ACT_NONE);
}

// DenseGCNConv degrees (layers.py:139-146): d_i = clamp(sum_j A^_ij, 1)^-1/2 with the unit diagonal of A^
CCSD_KERNEL void __launch_bounds__(128) big_deg_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, c = blockIdx.y, N = P->d.N, Np = L.big_Np;
  const float *pl = big_ptr(P, g, b, L.big_S) + (size_t)(g.ch_in + c) * L.big_PS;
  float *dv = big_ptr(P, g, b, L.big_DV) + c * Np;
  const int i1 = (blockIdx.x + 1) * 128 < N ? (blockIdx.x + 1) * 128 : N;
  for (int i = blockIdx.x * 128 + threadIdx.x; i < i1; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += (j == i) ? 1.f : pl[(size_t)j * Np + i];   // column i = row i (symmetric), coalesced
    dv[i] = 1.0f / sqrtf(fmaxf(s, 1.f));
  }
}

// Y = diag(d) (x W): the feature transforms of one channel's Q | K | V convolutions (xmode 0, one node-major
// buffer [N][2 adp + nhp]) or of one GCN layer of ScoreNetworkX (xmode 1), pre-scaled by d_j for the aggregation.
// One pass: item = (8 output columns, 4 rows); the item picks its weight matrix and scales its own tile.
CCSD_KERNEL void __launch_bounds__(128) big_xw_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, c = blockIdx.y, N = d.N, Np = L.big_Np;
  const int i0 = blockIdx.x * BIG_RC, R = (N - i0 < BIG_RC) ? N - i0 : BIG_RC, ngrp = (R + 3) >> 2;
  const float *W = P->W;
  const float *dv = big_ptr(P, g, b, L.big_DV) + c * Np;
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const ccsd_gcn_t &gc = d.netx.gcn[g.gk];
  const int ad = ly.attn_dim, nh = ly.conv_out, adp = round_up(ad, 8), nhp = round_up(nh, 8), w2 = 2 * adp;
  const int YW = g.xmode == 0 ? w2 + nhp : round_up(gc.dout, 8);
  const int kin = g.xmode == 0 ? ly.conv_in : gc.din;
  const float *in = g.xmode == 0 ? g.xin + (size_t)b * L.big_total : big_ptr(P, g, b, L.big_HC) + (size_t)g.in_row * Np;
  float *y = big_ptr(P, g, b, L.big_Y) + (g.xmode == 0 ? (size_t)c * N * YW : 0);
  for (int it = threadIdx.x; it < (YW >> 3) * ngrp; it += blockDim.x) {
    const int cc = it / ngrp, r0 = i0 + ((it - cc * ngrp) << 2), o0 = cc << 3;
    const float *Wm;
    int Opad, oc, valid;
    if (g.xmode == 1) { Wm = W + gc.w; Opad = YW; oc = o0; valid = gc.dout - oc; }
    else if (o0 < adp) { Wm = W + ly.q[c].w; Opad = adp; oc = o0; valid = ad - oc; }
    else if (o0 < w2) { Wm = W + ly.k[c].w; Opad = adp; oc = o0 - adp; valid = ad - oc; }
    else { Wm = W + ly.v[c].w; Opad = nhp; oc = o0 - w2; valid = nh - oc; }
    float acc[4][8];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[rr][j] = 0.f;
    dense_tile(acc, in, Np, kin, nullptr, 0, 0, Wm, Opad, r0, oc);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int i = r0 + rr;
      if (i < N) {
        const float di = dv[i];
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = j < valid ? acc[rr][j] * di : 0.f;
        float4 *dst = reinterpret_cast<float4 *>(y + (size_t)i * YW + o0);
        dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
        dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
    }
  }
}

// T = diag(d) A^ Y + bias: the aggregation of DenseGCNConv (layers.py:147-156).  A^ = plane with a unit
// diagonal: sum_j a_ij y_j over the stored plane plus the rank-one fix-up (1 - a_ii) y_i, applied by the item
// that owns the tile (no second pass).  Item = 8 rows x 8 columns in registers: four 16-byte loads per 64 FMAs
// (the operands come through L1, whose bandwidth -- not the FMA pipe -- bounds the 4 x 8 tile), software pipelined.
// xmode 0: Q | K rows -> TQK (feature-major [2 adp][Np]), V -> TV ([c nh + o][Np]);  xmode 1: tanh -> HC rows.
constexpr int BIG_RCA = 64;    // node rows per CTA of the aggregation kernel

struct AggOps { float4 a0, a1, w0, w1; };
__device__ __forceinline__ void agg_load(AggOps &t, const float *pa, const float *__restrict__ pw) {
  t.a0 = ld4(pa); t.a1 = ld4(pa + 4);
  t.w0 = __ldg(reinterpret_cast<const float4 *>(pw)); t.w1 = __ldg(reinterpret_cast<const float4 *>(pw + 4));
}
__device__ __forceinline__ void agg_fma(float acc[8][8], const AggOps &t) {
  const float av[8] = {t.a0.x, t.a0.y, t.a0.z, t.a0.w, t.a1.x, t.a1.y, t.a1.z, t.a1.w};
  const float wv[8] = {t.w0.x, t.w0.y, t.w0.z, t.w0.w, t.w1.x, t.w1.y, t.w1.z, t.w1.w};
#pragma unroll
  for (int rr = 0; rr < 8; ++rr)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[rr][j] += av[rr] * wv[j];
}

CCSD_KERNEL void __launch_bounds__(128) big_agg_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, c = blockIdx.y, N = d.N, Np = L.big_Np;
  const int i0 = blockIdx.x * BIG_RCA, R = (N - i0 < BIG_RCA) ? N - i0 : BIG_RCA, ngrp = (R + 7) >> 3;
  const float *W = P->W;
  const float *dv = big_ptr(P, g, b, L.big_DV) + c * Np;
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const ccsd_gcn_t &gc = d.netx.gcn[g.gk];
  const int ad = ly.attn_dim, nh = ly.conv_out, adp = round_up(ad, 8), nhp = round_up(nh, 8), w2 = 2 * adp;
  const int YW = g.xmode == 0 ? w2 + nhp : round_up(gc.dout, 8);
  const float *pl = big_ptr(P, g, b, L.big_S) + (g.xmode == 0 ? (size_t)(g.ch_in + c) * L.big_PS : 0);
  const float *y = big_ptr(P, g, b, L.big_Y) + (g.xmode == 0 ? (size_t)c * N * YW : 0);
  float *tqk = big_ptr(P, g, b, L.big_TQK) + (size_t)c * w2 * Np, *tv = big_ptr(P, g, b, L.big_TV) + (size_t)c * nh * Np;
  float *hc = big_ptr(P, g, b, L.big_HC) + (size_t)g.out_row * Np;
  for (int it = threadIdx.x; it < (YW >> 3) * ngrp; it += blockDim.x) {
    const int cc = it / ngrp, r0 = i0 + ((it - cc * ngrp) << 3), o0 = cc << 3;   // r0 + 7 < Np (Np is a multiple of 8)
    float acc[8][8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[rr][j] = 0.f;
    const float *pa = pl + r0, *pw = y + o0;
    AggOps A, Bq;
    agg_load(A, pa, pw);
    int k = 1;
#pragma unroll 1
    for (; k + 1 < N; k += 2) {
      agg_load(Bq, pa + (size_t)k * Np, pw + (size_t)k * YW);
      agg_fma(acc, A);
      agg_load(A, pa + (size_t)(k + 1) * Np, pw + (size_t)(k + 1) * YW);
      agg_fma(acc, Bq);
    }
    agg_fma(acc, A);
    if (k < N) {
      agg_load(Bq, pa + (size_t)k * Np, pw + (size_t)k * YW);
      agg_fma(acc, Bq);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = o0 + j;
      float *dst;
      float bias;
      int act = ACT_NONE;
      if (g.xmode == 1) {
        if (o >= gc.dout) continue;
        dst = hc + (size_t)o * Np; bias = __ldg(W + gc.b + o); act = ACT_TANH;
      } else if (o < w2) {
        const int oo = o < adp ? o : o - adp;
        if (oo >= ad) continue;
        dst = tqk + (size_t)o * Np; bias = __ldg(W + (o < adp ? ly.q[c].b : ly.k[c].b) + oo);
      } else {
        if (o - w2 >= nh) continue;
        dst = tv + (size_t)(o - w2) * Np; bias = __ldg(W + ly.v[c].b + (o - w2));
      }
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const int i = r0 + rr;
        if (i < N) {
          const float fix = 1.f - pl[(size_t)i * Np + i];
          dst[i] = act_fast(dv[i] * (acc[rr][j] + fix * y[(size_t)i * YW + o]) + bias, act);
        }
      }
    }
  }
}

// Attention scores of one channel (attention.py:111-130): heads = chunks of ds = ad / heads features
// (torch.split), att(i, j) = mean_h 0.5 (tanh(q_i.k_j s) + tanh(q_j.k_i s)).  Item = 4x4 node block I <= J.
CCSD_KERNEL void __launch_bounds__(128) big_attn_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, c = blockIdx.y, N = d.N, Np = L.big_Np;
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const int ad = ly.attn_dim, adp = round_up(ad, 8), heads = d.neta.num_heads;
  const int ds = ad / heads, nch = (ad + ds - 1) / ds;
  const float inv = 0.5f / (float)nch, scale = 1.0f / sqrtf((float)ly.conv_out);
  const float *Q = big_ptr(P, g, b, L.big_TQK) + (size_t)c * 2 * adp * Np, *Kf = Q + (size_t)adp * Np;
  float *att = big_ptr(P, g, b, L.big_ATT) + (size_t)c * L.big_PS;
  const int nb = (N + 3) >> 2, nblk = nb * (nb + 1) / 2;
  const int it1 = (blockIdx.x + 1) * 128 < nblk ? (blockIdx.x + 1) * 128 : nblk;
  for (int it = blockIdx.x * 128 + threadIdx.x; it < it1; it += blockDim.x) {
    // row-major upper triangle of nb x nb blocks: rows before I hold I nb - I (I - 1) / 2 blocks
    int I = (int)(((float)(2 * nb + 1) - sqrtf((float)(2 * nb + 1) * (float)(2 * nb + 1) - 8.f * (float)it)) * 0.5f);
    if (I < 0) I = 0;
    if (I > nb - 1) I = nb - 1;
    while (I > 0 && I * nb - I * (I - 1) / 2 > it) --I;
    while ((I + 1) * nb - (I + 1) * I / 2 <= it) ++I;
    const int J = I + (it - (I * nb - I * (I - 1) / 2));
    const int i0 = I << 2, j0 = J << 2;
    float sum[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) sum[u][v] = 0.f;
    for (int h = 0; h < nch; ++h) {
      float qa[4][4], qb[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) { qa[u][v] = 0.f; qb[u][v] = 0.f; }
      const int d0 = h * ds, d1 = (d0 + ds < ad) ? d0 + ds : ad;
      for (int dd = d0; dd < d1; ++dd) {   // (unrolling by 4 measured slower: 0.36 -> 0.40 ms at grid; the kernel is within 1.6x of its SFU floor)
        const float4 qi = ld4(Q + (size_t)dd * Np + i0), kj = ld4(Kf + (size_t)dd * Np + j0);
        const float4 qj = ld4(Q + (size_t)dd * Np + j0), ki = ld4(Kf + (size_t)dd * Np + i0);
        const float qiv[4] = {qi.x, qi.y, qi.z, qi.w}, kjv[4] = {kj.x, kj.y, kj.z, kj.w};
        const float qjv[4] = {qj.x, qj.y, qj.z, qj.w}, kiv[4] = {ki.x, ki.y, ki.z, ki.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) { qa[u][v] += qiv[u] * kjv[v]; qb[u][v] += qjv[v] * kiv[u]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) sum[u][v] += fast_tanh(qa[u][v] * scale) + fast_tanh(qb[u][v] * scale);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = i0 + u, j = j0 + v;
        if (i <= j && j < N) {
          const float o = inv * sum[u][v];
          att[(size_t)i * Np + j] = o;   // upper triangle only: the edge kernels read att at j >= i
        }
      }
  }
}

// node branch of AttentionLayer (attention.py:292-293): x_out = tanh(mask_x(MLP(cat_c V_c)))
CCSD_KERNEL void __launch_bounds__(128) big_node_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, N = d.N, Np = L.big_Np;
  const int i0 = blockIdx.x * BIG_RC, R = (N - i0 < BIG_RC) ? N - i0 : BIG_RC;
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const ccsd_mlp_t &mc = ly.multi_channel;
  const int nh = ly.conv_out, hid = mc.nl > 1 ? mc.dhid : 1;
  float *so = sm, *hA = sm + round_up(nh, 4) * BIG_RC, *hB = mc.nl > 2 ? hA + hid * BIG_RC : hA;
  const float *tv = big_ptr(P, g, b, L.big_TV);
  mlp_fm(mc, P->W, tv + i0, Np, ly.c_in * nh, nullptr, 0, 0, R, hA, hB, BIG_RC, so, 1, BIG_RC, ACT_ELU, ACT_NONE);
  float *xo = g.xout + (size_t)b * L.big_total;
  for (int p = threadIdx.x; p < nh * R; p += blockDim.x) {
    const int o = p / R, r = p - o * R, i = i0 + r;
    xo[(size_t)o * Np + i] = fast_tanh(so[o * BIG_RC + r] * g.a.flags[(size_t)b * N + i]);
  }
}

// segment of row i that this CTA of a per-pair kernel owns; false when it lies wholly below the diagonal
__device__ __forceinline__ bool big_segment(const DevPlan *P, int &i, int &j0, int &R) {
  const int nseg = P->xp.big_nseg, N = P->d.N;
  i = blockIdx.x / nseg;
  j0 = (blockIdx.x - i * nseg) * BIG_SEG;
  R = (N - j0 < BIG_SEG) ? N - j0 : BIG_SEG;
  return j0 + BIG_SEG > i;
}

// edge branch of AttentionLayer (attention.py:295-303): M = MLP([att_1..att_c, adj_1..adj_c]),
// adj_out = mask_adjs(M + M^T) = 2 M mask (M is symmetric because its inputs are)
CCSD_KERNEL void __launch_bounds__(128) big_edge_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, N = d.N, Np = L.big_Np, PS = L.big_PS;
  int i, j0, R;
  if (!big_segment(P, i, j0, R)) return;
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const ccsd_mlp_t &m = ly.mlp;
  const int hid = m.nl > 1 ? m.dhid : 1;
  float *so = sm, *hA = sm + ly.c_out * BIG_SEG, *hB = m.nl > 2 ? hA + hid * BIG_SEG : hA;
  float *S = big_ptr(P, g, b, L.big_S);
  const float *att = big_ptr(P, g, b, L.big_ATT);
  const size_t t0 = (size_t)i * Np + j0;
  mlp_fm(m, P->W, att + t0, PS, ly.c_in, S + (size_t)g.ch_in * PS + t0, PS, ly.c_in, R, hA, hB, BIG_SEG, so, 1, BIG_SEG, ACT_ELU,
         ACT_NONE);
  const float fi = g.a.flags[(size_t)b * N + i];
  for (int p = threadIdx.x; p < ly.c_out * R; p += blockDim.x) {
    const int o = p / R, r = p - o * R, j = j0 + r;
    if (j < i) continue;
    const float v = 2.0f * so[o * BIG_SEG + r] * fi * g.a.flags[(size_t)b * N + j];
    float *pl = S + (size_t)(g.ch_out + o) * PS;
    pl[(size_t)i * Np + j] = v;
    pl[(size_t)j * Np + i] = v;
  }
}

// Same edge branch with ONE NODE PAIR PER THREAD (every width <= 16: c <= 8 channels, hidden = 2 max(c_in, c_out)):
// the 2 c_in inputs and the hidden vector live in registers, the Linears are staged in shared memory as zero-padded
// [16][16] blocks read as broadcast 16-byte loads.  ~520 instructions per pair instead of ~1950 through the row-tile
// primitive (whose per-item epilogue dominates at K = 16).  CTA = 128 columns of row i; segments below the diagonal exit.
constexpr int BIG_ESEG = 128;
constexpr int BIG_EROWS = 4;    // rows per CTA (amortises the weight staging)
constexpr int BIG_EW = 16 * 16 + 16;   // floats per staged Linear
static inline bool big_edge_fast_ok(const ccsd_attn_layer_t &ly) {
  return 2 * ly.c_in <= 16 && ly.c_out <= 8 && ly.mlp.nl >= 1 && ly.mlp.nl <= 4 && (ly.mlp.nl == 1 || ly.mlp.dhid <= 16);
}

CCSD_KERNEL void __launch_bounds__(BIG_ESEG) big_edge_pair_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, N = d.N, Np = L.big_Np, PS = L.big_PS;
  const int nseg = (N + BIG_ESEG - 1) / BIG_ESEG;
  const int ib = (blockIdx.x / nseg) * BIG_EROWS, j0 = (blockIdx.x - (blockIdx.x / nseg) * nseg) * BIG_ESEG;
  if (j0 + BIG_ESEG <= ib) return;   // every row of this CTA lies below the segment
  const ccsd_attn_layer_t &ly = d.neta.layer[g.layer];
  const ccsd_mlp_t &m = ly.mlp;
  const float *W = P->W;
  // stage: Linear l as w[k][16] (k < 16) + b[16], zero padded
  for (int l = 0; l < m.nl; ++l) {
    const int din = l == 0 ? m.din : m.dhid, dout = l == m.nl - 1 ? m.dout : m.dhid, opad = round_up(dout, 8);
    for (int p = threadIdx.x; p < BIG_EW; p += blockDim.x) {
      float v = 0.f;
      if (p < 256) { const int k = p >> 4, o = p & 15; if (k < din && o < dout) v = __ldg(W + m.w[l] + k * opad + o); }
      else if (p - 256 < dout) v = __ldg(W + m.b[l] + (p - 256));
      sm[l * BIG_EW + p] = v;
    }
  }
  __syncthreads();
  float *S = big_ptr(P, g, b, L.big_S);
  const float *att = big_ptr(P, g, b, L.big_ATT);
  for (int r = threadIdx.x; r < BIG_ESEG * BIG_EROWS; r += blockDim.x) {
    const int i = ib + r / BIG_ESEG, j = j0 + r % BIG_ESEG;
    if (i >= N || j < i || j >= N) continue;
    const float fi = g.a.flags[(size_t)b * N + i];
    const size_t t = (size_t)i * Np + j;
    float h[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) h[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < ly.c_in) h[c] = att[(size_t)c * PS + t];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < ly.c_in) {
        const float v = S[(size_t)(g.ch_in + c) * PS + t];
        // input k = c_in + c: place with a static index
#pragma unroll
        for (int k = 1; k < 16; ++k)
          if (k == ly.c_in + c) h[k] = v;
      }
    for (int l = 0; l < m.nl; ++l) {
      const float *w = sm + l * BIG_EW;
      float t16[16];
#pragma unroll
      for (int o4 = 0; o4 < 4; ++o4) {
        const float4 bv = *reinterpret_cast<const float4 *>(w + 256 + 4 * o4);
        t16[4 * o4] = bv.x; t16[4 * o4 + 1] = bv.y; t16[4 * o4 + 2] = bv.z; t16[4 * o4 + 3] = bv.w;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int o4 = 0; o4 < 4; ++o4) {
          const float4 wv = *reinterpret_cast<const float4 *>(w + k * 16 + 4 * o4);
          t16[4 * o4] += h[k] * wv.x; t16[4 * o4 + 1] += h[k] * wv.y; t16[4 * o4 + 2] += h[k] * wv.z; t16[4 * o4 + 3] += h[k] * wv.w;
        }
      }
      const bool last = l == m.nl - 1;
#pragma unroll
      for (int o = 0; o < 16; ++o) h[o] = last ? t16[o] : fast_elu(t16[o]);   // padded outputs: elu(0) = 0
    }
    const float f2 = 2.0f * fi * g.a.flags[(size_t)b * N + j];
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < ly.c_out) {
        float *pl = S + (size_t)(g.ch_out + o) * PS;
        const float v = h[o] * f2;
        pl[t] = v;   // upper triangle; big_mirror_kernel fills the lower one with coalesced stores
      }
  }
}

// lower triangle <- upper triangle of planes [ch_out, ch_out + nch): 32 x 32 tiles through shared memory, so that both
// the reads and the writes are coalesced (a per-row kernel can only write its mirrored entries as scattered 4-byte stores)
CCSD_KERNEL void __launch_bounds__(256) big_mirror_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  CCSD_SMEM(sm);
  const XpLayout &L = P->xp;
  const int b = blockIdx.z, N = P->d.N, Np = L.big_Np;
  float *pl = big_ptr(P, g, b, L.big_S) + (size_t)(g.ch_out + blockIdx.y) * L.big_PS;
  const int nt = (N + 31) >> 5, it = blockIdx.x;
  int I = 0, rem = it;
  while (rem >= nt - I) { rem -= nt - I; ++I; }
  const int J = I + rem;
  for (int p = threadIdx.x; p < 1024; p += blockDim.x) {
    const int r = p >> 5, q = p & 31, i = I * 32 + r, j = J * 32 + q;
    sm[r * 33 + q] = (i < N && j < N) ? pl[(size_t)i * Np + j] : 0.f;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < 1024; p += blockDim.x) {
    const int r = p >> 5, q = p & 31, jj = J * 32 + r, ii = I * 32 + q;
    if (jj < N && ii < N && jj > ii) pl[(size_t)jj * Np + ii] = sm[q * 33 + r];
  }
}

// final per-edge MLP of ScoreNetworkA (ScoreNetwork_A.py:529-539) + the adjacency sampler epilogue
// (same arithmetic as afinal_kernel)
CCSD_KERNEL void __launch_bounds__(128) big_final_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const XaArgs &a = g.a;
  const int b = blockIdx.z, N = d.N, Np = L.big_Np, PS = L.big_PS;
  int i, j0, R;
  float *np = a.norm_part + ((size_t)(1 * d.B + b) * P->ntile_max + blockIdx.x) * 2;
  if (!big_segment(P, i, j0, R)) {
    if (a.mode == MODE_SCORE && threadIdx.x == 0) { np[0] = 0.f; np[1] = 0.f; }
    return;
  }
  const ccsd_mlp_t &m = d.neta.fin;
  const int hid = m.nl > 1 ? m.dhid : 1;
  float *so = sm, *red = sm + BIG_SEG, *hA = red + 40, *hB = m.nl > 2 ? hA + hid * BIG_SEG : hA;
  const float *S = big_ptr(P, g, b, L.big_S);
  mlp_fm(m, P->W, S + (size_t)i * Np + j0, PS, g.ch_out /* planes in the stack */, nullptr, 0, 0, R, hA, hB, BIG_SEG, so, 1, 0,
         ACT_ELU, ACT_NONE);
  const size_t ga = (size_t)b * N * N;
  const int stp = a.mode == MODE_EVAL ? 0 : nz_step(a.nz);
  const ccsd_objcoef_t ca = a.mode == MODE_EVAL ? ccsd_objcoef_t() : P->sched[stp * 3 + 1];
  const unsigned long long gsid = (unsigned long long)(a.nz.sample_offset + b);
  const float fi = a.flags[(size_t)b * N + i];
  float s2 = 0.f, z2 = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int j = j0 + r;
    if (j < i) continue;
    const float fij = fi * a.flags[(size_t)b * N + j];
    const float o = (i == j) ? 0.f : so[r] * fij;   // (1 - I) mask and mask_adjs
    if (a.mode == MODE_EVAL) {
      a.out_adj[ga + (size_t)i * N + j] = o;
      a.out_adj[ga + (size_t)j * N + i] = o;
      continue;
    }
    const float s = ca.score_scale * o;
    float z = 0.f;
    if (i != j) {
      const int q = i * N + j;
      z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gsid, draw_id(1, stp, a.slot), q)) * fij;
    }
    if (a.mode == MODE_SCORE) {
      a.out_adj[ga + (size_t)i * N + j] = s;
      if (i != j) {
        a.out_adj[ga + (size_t)j * N + i] = s;
        s2 += 2.f * s * s;
        z2 += 2.f * z * z;
      }
    } else {
      const float mu = ca.pa * a.adj[ga + (size_t)i * N + j] + ca.pb * s;
      const float v = mu + ca.pc * z;
      a.out_adj[ga + (size_t)i * N + j] = v;
      a.mean_adj[ga + (size_t)i * N + j] = mu;
      float *tja = a.nz.sd ? a.nz.sd->ta : a.traj_adj;
      if (tja && b == 0) tja[(size_t)i * N + j] = a.denoise ? mu : v;
      if (i != j) {
        a.out_adj[ga + (size_t)j * N + i] = v;
        a.mean_adj[ga + (size_t)j * N + i] = mu;
        if (tja && b == 0) tja[(size_t)j * N + i] = a.denoise ? mu : v;
      }
    }
  }
  if (a.mode == MODE_SCORE) {
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) { np[0] = s2; np[1] = z2; }
  }
}

// final MLP of ScoreNetworkX over a row chunk (ScoreNetwork_X.py:127-133) + the x sampler epilogue
// (same arithmetic as x_net_kernel)
CCSD_KERNEL void __launch_bounds__(128) big_xfin_kernel(const DevPlan *__restrict__ P, BigArgs g) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const XaArgs &a = g.a;
  const int b = blockIdx.z, N = d.N, F = d.F, Np = L.big_Np;
  const int i0 = blockIdx.x * BIG_RCX, R = (N - i0 < BIG_RCX) ? N - i0 : BIG_RCX;
  const ccsd_netx_t &X = d.netx;
  const ccsd_mlp_t &m = X.fin;
  const int hid = m.nl > 1 ? m.dhid : 1;
  float *so = sm, *red = sm + round_up(F, 4) * BIG_RCX, *hA = red + 40, *hB = m.nl > 2 ? hA + hid * BIG_RCX : hA;
  const float *hc = big_ptr(P, g, b, L.big_HC);
  mlp_fm(m, P->W, hc + i0, Np, X.fdim, nullptr, 0, 0, R, hA, hB, BIG_RCX, so, 1, BIG_RCX, ACT_ELU, ACT_NONE);
  const size_t gxo = (size_t)b * N * F;
  const int stp = a.mode == MODE_EVAL ? 0 : nz_step(a.nz);
  const ccsd_objcoef_t cx = a.mode == MODE_EVAL ? ccsd_objcoef_t() : P->sched[stp * 3 + 0];
  const unsigned long long gsid = (unsigned long long)(a.nz.sample_offset + b);
  float s2 = 0.f, z2 = 0.f;
  for (int q = threadIdx.x; q < R * F; q += blockDim.x) {
    const int r = q / F, f = q - r * F, i = i0 + r, p = i * F + f;
    const float fl = a.flags[(size_t)b * N + i];
    const float o = so[f * BIG_RCX + r] * fl;   // mask_x
    if (a.mode == MODE_EVAL) { a.out_x[gxo + p] = o; continue; }
    const float s = cx.score_scale * o;
    const float z = (a.noise_x ? a.noise_x[gxo + p] : normal1(a.nz.seed, gsid, draw_id(0, stp, a.slot), p)) * fl;
    if (a.mode == MODE_SCORE) {
      a.out_x[gxo + p] = s;
      s2 += s * s;
      z2 += z * z;
    } else {
      const float mu = cx.pa * a.x[gxo + p] + cx.pb * s;
      const float v = mu + cx.pc * z;
      a.out_x[gxo + p] = v;
      a.mean_x[gxo + p] = mu;
      if (b == 0) { float *tjx = a.nz.sd ? a.nz.sd->tx : a.traj_x; if (tjx) tjx[p] = a.denoise ? mu : v; }
    }
  }
  if (a.mode == MODE_SCORE) {
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(0 * d.B + b) * P->ntile_max + blockIdx.x) * 2;
      np[0] = s2; np[1] = z2;
    }
  }
}

}  // namespace ccsd
