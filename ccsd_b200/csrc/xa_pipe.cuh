// xa_pipe.cuh -- ScoreNetworkX and ScoreNetworkA / ScoreNetworkA_CC as a pipeline of small kernels with
// (graph x channel) grid parallelism, plus the x / adj sampler epilogues.
//
// Reference: ScoreNetwork_X.py:102-133, ScoreNetwork_A.py:505-541, ScoreNetwork_A_CC.py:275-332,
// attention.py:84-132,270-304, hodge_attention.py:80-129,290-325, layers.py:115-158.
//
// Why a pipeline: a graph is tiny (N <= 64), every phase of the networks has O(100) independent work
// items, and an attention layer is a chain of ~6 such phases PER CHANNEL.  One CTA per graph running the
// whole network keeps few warps per SM (the channel stack needs ~75 KB of shared memory) and spends its
// time in barrier bubbles (measured: 37 k cycles per channel for ~1.5 k cycles of FMAs).  Here the unit of
// work is small and there are many of them:
//   x_net_kernel          one CTA per graph            ScoreNetworkX + x epilogue; adjacency powers
//   attn_channel_kernel   one CTA per (graph, channel) GCN Q/K/V, attention scores, V's share of the node MLP
//   attn_finish_kernel    one CTA per graph            node MLP -> next node features; per-edge MLP -> next channels
//   hodge_kernel          one CTA per graph            hodge branch of ScoreNetworkA_CC (reduced form)
//   afinal_kernel         one CTA per (row chunk, graph) final per-edge MLP + adj epilogue
// Each CTA needs 15-45 KB of shared memory, so 5-12 of them share an SM and hide each other's barriers.
// Intermediates (channel stack, attention maps, node features: ~45 KB per graph) travel through global
// memory and stay in the 126 MB L2 for the batch sizes of interest.
//
// Every adjacency-shaped tensor is stored as its upper triangle (N(N+1)/2 pairs, "tri" index): the
// adjacency state is symmetric and so is everything derived from it (powers, symmetrised attention,
// M + M^T), which also halves the per-edge MLP work.
//
// Hodge branch (ScoreNetworkA_CC): the Hodge-dual adjacency built by adj_to_hodgedual
// (cc_utils.py:1503-1538) is diagonal and hodgedual_to_adj (cc_utils.py:1541-1588) reads only
// diagonals back, so layer 0 is a per-edge row scaling of the projections rank2 @ W_{q,k}
// (computed by the Gram kernel as P0) and the last hodge layer only needs diag(attention); its
// value branch is dead.  With two hodge layers the first layer's E x E attention output is
// materialised in shared memory and the second layer aggregates the projections P1 of the first
// layer's value output with it.
#pragma once
#include "prims.cuh"
#include "r2_kernels.cuh"   // zero_mask_of, Philox helpers of the sampler epilogues

#ifndef XP_MINB_C
#define XP_MINB_C 1
#endif
#ifndef XP_STAGE_W
#define XP_STAGE_W 1   // stage the channel's Q/K/V/W1 weights in shared memory (0: read them through L1/L2)
#endif
#ifndef XP_MINB_M
#define XP_MINB_M 1
#endif

namespace ccsd {

struct XaArgs {
  const float *x, *adj, *flags;  // [B,N,F] [B,N,N] [B,N]
  const float *P0, *P1;          // hodge projections [B,E,PR0] [B,E,PR1] (CC only)
  const float *r2;               // rank-2 state [B,E,K] (CC only: proj1_kernel reads it)
  int mode;                      // MODE_EVAL / MODE_SCORE / MODE_PRED
  int which;                     // bit0: evaluate X net, bit1: evaluate A net
  float *out_x, *out_adj;        // EVAL: raw net output; SCORE: scaled score; PRED: new state
  float *mean_x, *mean_adj;      // PRED: means
  float *norm_part;              // SCORE: [3][B][ntile_max][2]
  const float *noise_x, *noise_adj;  // raw normals for this draw ([B,...]) or nullptr (Philox)
  float *traj_x, *traj_adj;      // PRED: destination for sample 0 of this shard (or nullptr)
  int slot;                      // draw slot within the step
  int denoise;
  NoiseCtx nz;
  // pipeline scratch in global memory (per graph strides in XpLayout)
  float *g_stack, *g_att, *g_hmc, *g_x0, *g_x1;
  // per-launch (attention layers)
  int layer, ch_in, ch_out;
  const float *g_xin;            // node features read by this layer  [kin x N4] per graph
  float *g_xout;                 // node features written by this layer
  int gmh;                       // the attention layers of ScoreNetworkX_GMH (netx.glayer) instead of ScoreNetworkA's
  int gmh_phase;                 // x_net_kernel of ScoreNetworkX_GMH: 1 = adjacency powers + hand-over of x, 2 = final MLP + epilogue
  int skip_edge;                 // attn_finish_kernel: the per-edge MLP runs on the tensor cores (tc_edge.cuh) instead
  float *g_hcat;                 // not null: x_net_kernel stops after the GCN stack and writes [x, h_1 .. h_D] ([fdim x N4] per
                                 // graph) here for the tensor-core final MLP (tc_xfin.cuh)
  const float *g_hu;             // hodge_kernel: per-graph sums of the folded layer-1 projection weights over the live cells
                                 // ([B x n1]; they depend on the node flags only, hodge_u_kernel fills them once per run) or nullptr
  long long *trace;              // debug phase timeline of tc_attn_kernel (ccsd_debug_apply_trace), normally nullptr
};

__device__ __forceinline__ const ccsd_attn_layer_t &xa_layer(const DevPlan *P, const XaArgs &a) {
  return a.gmh ? P->d.netx.glayer[a.layer] : P->d.neta.layer[a.layer];
}
__device__ __forceinline__ int xa_heads(const DevPlan *P, const XaArgs &a) { return a.gmh ? P->d.netx.gmh_heads : P->d.neta.num_heads; }

// =============================================================================================
// x_net_kernel: ScoreNetworkX, x epilogue, adjacency powers + feature-major x for the A pipeline
// =============================================================================================
CCSD_KERNEL void __launch_bounds__(128) x_net_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const int b = blockIdx.x;
  const int N = d.N, F = d.F, N4 = L.N4, NT = L.NT, ldp = L.ldp;
  const float *W = P->W;
  float *flags = sm + L.x_flags, *dvec = sm + L.x_dvec, *adj = sm + L.x_adj, *an = sm + L.x_an, *x0 = sm + L.x_x0;
  float *sx = sm + L.x_sx, *red = sm + L.x_red;

  for (int i = threadIdx.x; i < N4; i += blockDim.x) flags[i] = i < N ? a.flags[(size_t)b * N + i] : 0.f;
  for (int p = threadIdx.x; p < F * N4; p += blockDim.x) {
    const int f = p / N4, i = p - f * N4;
    x0[p] = i < N ? a.x[((size_t)b * N + i) * F + f] : 0.f;
  }
  for (int t = threadIdx.x; t < NT; t += blockDim.x) {
    const int ij = P->tri_ij[t];
    adj[t] = a.adj[(size_t)b * N * N + (ij >> 8) * N + (ij & 255)];
  }
  __syncthreads();

  if ((a.which & 2) || a.gmh_phase == 1) {
    // pow_tensor (graph_utils.py:274-292): A^c = A^(c-1) . A, symmetric; channels 0..c_init-1 of the stack
    const int c0 = a.gmh_phase == 1 ? d.netx.gmh_c_init : d.neta.c_init;
    for (int c = 1; c < c0; ++c) {
      for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        const int ij = P->tri_ij[t], i = ij >> 8, j = ij & 255;
        float s = 0.f;
        for (int k = 0; k < N; ++k) s += adj[(c - 1) * ldp + tri_index_any(i, k, N)] * adj[tri_index_any(k, j, N)];
        adj[c * ldp + t] = s;
      }
      __syncthreads();
    }
    float *gs = a.g_stack + (size_t)b * L.g_stack;
    for (int p = threadIdx.x; p < c0 * ldp; p += blockDim.x) gs[p] = (p % ldp) < NT ? adj[p] : 0.f;
    float *gx = a.g_x0 + (size_t)b * L.g_x;
    for (int p = threadIdx.x; p < F * N4; p += blockDim.x) gx[p] = x0[p];
  }
  if (!(a.which & 1)) return;

  // ---- ScoreNetworkX ----
  const ccsd_netx_t &X = d.netx;
  float *hcat = sm + L.x_hcat, *ax = sm + L.x_ax, *hA = sm + L.x_ha, *hB = sm + L.x_hb;
  if (a.gmh_phase == 1) {   // ScoreNetworkX_GMH: the attention layers fill rows [F, fdim) of g_hcat; x itself is rows [0, F)
    float *gh = a.g_hcat + (size_t)b * (size_t)X.fdim * N4;
    for (int p = threadIdx.x; p < F * N4; p += blockDim.x) gh[p] = x0[p];
    return;
  }
  const float *in = x0;
  int din = F, row = 0;
  if (a.gmh_phase == 2) {   // the layers' (tanh'ed) node outputs come back from g_hcat
    const float *gh = a.g_hcat + (size_t)b * (size_t)X.fdim * N4 + F * N4;
    row = X.depth * X.nhid;
    for (int p = threadIdx.x; p < row * N4; p += blockDim.x) hcat[p] = gh[p];
    __syncthreads();
  } else
    gcn_norm_tri(adj, N, N4, dvec, an);
  for (int k = 0; k < (a.gmh_phase == 2 ? 0 : X.depth); ++k) {
    const ccsd_gcn_t &g = X.gcn[k];
    gcn_aggregate_fm(an, N, N4, in, din, ax);
    __syncthreads();
    dense_fm(ax, N4, din, nullptr, 0, 0, W + g.w, W + g.b, g.dout, hcat + row * N4, 1, N4, N, ACT_TANH);
    __syncthreads();
    in = hcat + row * N4;
    din = g.dout;
    row += g.dout;
  }
  if (a.g_hcat && a.gmh_phase == 0) {
    float *gh = a.g_hcat + (size_t)b * (size_t)X.fdim * N4;
    for (int p = threadIdx.x; p < F * N4; p += blockDim.x) gh[p] = x0[p];
    for (int p = threadIdx.x; p < row * N4; p += blockDim.x) gh[F * N4 + p] = hcat[p];
    return;
  }
  mlp_fm(X.fin, W, x0, N4, F, hcat, N4, row, N, hA, hB, N4, sx, 1, N4, ACT_ELU, ACT_NONE);
  for (int p = threadIdx.x; p < F * N4; p += blockDim.x) {
    const int i = p % N4;
    if (i < N) sx[p] *= flags[i];
  }
  __syncthreads();

  // ---- x epilogue ----
  const size_t gxo = (size_t)b * N * F;
  if (a.mode == MODE_EVAL) {
    for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
      const int i = p / F, f = p - i * F;
      a.out_x[gxo + p] = sx[f * N4 + i];
    }
    return;
  }
  const int stp = nz_step(a.nz);
  const ccsd_objcoef_t cx = P->sched[stp * 3 + 0];
  const unsigned long long gsid = (unsigned long long)(a.nz.sample_offset + b);
  if (a.mode == MODE_SCORE) {
    // scaled score + per-sample squared norms of score and (masked) noise (solver.py:693-699, 1299-1305)
    float s2 = 0.f, z2 = 0.f;
    for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
      const int i = p / F, f = p - i * F;
      const float s = cx.score_scale * sx[f * N4 + i];
      a.out_x[gxo + p] = s;
      const float z = (a.noise_x ? a.noise_x[gxo + p] : normal1(a.nz.seed, gsid, draw_id(0, stp, a.slot), p)) * flags[i];
      s2 += s * s;
      z2 += z * z;
    }
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(0 * d.B + b) * P->ntile_max) * 2;
      np[0] = s2; np[1] = z2;
    }
    return;
  }
  // MODE_PRED: mean = pa*obj + pb*score ; new = mean + pc*z   (solver.py:230-244, 386-398; sde.py:200-235)
  for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
    const int i = p / F, f = p - i * F;
    const float s = cx.score_scale * sx[f * N4 + i];
    const float z = (a.noise_x ? a.noise_x[gxo + p] : normal1(a.nz.seed, gsid, draw_id(0, stp, a.slot), p)) * flags[i];
    const float m = cx.pa * x0[f * N4 + i] + cx.pb * s;
    const float v = m + cx.pc * z;
    a.out_x[gxo + p] = v;
    a.mean_x[gxo + p] = m;
    if (b == 0) { float *tjx = a.nz.sd ? a.nz.sd->tx : a.traj_x; if (tjx) tjx[p] = a.denoise ? m : v; }
  }
}

// =============================================================================================
// attn_channel_kernel: Attention.forward of ONE channel of one graph (attention.py:84-132)
// =============================================================================================
CCSD_KERNEL void __launch_bounds__(128, XP_MINB_C) attn_channel_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_attn_layer_t &ly = xa_layer(P, a);
  const int heads = xa_heads(P, a);
  const int c = blockIdx.x, b = blockIdx.y;
  const int N = d.N, N4 = L.N4, NT = L.NT, ldp = L.ldp;
  const float *W = P->W;
  float *dvec = sm + L.c_dvec, *adjc = sm + L.c_adj, *an = sm + L.c_an, *xin = sm + L.c_xin, *ax = sm + L.c_ax;
  float *q = sm + L.c_q, *kf = sm + L.c_k, *v = sm + L.c_v, *atp = sm + L.c_atp;
  const int kin = ly.conv_in, ad = ly.attn_dim, nh = ly.conv_out;
  const float scale = 1.0f / sqrtf((float)nh);  // / sqrt(out_dim)  (attention.py:125)

  // this channel's weights (Q, K, V transforms and V's slice of the node MLP) stream into shared memory
  // while the normalisation and aggregation phases run
  const ccsd_mlp_t &mc = ly.multi_channel;
  const int o1 = mc.nl == 1 ? mc.dout : mc.dhid, o1p = round_up(o1, 8);
  const int adp8 = round_up(ad, 8), nhp8 = round_up(nh, 8);
  float *wq = sm + L.c_w, *wk = wq + kin * adp8, *wv = wk + kin * adp8, *w1 = wv + kin * nhp8;
#if XP_STAGE_W
  if (!ly.conv_mlp) {
    stage_async(wq, W + ly.q[c].w, kin * adp8);
    stage_async(wk, W + ly.k[c].w, kin * adp8);
  }
  stage_async(wv, W + ly.v[c].w, kin * nhp8);
  stage_async(w1, W + mc.w[0] + (size_t)c * nh * o1p, nh * o1p);
#else
  wq = const_cast<float *>(W + ly.q[c].w); wk = const_cast<float *>(W + ly.k[c].w); wv = const_cast<float *>(W + ly.v[c].w);
  w1 = const_cast<float *>(W + mc.w[0] + (size_t)c * nh * o1p);
#endif
  const float *gadj = a.g_stack + (size_t)b * L.g_stack + (size_t)(a.ch_in + c) * ldp;
  for (int t = threadIdx.x; t < ldp; t += blockDim.x) adjc[t] = gadj[t];
  const float *gx = a.g_xin + (size_t)b * L.g_x;
  for (int p = threadIdx.x; p < kin * N4; p += blockDim.x) xin[p] = gx[p];
  __syncthreads();
  gcn_norm_tri(adjc, N, N4, dvec, an);
  gcn_aggregate_fm(an, N, N4, xin, kin, ax);
  stage_wait();
  __syncthreads();
  {
    // Q | K | V = (A x) W_{q,k,v} + b as one item space (same input tile, three weight matrices)
    const int ngrp = N4 >> 2;
    const int nq = ly.conv_mlp ? 0 : (round_up(ad, 8) >> 3) * ngrp, nv = (round_up(nh, 8) >> 3) * ngrp;
    for (int it = threadIdx.x; it < 2 * nq + nv; it += blockDim.x) {
      const int w = it < nq ? 0 : (it < 2 * nq ? 1 : 2);
      const int li = it - (w == 0 ? 0 : (w == 1 ? nq : 2 * nq));
      const ccsd_gcn_t &g = w == 0 ? ly.q[c] : (w == 1 ? ly.k[c] : ly.v[c]);
      float *dst = w == 0 ? q : (w == 1 ? kf : v);
      const float *ws = w == 0 ? wq : (w == 1 ? wk : wv);
      const int O = g.dout, Opad = round_up(O, 8);
      const int chunk = li / ngrp, r0 = (li - chunk * ngrp) << 2, oc = chunk << 3;
      float acc[4][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float bv = __ldg(W + g.b + oc + j);
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) acc[rr][j] = bv;
      }
      dense_tile<XP_STAGE_W != 0>(acc, ax, N4, kin, nullptr, 0, 0, ws, Opad, r0, oc);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (oc + j < O) {
          // padded node rows are written too (finite values: the attention blocks read them)
          float *o = dst + (oc + j) * N4 + r0;
          o[0] = acc[0][j]; o[1] = acc[1][j]; o[2] = acc[2][j]; o[3] = acc[3][j];
        }
    }
  }
  __syncthreads();
  if (ly.conv_mlp) {
    // conv == "MLP" (attention.py:170-180): Q, K = 2-layer tanh MLPs of x alone (no adjacency); mlp_fm ends with a barrier
    float *mh = sm + L.c_mh;
    mlp_fm(ly.qm[c], W, xin, N4, kin, nullptr, 0, 0, N, mh, mh, N4, q, 1, N4, ACT_TANH, ACT_NONE);
    mlp_fm(ly.km[c], W, xin, N4, kin, nullptr, 0, 0, N, mh, mh, N4, kf, 1, N4, ACT_TANH, ACT_NONE);
    // rows [N, N4) of q / k feed only discarded outputs of the 4 x 4 score blocks: keep them finite
    for (int p = threadIdx.x; p < ad * (N4 - N); p += blockDim.x) {
      const int dd = p / (N4 - N), i = N + p - dd * (N4 - N);
      q[dd * N4 + i] = 0.f; kf[dd * N4 + i] = 0.f;
    }
    __syncthreads();
  }
  attn_scores_blk(q, kf, N, N4, ad, heads, scale, atp, ldp);
  {
    // V's share of the first Linear of multi_channel, which is linear in the channel concat
    // (attention.py:292): hmc_c(o, i) = sum_f V_c(i, f) W1[c*nh + f, o]; the finish kernel sums over c
    float *gh = a.g_hmc + (size_t)b * L.g_hmc + (size_t)c * L.mc_o1_max * N4;
    dense_fm<XP_STAGE_W != 0>(v, N4, nh, nullptr, 0, 0, w1, nullptr, o1, gh, 1, N4, N, ACT_NONE, false, (int)blockDim.x - 32);
  }
  __syncthreads();
  {
    const int ds = ad / heads, nch = (ad + ds - 1) / ds;
    float *ga = a.g_att + (size_t)b * L.g_att + (size_t)c * ldp;
    for (int t = threadIdx.x; t < ldp; t += blockDim.x) {
      float s = 0.f;
      if (t < NT)
        for (int h = 0; h < nch; ++h) s += atp[h * ldp + t];
      ga[t] = s;
    }
  }
}

// =============================================================================================
// attn_finish_kernel: the rest of AttentionLayer.forward (attention.py:292-302) for one graph
// =============================================================================================
CCSD_KERNEL void __launch_bounds__(128) attn_finish_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_attn_layer_t &ly = xa_layer(P, a);
  const int b = blockIdx.x;
  const int N = d.N, N4 = L.N4, NT = L.NT, ldp = L.ldp;
  const float *W = P->W;
  float *flags = sm + L.f_flags, *hs = sm + L.f_hs, *hs2 = sm + L.f_hs2, *ehA = sm + L.f_eha, *ehB = sm + L.f_ehb;
  const int nh = ly.conv_out;
  const ccsd_mlp_t &mc = ly.multi_channel;
  const int o1 = mc.nl == 1 ? mc.dout : mc.dhid;

  for (int i = threadIdx.x; i < N4; i += blockDim.x) flags[i] = i < N ? a.flags[(size_t)b * N + i] : 0.f;
  // node branch: x_out = tanh(mask_x(MLP(cat V)))  (attention.py:292-293)
  {
    const float *gh = a.g_hmc + (size_t)b * L.g_hmc;
    for (int p = threadIdx.x; p < o1 * N4; p += blockDim.x) {
      const int o = p / N4, i = p - o * N4;
      float s = 0.f;
      if (i < N) {
        s = __ldg(W + mc.b[0] + o);
        for (int c = 0; c < ly.c_in; ++c) s += gh[(size_t)c * L.mc_o1_max * N4 + p];
        if (mc.nl > 1) s = fast_elu(s);
      }
      hs[p] = s;
    }
    __syncthreads();
    float *cur = hs, *oth = hs2;
    for (int li = 1; li < mc.nl; ++li) {
      const bool last = li == mc.nl - 1;
      const int O = last ? mc.dout : mc.dhid;
      dense_fm(cur, N4, li == 1 ? o1 : mc.dhid, nullptr, 0, 0, W + mc.w[li], W + mc.b[li], O, oth, 1, N4, N,
               last ? ACT_NONE : ACT_ELU);
      __syncthreads();
      float *t = cur; cur = oth; oth = t;
    }
    float *gx = a.g_xout + (size_t)b * L.g_x;
    // ScoreNetworkX_GMH: one more tanh on the layer's node output (ScoreNetwork_X.py:300), kept for the final concat
    float *ghc = a.gmh ? a.g_hcat + (size_t)b * (size_t)d.netx.fdim * N4 + (size_t)(d.F + a.layer * nh) * N4 : nullptr;
    for (int p = threadIdx.x; p < nh * N4; p += blockDim.x) {
      const int i = p % N4;
      float v = i < N ? fast_tanh(cur[p] * flags[i]) : 0.f;
      if (ghc) { v = i < N ? fast_tanh(v) : 0.f; ghc[p] = v; }
      gx[p] = v;
    }
  }
  if (a.skip_edge) return;
  // edge branch: M = MLP(cat[A_1..A_c, adj_1..adj_c]) ; adj_out = mask_adjs(M + M^T) = 2 M mask (M symmetric)
  float *gs = a.g_stack + (size_t)b * L.g_stack;
  const float *ga = a.g_att + (size_t)b * L.g_att;
  mlp_fm(ly.mlp, W, ga, ldp, ly.c_in, gs + (size_t)a.ch_in * ldp, ldp, ly.c_in, NT, ehA, ehB, ldp, gs + (size_t)a.ch_out * ldp,
         1, ldp, ACT_ELU, ACT_NONE);   // ends with __syncthreads: this CTA's global writes are visible to it
  for (int p = threadIdx.x; p < ly.c_out * ldp; p += blockDim.x) {
    const int c = p / ldp, t = p - c * ldp;
    float *pl = gs + (size_t)(a.ch_out + c) * ldp + t;
    if (t < NT) {
      const int ij = P->tri_ij[t];
      *pl = 2.0f * *pl * flags[ij >> 8] * flags[ij & 255];
    } else {
      *pl = 0.f;
    }
  }
}

// =============================================================================================
// hodge_kernel: hodge branch of ScoreNetworkA_CC on the channel stack in global memory
// =============================================================================================
// mlp_attention of a hodge layer on a register vector of <= 8 channels -> <= 8 channels: the single-Linear case (every
// shipped two-layer checkpoint) runs on statically indexed registers, anything else through small_mlp
__device__ __forceinline__ void hodge_channel_mlp(const ccsd_mlp_t &m, const float *__restrict__ W, const float att[CCSD_MAX_CH],
                                                  float out[CCSD_MAX_CH]) {
  if (m.nl == 1 && m.din <= CCSD_MAX_CH && m.dout <= CCSD_MAX_CH) {
    const float *w = W + m.w[0], *bb = W + m.b[0];   // (in, out_pad = 8)
#pragma unroll
    for (int o = 0; o < CCSD_MAX_CH; ++o) out[o] = o < m.dout ? __ldg(bb + o) : 0.f;
#pragma unroll
    for (int c = 0; c < CCSD_MAX_CH; ++c)
      if (c < m.din) {
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(w + c * 8)), w1 = __ldg(reinterpret_cast<const float4 *>(w + c * 8 + 4));
        out[0] += att[c] * w0.x; out[1] += att[c] * w0.y; out[2] += att[c] * w0.z; out[3] += att[c] * w0.w;
        out[4] += att[c] * w1.x; out[5] += att[c] * w1.y; out[6] += att[c] * w1.z; out[7] += att[c] * w1.w;
      }
    return;
  }
  if (m.nl == 2 && m.din <= CCSD_MAX_CH && m.dhid <= 8 && m.dout <= CCSD_MAX_CH) {   // din -> hid (<= 8) -> dout, elu
    float h[8];
    {
      const float *w = W + m.w[0], *bb = W + m.b[0];
#pragma unroll
      for (int o = 0; o < 8; ++o) h[o] = o < m.dhid ? __ldg(bb + o) : 0.f;
#pragma unroll
      for (int c = 0; c < CCSD_MAX_CH; ++c)
        if (c < m.din) {
          const float4 w0 = __ldg(reinterpret_cast<const float4 *>(w + c * 8)), w1 = __ldg(reinterpret_cast<const float4 *>(w + c * 8 + 4));
          h[0] += att[c] * w0.x; h[1] += att[c] * w0.y; h[2] += att[c] * w0.z; h[3] += att[c] * w0.w;
          h[4] += att[c] * w1.x; h[5] += att[c] * w1.y; h[6] += att[c] * w1.z; h[7] += att[c] * w1.w;
        }
#pragma unroll
      for (int o = 0; o < 8; ++o) h[o] = o < m.dhid ? fast_elu(h[o]) : 0.f;
    }
    const float *w = W + m.w[1], *bb = W + m.b[1];
#pragma unroll
    for (int o = 0; o < CCSD_MAX_CH; ++o) out[o] = o < m.dout ? __ldg(bb + o) : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < m.dhid) {
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(w + c * 8)), w1 = __ldg(reinterpret_cast<const float4 *>(w + c * 8 + 4));
        out[0] += h[c] * w0.x; out[1] += h[c] * w0.y; out[2] += h[c] * w0.z; out[3] += h[c] * w0.w;
        out[4] += h[c] * w1.x; out[5] += h[c] * w1.y; out[6] += h[c] * w1.z; out[7] += h[c] * w1.w;
      }
    return;
  }
  float o32[SMALL_MAX];
  small_mlp(m, W, att, o32, ACT_ELU);
#pragma unroll
  for (int o = 0; o < CCSD_MAX_CH; ++o) out[o] = o < m.dout ? o32[o] : 0.f;
}

__device__ __forceinline__ float hodge_diag_att(const float *q, const float *k, int ad, int heads, float scale) {
  if (ad == 4 && heads == 2)   // every shipped checkpoint (adim_h = 4, num_heads_h = 2): no loops, no divisions
    return 0.5f * (fast_tanh((q[0] * k[0] + q[1] * k[1]) * scale) + fast_tanh((q[2] * k[2] + q[3] * k[3]) * scale));
  const int ds = ad / heads;
  const int nch = (ad + ds - 1) / ds;
  float s = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int d0 = c * ds, d1 = (d0 + ds < ad) ? d0 + ds : ad;
    float a = 0.f;
    for (int dd = d0; dd < d1; ++dd) a += q[dd] * k[dd];
    s += fast_tanh(a * scale);
  }
  return s / (float)nch;
}

// u[b][r] = sum over the live cells of sample b of the folded layer-1 projection weights (see hodge_kernel): a function of the
// node flags only, so it is computed once per sampler run (ccsd_plan_init) instead of in every evaluation.  Same summation
// order as the in-kernel fallback (lanes along the cells, warp tree, sum over warps): bit-identical.
CCSD_KERNEL void __launch_bounds__(128) hodge_u_kernel(const DevPlan *__restrict__ P, const float *__restrict__ flags_g, float *__restrict__ g_hu, int nthr) {
  CCSD_SMEM(sm);   // [32] partial sums + [64] flags
  float *part = sm, *fl = sm + 32;
  const ccsd_plan_desc_t &d = P->d;
  const ccsd_neta_t &A = d.neta;
  const int b = blockIdx.x, N = d.N, K = d.K, Kw = P->Kp, PR1 = A.n_proj_rows[1];
  // the reduction tree must match hodge_kernel's block size (nthr): threads beyond it idle
  for (int i = threadIdx.x; i < 64; i += blockDim.x) fl[i] = i < N ? flags_g[(size_t)b * N + i] : 0.f;
  __syncthreads();
  const unsigned long long zm = zero_mask_of(fl, N);
  const float *Wp = P->W + A.proj_w + (size_t)P->PR0h * Kw;
#ifdef CCSD_EMU
  const int lane = 0, warp = 0, nwarp = 1;
#else
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (nthr + 31) >> 5;
#endif
  for (int r0 = 0; r0 < PR1; r0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#ifdef CCSD_EMU
    const int stride = blockDim.x;   // (host emulation: one thread per block walks every cell, like hodge_kernel's own loop)
#else
    const int stride = nthr;
#endif
    if ((int)threadIdx.x < stride)
      for (int k = threadIdx.x; k < K; k += stride) {
        if (!(P->cell_mask[k] & zm)) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (r0 + j < PR1) acc[j] += __ldg(Wp + (size_t)(r0 + j) * Kw + k);
        }
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#ifndef CCSD_EMU
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
#endif
      if (lane == 0 && warp < 4) part[warp * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 8 && r0 + j < PR1; j += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < nwarp; ++w) t += part[w * 8 + j];
      g_hu[(size_t)b * PR1 + r0 + j] = t;
    }
    __syncthreads();
  }
}

CCSD_KERNEL void __launch_bounds__(128) hodge_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_neta_t &A = d.neta;
  const int b = blockIdx.x;
  const int N = d.N, E = d.E, N4 = L.N4, ldp = L.ldp, NT = L.NT;
  const float *W = P->W;
  float *flags = sm + L.h_flags;
  float *stack = a.g_stack + (size_t)b * L.g_stack;
  const int ch_hodge0 = a.ch_in;   // first hodge channel of the stack
  const float scale = 1.0f / sqrtf((float)d.K);  // HodgeAttention out_dim = K (hodge_attention.py:236-239)
  const int c0 = A.c_init;
  const ccsd_hodge_layer_t &h0 = A.hodge[0];
  const int ad0 = h0.attn_dim;
  const int PR0 = P->PR0;
  const float *P0 = a.P0 + (size_t)b * E * PR0;

  for (int i = threadIdx.x; i < N4; i += blockDim.x) flags[i] = i < N ? a.flags[(size_t)b * N + i] : 0.f;
  // channels [ch_hodge0, ch_hodge0 + c0): hodgedual_to_adj(adj_to_hodgedual(adjc)) = adjc with zero diagonal;
  // the hodge output channels start as zero (diagonals stay zero, off-diagonals are filled per edge)
  {
    const int nout = h0.c_out + (A.num_layers_h == 2 ? A.hodge[1].c_out : 0);
    for (int p = threadIdx.x; p < (c0 + nout) * ldp; p += blockDim.x) {
      const int c = p / ldp, t = p - c * ldp;
      float v = 0.f;
      if (c < c0 && t < NT) {
        const int ij = P->tri_ij[t];
        if ((ij >> 8) != (ij & 255)) v = stack[c * ldp + t];
      }
      stack[(ch_hodge0 + c) * ldp + t] = v;
    }
  }
  __syncthreads();

  if (A.num_layers_h == 1) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
      const int t = tri_index(i, j, N);
      const float fe = flags[i] * flags[j];
      float att[CCSD_MAX_CH], q[SMALL_MAX], k[SMALL_MAX], out[SMALL_MAX];
      for (int c = 0; c < c0; ++c) {
        const float av = stack[c * ldp + t];
        const float dg = 1.0f / sqrtf(fmaxf(av, 1.f));
        const float nrm = dg * av * dg;
        const float *pq = P0 + (size_t)e * PR0 + h0.proj_row + (c * 2 + 0) * ad0;
        const float *pk = pq + ad0;
        for (int dd = 0; dd < ad0; ++dd) {
          q[dd] = nrm * pq[dd] + __ldg(W + h0.bq[c] + dd);
          k[dd] = nrm * pk[dd] + __ldg(W + h0.bk[c] + dd);
        }
        att[c] = hodge_diag_att(q, k, ad0, A.num_heads_h, scale);
      }
      small_mlp(h0.mlp_attention, W, att, out, ACT_ELU);
      for (int c = 0; c < h0.c_out; ++c) stack[(ch_hodge0 + c0 + c) * ldp + t] = 2.0f * tanhf(fe * fe * out[c]);
    }
    return;
  }

  // ---- two hodge layers ----
  const ccsd_hodge_layer_t &h1 = A.hodge[1];
  const int c1 = h0.c_out, ad1 = h1.attn_dim, lde = L.lde;
  float *hq = sm + L.h_hq, *hk = sm + L.h_hk, *H1 = sm + L.h_h1, *hdeg = sm + L.h_hdeg;
  int PR1 = P->PR1;
  const float *P1 = a.P1 + (size_t)b * E * PR1;
  if (P->p1_fold) {
    // Layer 0's value MLP is one Linear, so its output is rank2' = mask (alpha_e F + beta) with
    // alpha_e = sum_c w_c (A^c)_ij, and the layer-1 projections are
    //   P1[e, r] = alpha_e (F Wp1^T)[e, r] + beta fe_e u[r],   u[r] = sum_k fc_k Wp1[r, k]
    // where F Wp1^T are Gram columns [PR0h, PR0) of P0 (F is masked, so fe fc F = F).
    PR1 = A.n_proj_rows[1];
    float *p1s = sm + L.h_p1, *u = sm + L.h_u;
    const int Kw = P->Kp, K = d.K;
    const unsigned long long zm = zero_mask_of(flags, N);
    const float *Wp = W + A.proj_w + (size_t)P->PR0h * Kw;
    // u[r] = sum of the projection weights over the live cells: lanes run along the cells (coalesced weight rows, one
    // mask test per cell for 8 projection rows), fixed-order warp tree + sum over warps
    float *part = sm + L.h_part, *alpha_e = sm + L.h_alpha;
#ifdef CCSD_EMU
    const int lane = 0, warp = 0, nwarp = 1;
#else
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#endif
    if (a.g_hu) {
      for (int r = threadIdx.x; r < PR1; r += blockDim.x) u[r] = a.g_hu[(size_t)b * PR1 + r];
      __syncthreads();
    }
    for (int r0 = a.g_hu ? PR1 : 0; r0 < PR1; r0 += 8) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int k = threadIdx.x; k < K; k += blockDim.x) {
        if (!(P->cell_mask[k] & zm)) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (r0 + j < PR1) acc[j] += __ldg(Wp + (size_t)(r0 + j) * Kw + k);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#ifndef CCSD_EMU
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
#endif
        if (lane == 0) part[warp * 8 + j] = acc[j];
      }
      __syncthreads();
      for (int j = threadIdx.x; j < 8 && r0 + j < PR1; j += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nwarp; ++w) t += part[w * 8 + j];
        u[r0 + j] = t;
      }
      __syncthreads();
    }
    const ccsd_mlp_t &mv = h0.mlp_value;
    const float beta = __ldg(W + mv.b[0]);
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const int t = tri_index(P->edge_ij[2 * e], P->edge_ij[2 * e + 1], N);
      float alpha = 0.f;
      for (int c = 0; c < c0; ++c) alpha += __ldg(W + mv.w[0] + c * 8) * stack[c * ldp + t];
      alpha_e[e] = alpha;
    }
    __syncthreads();
    {
      // four Gram columns per thread in flight (the one-at-a-time loop paid an L2 round trip per element: 14 % of the kernel)
      const int tot = E * PR1, PR0h = P->PR0h;
      for (int p0 = threadIdx.x; p0 < tot; p0 += 4 * blockDim.x) {
        float g4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int p = p0 + q * blockDim.x;
          const int e = p / PR1, r = p - e * PR1;
          g4[q] = p < tot ? __ldg(P0 + (size_t)e * PR0 + PR0h + r) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int p = p0 + q * blockDim.x;
          if (p < tot) {
            const int e = p / PR1, r = p - e * PR1;
            const float fe = flags[P->edge_ij[2 * e]] * flags[P->edge_ij[2 * e + 1]];
            p1s[p] = alpha_e[e] * g4[q] + beta * fe * u[r];
          }
        }
      }
    }
    __syncthreads();
    P1 = p1s;
  }
  // layer-0 Q, K for every edge and channel
  for (int p = threadIdx.x; p < c0 * E; p += blockDim.x) {
    const int c = p / E, e = p - c * E;
    const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
    const float av = stack[c * ldp + tri_index(i, j, N)];
    const float dg = 1.0f / sqrtf(fmaxf(av, 1.f));
    const float nrm = dg * av * dg;
    const float *pq = P0 + (size_t)e * PR0 + h0.proj_row + (c * 2 + 0) * ad0;
    const float *pk = pq + ad0;
    for (int dd = 0; dd < ad0; ++dd) {
      hq[(c * E + e) * ad0 + dd] = nrm * pq[dd] + __ldg(W + h0.bq[c] + dd);
      hk[(c * E + e) * ad0 + dd] = nrm * pk[dd] + __ldg(W + h0.bk[c] + dd);
    }
  }
  __syncthreads();
  // layer-0 output  H1[c'][e][e'] = 2 tanh(fe fe' MLP_att(A_.[e,e'])),  A symmetric
  for (int p = threadIdx.x; p < E * (E + 1) / 2; p += blockDim.x) {
    // row-major upper triangle: rows before e hold e E - e (e - 1) / 2 pairs
    int e = (int)(((float)(2 * E + 1) - sqrtf((float)(2 * E + 1) * (float)(2 * E + 1) - 8.f * (float)p)) * 0.5f);
    e = e < 0 ? 0 : (e > E - 1 ? E - 1 : e);
    while (e > 0 && e * E - e * (e - 1) / 2 > p) --e;
    while ((e + 1) * E - (e + 1) * e / 2 <= p) ++e;
    const int e2 = e + (p - (e * E - e * (e - 1) / 2));
    const float fe = flags[P->edge_ij[2 * e]] * flags[P->edge_ij[2 * e + 1]];
    const float fe2 = flags[P->edge_ij[2 * e2]] * flags[P->edge_ij[2 * e2 + 1]];
    float att[CCSD_MAX_CH], out[CCSD_MAX_CH];
#pragma unroll
    for (int c = 0; c < CCSD_MAX_CH; ++c) {
      att[c] = 0.f;
      if (c < c0) {
        const float s1 = hodge_diag_att(hq + (c * E + e) * ad0, hk + (c * E + e2) * ad0, ad0, A.num_heads_h, scale);
        const float s2 = hodge_diag_att(hq + (c * E + e2) * ad0, hk + (c * E + e) * ad0, ad0, A.num_heads_h, scale);
        att[c] = 0.5f * (s1 + s2);
      }
    }
    hodge_channel_mlp(h0.mlp_attention, W, att, out);
#pragma unroll
    for (int c = 0; c < CCSD_MAX_CH; ++c)
      if (c < c1) {
        const float v = 2.0f * fast_tanh(fe * fe2 * out[c]);
        H1[(c * E + e) * lde + e2] = v;
        H1[(c * E + e2) * lde + e] = v;
      }
  }
  __syncthreads();
  // diag of layer-0 output -> stack ; DenseHCNConv degrees of layer 1 (hodge_layers.py:186)
  for (int p = threadIdx.x; p < c1 * E; p += blockDim.x) {
    const int c = p / E, e = p - c * E;
    const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
    const float *row = H1 + (c * E + e) * lde;
    float s = 0.f;
    for (int e2 = 0; e2 < E; ++e2) s += row[e2];
    hdeg[c * E + e] = 1.0f / sqrtf(fmaxf(s, 1.f));
    stack[(ch_hodge0 + c0 + c) * ldp + tri_index(i, j, N)] = row[e];
  }
  __syncthreads();
  // layer 1 (last): only diag(attention) is read back (cc_utils.py:1571).  Item = (edge, channel): aggregate the
  // layer-1 projections with row e of the layer-0 output, then the diagonal attention value
  float *att1 = sm + L.h_att1;
  for (int p = threadIdx.x; p < c1 * E; p += blockDim.x) {
    const int c = p / E, e = p - c * E;
    float q[SMALL_MAX], k[SMALL_MAX];
    for (int dd = 0; dd < ad1; ++dd) { q[dd] = 0.f; k[dd] = 0.f; }
    const float *row = H1 + (c * E + e) * lde;
    const float de = hdeg[c * E + e];
    const float *pq0 = P1 + h1.proj_row + (c * 2 + 0) * ad1;
    if (ad1 == 4) {   // the shipped width: registers instead of the dynamically indexed arrays
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, k0 = 0.f, k1 = 0.f, k2 = 0.f, k3 = 0.f;
      for (int e2 = 0; e2 < E; ++e2) {
        const float w = de * row[e2] * hdeg[c * E + e2];
        const float *pq = pq0 + (size_t)e2 * PR1;
        q0 += w * pq[0]; q1 += w * pq[1]; q2 += w * pq[2]; q3 += w * pq[3];
        k0 += w * pq[4]; k1 += w * pq[5]; k2 += w * pq[6]; k3 += w * pq[7];
      }
      q[0] = q0; q[1] = q1; q[2] = q2; q[3] = q3; k[0] = k0; k[1] = k1; k[2] = k2; k[3] = k3;
    } else {
      for (int e2 = 0; e2 < E; ++e2) {
        const float w = de * row[e2] * hdeg[c * E + e2];
        const float *pq = pq0 + (size_t)e2 * PR1;
        const float *pk = pq + ad1;
        for (int dd = 0; dd < ad1; ++dd) { q[dd] += w * pq[dd]; k[dd] += w * pk[dd]; }
      }
    }
    for (int dd = 0; dd < ad1; ++dd) {
      q[dd] += __ldg(W + h1.bq[c] + dd);
      k[dd] += __ldg(W + h1.bk[c] + dd);
    }
    att1[c * E + e] = hodge_diag_att(q, k, ad1, A.num_heads_h, scale);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
    const float fe = flags[i] * flags[j];
    float att[CCSD_MAX_CH], out[CCSD_MAX_CH];
#pragma unroll
    for (int c = 0; c < CCSD_MAX_CH; ++c) att[c] = c < c1 ? att1[c * E + e] : 0.f;
    hodge_channel_mlp(h1.mlp_attention, W, att, out);
#pragma unroll
    for (int c = 0; c < CCSD_MAX_CH; ++c)
      if (c < h1.c_out) stack[(ch_hodge0 + c0 + c1 + c) * ldp + tri_index(i, j, N)] = 2.0f * fast_tanh(fe * fe * out[c]);
  }
}

// =============================================================================================
// hodge_base_kernel: hodge branch of ScoreNetworkA_Base_CC (ScoreNetwork_A_Base_CC.py:296-312,
// hodge_layers.py:202-416) on the channel stack in global memory.
//
// The hodge-adjacency chain of the baseline network does not depend on rank2 at all: BaselineBlock's
// hodge output is sym(tanh(MLP(hodge_adj))) and only the rank-2 outputs (never read by the adjacency score)
// use rank2.  Layer 0 sees the diagonal Hodge dual diag(a_c), so its block c is
//     M_c[e, e'] = tanh(b2_c[e'] + W2_c[e', :] . u_c[e]),   u_c[e] = elu(W1_c[:, e] a_c[e] + b1_c)
// i.e. `hid` MACs per entry: the E x E layer-0 output is recomputed on the fly (never stored) while the thread
// that owns row e accumulates the first Linear of every layer-1 block over e'.  Only the DIAGONALS of the
// layer outputs reach the final MLP (hodgedual_to_adj, cc_utils.py:1571).
// =============================================================================================
CCSD_KERNEL void __launch_bounds__(128) hodge_base_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_neta_t &A = d.neta;
  const int b = blockIdx.x;
  const int N = d.N, E = d.E, N4 = L.N4, ldp = L.ldp, NT = L.NT;
  const float *W = P->W;
  float *flags = sm + L.h_flags, *u0 = sm + L.hb_u0, *fes = sm + L.hb_fe, *tix = sm + L.hb_tri;
  float *stack = a.g_stack + (size_t)b * L.g_stack;
  const int ch_h0 = a.ch_in;                       // first hodge channel of the stack
  const ccsd_hbase_layer_t &h0 = A.hbase[0], &h1 = A.hbase[1];
  const int c0 = A.c_init, hid0 = h0.hid, hp0 = round_up(hid0, 8), c1 = h0.c_out;
  const int Lh = A.num_layers_h;

  for (int i = threadIdx.x; i < N4; i += blockDim.x) flags[i] = i < N ? a.flags[(size_t)b * N + i] : 0.f;
  // channels [ch_h0, ch_h0 + c0): adjc with zero diagonal; hodge output channels start as zero
  {
    const int nout = h0.c_out + (Lh == 2 ? h1.c_out : 0);
    for (int p = threadIdx.x; p < (c0 + nout) * ldp; p += blockDim.x) {
      const int c = p / ldp, t = p - c * ldp;
      float v = 0.f;
      if (c < c0 && t < NT) {
        const int ij = P->tri_ij[t];
        if ((ij >> 8) != (ij & 255)) v = stack[c * ldp + t];
      }
      stack[(ch_h0 + c) * ldp + t] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
    fes[e] = flags[i] * flags[j];
    reinterpret_cast<int *>(tix)[e] = tri_index(i, j, N);
  }
  __syncthreads();
  // u0[c][e][h] = elu(W1_c[e][h] a_c[e] + b1_c[h])
  for (int p = threadIdx.x; p < c0 * E * hp0; p += blockDim.x) {
    const int c = p / (E * hp0), r = p - c * E * hp0, e = r / hp0, h = r - e * hp0;
    float v = 0.f;
    if (h < hid0) {
      const float ac = stack[c * ldp + reinterpret_cast<const int *>(tix)[e]];
      v = fast_elu(__ldg(W + h0.w1[c] + e * hp0 + h) * ac + __ldg(W + h0.b1[c] + h));
    }
    u0[p] = v;
  }
  __syncthreads();

  const int hid1 = Lh == 2 ? h1.hid : 0;   // <= 8 (validated): hid_pad = 8, so acc[c][0..8) covers a whole padded row
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float fe = fes[e];
    const int te = reinterpret_cast<const int *>(tix)[e];
    float acc[CCSD_MAX_CH][8];   // first Linear of the layer-1 blocks: acc[c][h] = sum_e2 W1'_c[e2][h] O0_c[e, e2]
#pragma unroll
    for (int c = 0; c < CCSD_MAX_CH; ++c)
#pragma unroll
      for (int h = 0; h < 8; ++h) acc[c][h] = 0.f;
    for (int e2 = (Lh == 1 ? e : 0); e2 < (Lh == 1 ? e + 1 : E); ++e2) {   // one hodge layer: only its diagonal
      // layer 0, entry (e, e2): symmetrised block outputs -> mlp_hodge -> mask -> tanh -> + transpose
      float S[CCSD_MAX_CH], Y[SMALL_MAX];
      for (int c = 0; c < c0; ++c) {
        const float *ue = u0 + ((size_t)c * E + e) * hp0, *ue2 = u0 + ((size_t)c * E + e2) * hp0;
        const float *w2a = W + h0.w2[c] + (size_t)e2 * hp0, *w2b = W + h0.w2[c] + (size_t)e * hp0;
        float m1 = __ldg(W + h0.b2[c] + e2), m2 = __ldg(W + h0.b2[c] + e);
        for (int h = 0; h < hid0; ++h) { m1 += __ldg(w2a + h) * ue[h]; m2 += __ldg(w2b + h) * ue2[h]; }
        S[c] = 0.5f * (fast_tanh(m1) + fast_tanh(m2));
      }
      small_mlp(h0.mlp_hodge, W, S, Y, ACT_ELU);
      const float fm = fe * fes[e2];
#pragma unroll
      for (int c = 0; c < CCSD_MAX_CH; ++c)
        if (c < c1) {
          const float o = 2.0f * fast_tanh(fm * Y[c]);
          if (e2 == e) stack[(ch_h0 + c0 + c) * ldp + te] = o;
          if (Lh == 2) {
            const float *w1 = W + h1.w1[c] + (size_t)e2 * 8;
#pragma unroll
            for (int h = 0; h < 8; ++h) acc[c][h] += __ldg(w1 + h) * o;   // padded columns hold zeros
          }
        }
    }
    if (Lh == 2) {
      // layer 1: only the diagonal entry (e, e) of every block and of the layer output is needed
      float Sd[CCSD_MAX_CH], Y[SMALL_MAX];
#pragma unroll
      for (int c = 0; c < CCSD_MAX_CH; ++c)
        if (c < c1) {
          float m = __ldg(W + h1.b2[c] + e);
#pragma unroll
          for (int h = 0; h < 8; ++h)
            if (h < hid1) m += __ldg(W + h1.w2[c] + (size_t)e * 8 + h) * fast_elu(acc[c][h] + __ldg(W + h1.b1[c] + h));
          Sd[c] = fast_tanh(m);
        }
      small_mlp(h1.mlp_hodge, W, Sd, Y, ACT_ELU);
      for (int c = 0; c < h1.c_out; ++c) stack[(ch_h0 + c0 + c1 + c) * ldp + te] = 2.0f * fast_tanh(fe * fe * Y[c]);
    }
  }
}

// =============================================================================================
// afinal_kernel: final per-edge MLP of ScoreNetworkA (ScoreNetwork_A.py:529-539) on a chunk of node
// pairs + the adjacency sampler epilogue
// =============================================================================================
CCSD_KERNEL void __launch_bounds__(128, XP_MINB_M) afinal_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XpLayout &L = P->xp;
  const ccsd_neta_t &A = d.neta;
  const int b = blockIdx.y;
  const int N = d.N, NP = N * N, NT = L.NT, ldp = L.ldp, RC = L.m_rows;
  const int r0 = blockIdx.x * RC;
  const int R = (NT - r0 < RC) ? NT - r0 : RC;
  float *flags = sm + L.m_flags, *fA = sm + L.m_fa, *fB = sm + L.m_fb, *so = sm + L.m_out, *red = sm + L.m_red;
  const float *gs = a.g_stack + (size_t)b * L.g_stack;
  for (int i = threadIdx.x; i < L.N4; i += blockDim.x) flags[i] = i < N ? a.flags[(size_t)b * N + i] : 0.f;
  mlp_fm(A.fin, P->W, gs + r0, ldp, a.ch_out /* = channels in the stack */, nullptr, 0, 0, R, fA, fB, RC, so, 1, 0, ACT_ELU,
         ACT_NONE);
  const size_t ga = (size_t)b * NP;
  const int stp = a.mode == MODE_EVAL ? 0 : nz_step(a.nz);
  const ccsd_objcoef_t ca = a.mode == MODE_EVAL ? ccsd_objcoef_t() : P->sched[stp * 3 + 1];
  const unsigned long long gsid = (unsigned long long)(a.nz.sample_offset + b);
  float s2 = 0.f, z2 = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int ij = P->tri_ij[r0 + r], i = ij >> 8, j = ij & 255;
    // (1 - I) mask and mask_adjs
    const float o = (i == j) ? 0.f : so[r] * flags[i] * flags[j];
    if (a.mode == MODE_EVAL) {
      a.out_adj[ga + i * N + j] = o;
      a.out_adj[ga + j * N + i] = o;
      continue;
    }
    const float s = ca.score_scale * o;
    float z = 0.f;
    if (i != j) {
      const int q = i * N + j;
      z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gsid, draw_id(1, stp, a.slot), q)) * flags[i] * flags[j];
    }
    if (a.mode == MODE_SCORE) {
      a.out_adj[ga + i * N + j] = s;
      if (i != j) {
        a.out_adj[ga + j * N + i] = s;
        s2 += 2.f * s * s;
        z2 += 2.f * z * z;
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
  if (a.mode == MODE_SCORE) {
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(1 * d.B + b) * P->ntile_max + blockIdx.x) * 2;
      np[0] = s2; np[1] = z2;
    }
  }
}

}  // namespace ccsd
