// xa_kernel.cuh -- one CTA per graph: ScoreNetworkX + ScoreNetworkA / ScoreNetworkA_CC entirely in
// shared memory, with the sampler epilogue (score scaling, Langevin norms, predictor update,
// Philox / injected noise, masks) fused in.
//
// Reference: ScoreNetwork_X.py:102-133, ScoreNetwork_A.py:505-541, ScoreNetwork_A_CC.py:275-332,
// attention.py:84-132,270-304, hodge_attention.py:80-129,290-325, layers.py:115-158.
//
// Shape of the kernel: the graphs are tiny (N <= 64), so a CTA is small (64 or 128 threads, chosen
// per plan) and several CTAs share an SM; every phase has on the order of 100 work items (4x8
// register tiles of the feature transforms, 4x4 node blocks of the attention products), and the
// barrier bubbles of one graph are filled by the other graphs resident on the SM.  To make that
// residency possible every adjacency-shaped tensor lives in upper-triangle storage (N(N+1)/2 node
// pairs instead of N^2): the adjacency state is symmetric, and so is every channel derived from it
// (matrix powers, symmetrised attention, M + M^T) -- which also halves the per-edge MLP work.
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

namespace ccsd {

struct XaArgs {
  const float *x, *adj, *flags;  // [B,N,F] [B,N,N] [B,N]
  const float *P0, *P1;          // hodge projections [B,E,PR0] [B,E,PR1] (CC only)
  int mode;                      // MODE_*
  int which;                     // bit0: evaluate X net, bit1: evaluate A net
  float *out_x, *out_adj;        // EVAL: raw net output; SCORE: scaled score; PRED: new state
  float *mean_x, *mean_adj;      // PRED: means
  float *norm_part;              // SCORE: [3][B][ntile_max][2]
  const float *noise_x, *noise_adj;  // raw normals for this draw ([B,...]) or nullptr (Philox)
  float *traj_x, *traj_adj;      // PRED: destination for sample 0 of this shard (or nullptr)
  int slot;                      // draw slot within the step
  int denoise;
  NoiseCtx nz;
};

// ---- hodge branch -----------------------------------------------------------------------------
__device__ __forceinline__ float hodge_diag_att(const float *q, const float *k, int ad, int heads, float scale) {
  const int ds = ad / heads;
  const int nch = (ad + ds - 1) / ds;
  float s = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int d0 = c * ds, d1 = (d0 + ds < ad) ? d0 + ds : ad;
    float a = 0.f;
    for (int dd = d0; dd < d1; ++dd) a += q[dd] * k[dd];
    s += tanhf(a * scale);
  }
  return s / (float)nch;
}

__device__ void hodge_branch(const DevPlan *__restrict__ P, const XaArgs &a, float *sm, int b, int ch_hodge0) {
  const ccsd_plan_desc_t &d = P->d;
  const XaLayout &L = P->xa;
  const ccsd_neta_t &A = d.neta;
  const int N = d.N, E = d.E, ldp = L.ldp, NT = L.NT;
  const float *W = P->W;
  float *stack = sm + L.stack;
  const float *flags = sm + L.flags;
  const float scale = 1.0f / sqrtf((float)d.K);  // HodgeAttention out_dim = K (hodge_attention.py:236-239)
  const int c0 = A.c_init;
  const ccsd_hodge_layer_t &h0 = A.hodge[0];
  const int ad0 = h0.attn_dim;
  const int PR0 = P->PR0;
  const float *P0 = a.P0 + (size_t)b * E * PR0;
  const int *pij = reinterpret_cast<const int *>(sm + L.pij);

  // channels [ch_hodge0, ch_hodge0 + c0): hodgedual_to_adj(adj_to_hodgedual(adjc)) = adjc with zero diagonal;
  // the hodge output channels start as zero (diagonals stay zero, off-diagonals are filled per edge)
  {
    const int nout = h0.c_out + (A.num_layers_h == 2 ? A.hodge[1].c_out : 0);
    for (int p = threadIdx.x; p < (c0 + nout) * NT; p += blockDim.x) {
      const int c = p / NT, t = p - c * NT;
      const int ij = pij[t];
      float v = 0.f;
      if (c < c0 && (ij >> 8) != (ij & 255)) v = stack[c * ldp + t];
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
    __syncthreads();
    return;
  }

  // ---- two hodge layers ----
  const ccsd_hodge_layer_t &h1 = A.hodge[1];
  const int c1 = h0.c_out, ad1 = h1.attn_dim, PR1 = P->PR1, lde = L.lde;
  float *hq = sm + L.scratch + L.hq, *hk = sm + L.scratch + L.hk;
  float *H1 = sm + L.scratch + L.h1, *hdeg = sm + L.scratch + L.hdeg;
  const float *P1 = a.P1 + (size_t)b * E * PR1;
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
    int e = 0, rem = p;
    while (rem >= E - e) { rem -= E - e; ++e; }
    const int e2 = e + rem;
    const float fe = flags[P->edge_ij[2 * e]] * flags[P->edge_ij[2 * e + 1]];
    const float fe2 = flags[P->edge_ij[2 * e2]] * flags[P->edge_ij[2 * e2 + 1]];
    float att[CCSD_MAX_CH], out[SMALL_MAX];
    for (int c = 0; c < c0; ++c) {
      const float s1 = hodge_diag_att(hq + (c * E + e) * ad0, hk + (c * E + e2) * ad0, ad0, A.num_heads_h, scale);
      const float s2 = hodge_diag_att(hq + (c * E + e2) * ad0, hk + (c * E + e) * ad0, ad0, A.num_heads_h, scale);
      att[c] = 0.5f * (s1 + s2);
    }
    small_mlp(h0.mlp_attention, W, att, out, ACT_ELU);
    for (int c = 0; c < c1; ++c) {
      const float v = 2.0f * tanhf(fe * fe2 * out[c]);
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
  // layer 1 (last): only diag(attention) is read back (cc_utils.py:1571)
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
    const float fe = flags[i] * flags[j];
    float att[CCSD_MAX_CH], q[SMALL_MAX], k[SMALL_MAX], out[SMALL_MAX];
    for (int c = 0; c < c1; ++c) {
      for (int dd = 0; dd < ad1; ++dd) { q[dd] = 0.f; k[dd] = 0.f; }
      const float *row = H1 + (c * E + e) * lde;
      const float de = hdeg[c * E + e];
      for (int e2 = 0; e2 < E; ++e2) {
        const float w = de * row[e2] * hdeg[c * E + e2];
        const float *pq = P1 + (size_t)e2 * PR1 + h1.proj_row + (c * 2 + 0) * ad1;
        const float *pk = pq + ad1;
        for (int dd = 0; dd < ad1; ++dd) { q[dd] += w * pq[dd]; k[dd] += w * pk[dd]; }
      }
      for (int dd = 0; dd < ad1; ++dd) {
        q[dd] += __ldg(W + h1.bq[c] + dd);
        k[dd] += __ldg(W + h1.bk[c] + dd);
      }
      att[c] = hodge_diag_att(q, k, ad1, A.num_heads_h, scale);
    }
    small_mlp(h1.mlp_attention, W, att, out, ACT_ELU);
    for (int c = 0; c < h1.c_out; ++c)
      stack[(ch_hodge0 + c0 + c1 + c) * ldp + tri_index(i, j, N)] = 2.0f * tanhf(fe * fe * out[c]);
  }
  __syncthreads();
}

// ---- the kernel -------------------------------------------------------------------------------
__global__ void __launch_bounds__(XA_MAX_THREADS, XA_MIN_BLOCKS) xa_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XaLayout &L = P->xa;
  const int b = blockIdx.x;
  const int N = d.N, F = d.F, NP = N * N, N4 = L.N4, NT = L.NT, ldp = L.ldp;
  const float *W = P->W;
  float *flags = sm + L.flags, *dvec = sm + L.dvec, *an = sm + L.an, *x0 = sm + L.x0;
  float *stack = sm + L.stack, *sx = sm + L.sx, *sadj = sm + L.sadj, *red = sm + L.red;
  float *scr = sm + L.scratch;
  int *pij = reinterpret_cast<int *>(sm + L.pij);

  // ---- load the graph tile ----
  for (int i = threadIdx.x; i < N4; i += blockDim.x) flags[i] = i < N ? a.flags[(size_t)b * N + i] : 0.f;
  for (int p = threadIdx.x; p < F * N4; p += blockDim.x) {
    const int f = p / N4, i = p - f * N4;
    x0[p] = i < N ? a.x[((size_t)b * N + i) * F + f] : 0.f;
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int t0 = tri_index(i, i, N);
    for (int j = i; j < N; ++j) {
      pij[t0 + j - i] = (i << 8) | j;
      stack[t0 + j - i] = a.adj[(size_t)b * NP + i * N + j];
    }
  }
  __syncthreads();

  // ================= ScoreNetworkX =================
  if (a.which & 1) {
    const ccsd_netx_t &X = d.netx;
    float *hcat = sm + L.xh_cat, *ax = sm + L.xh_ax, *hA = sm + L.xh_a, *hB = sm + L.xh_b;
    gcn_norm_tri(stack, N, N4, dvec, an);
    const float *in = x0;
    int din = F, row = 0;
    for (int k = 0; k < X.depth; ++k) {
      const ccsd_gcn_t &g = X.gcn[k];
      gcn_aggregate_fm(an, N, N4, in, din, ax);
      __syncthreads();
      dense_fm(ax, N4, din, nullptr, 0, 0, W + g.w, W + g.b, g.dout, hcat + row * N4, 1, N4, N, ACT_TANH);
      __syncthreads();
      in = hcat + row * N4;
      din = g.dout;
      row += g.dout;
    }
    mlp_fm(X.fin, W, x0, N4, F, hcat, N4, row, N, hA, hB, N4, sx, 1, N4, ACT_ELU, ACT_NONE);
    for (int p = threadIdx.x; p < F * N4; p += blockDim.x) {
      const int i = p % N4;
      if (i < N) sx[p] *= flags[i];
    }
    __syncthreads();
  }

  // ================= ScoreNetworkA / ScoreNetworkA_CC =================
  if (a.which & 2) {
    const ccsd_neta_t &A = d.neta;
    float *att = scr + L.att, *ax = scr + L.ax, *hmc = scr + L.hmc, *hmc2 = scr + L.hmc2;
    float *q = scr + L.q, *kf = scr + L.k, *v = scr + L.v, *atp = scr + L.atp;
    float *ehA = scr + L.eh_a, *ehB = scr + L.eh_b;
    // pow_tensor (graph_utils.py:274-292): A^c = A^(c-1) . A, symmetric
    for (int c = 1; c < A.c_init; ++c) {
      for (int t = threadIdx.x; t < NT; t += blockDim.x) {
        const int ij = pij[t], i = ij >> 8, j = ij & 255;
        float s = 0.f;
        for (int k = 0; k < N; ++k) s += stack[(c - 1) * ldp + tri_index_any(i, k, N)] * stack[tri_index_any(k, j, N)];
        stack[c * ldp + t] = s;
      }
      __syncthreads();
    }
    const float *xin = x0;  // layer 0 reads the raw node features
    int kin = F;
    float *xnext = sm + L.xa, *xother = sm + L.xb;
    int ch_in = 0, ch_out = A.c_init;
    for (int l = 0; l < A.num_layers; ++l) {
      const ccsd_attn_layer_t &ly = A.layer[l];
      const int ad = ly.attn_dim, nh = ly.conv_out;
      const float scale = 1.0f / sqrtf((float)nh);  // / sqrt(out_dim)  (attention.py:125)
      const ccsd_mlp_t &mc = ly.multi_channel;
      const int mc_o1 = mc.nl == 1 ? mc.dout : mc.dhid, mc_o1p = round_up(mc_o1, 8);
      const int nch = (ad + ad / A.num_heads - 1) / (ad / A.num_heads);
      for (int c = 0; c < ly.c_in; ++c) {
        gcn_norm_tri(stack + (ch_in + c) * ldp, N, N4, dvec, an);
        gcn_aggregate_fm(an, N, N4, xin, kin, ax);
        __syncthreads();
        {
          // Q | K | V = A x W_{q,k,v} + b as one item space (same input tile, three weight matrices)
          const int ngrp = N4 >> 2;
          const int nq = (round_up(ad, 8) >> 3) * ngrp, nv = (round_up(nh, 8) >> 3) * ngrp;
          for (int it = threadIdx.x; it < 2 * nq + nv; it += blockDim.x) {
            const int w = it < nq ? 0 : (it < 2 * nq ? 1 : 2);
            const int li = it - (w == 0 ? 0 : (w == 1 ? nq : 2 * nq));
            const ccsd_gcn_t &g = w == 0 ? ly.q[c] : (w == 1 ? ly.k[c] : ly.v[c]);
            float *dst = w == 0 ? q : (w == 1 ? kf : v);
            const int O = g.dout, Opad = round_up(O, 8);
            const int chunk = li / ngrp, r0 = (li - chunk * ngrp) << 2, oc = chunk << 3;
            float acc[4][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float bv = __ldg(W + g.b + oc + j);
#pragma unroll
              for (int rr = 0; rr < 4; ++rr) acc[rr][j] = bv;
            }
            dense_tile(acc, ax, N4, kin, nullptr, 0, 0, W + g.w, Opad, r0, oc);
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
        attn_scores_blk(q, kf, N, N4, ad, A.num_heads, scale, atp, ldp);
        // node branch, first Linear of multi_channel folded over the channel concat (attention.py:292):
        // hmc(o, i) += sum_f V_c(i, f) W1[c*nh + f, o]
        dense_fm(v, N4, nh, nullptr, 0, 0, W + mc.w[0] + (size_t)c * nh * mc_o1p, nullptr, mc_o1, hmc, 1, N4, N, ACT_NONE,
                 c > 0, /*first thread*/ (int)blockDim.x - 32);
        __syncthreads();
        for (int t = threadIdx.x; t < NT; t += blockDim.x) {
          float s = 0.f;
          for (int h = 0; h < nch; ++h) s += atp[h * ldp + t];
          att[c * ldp + t] = s;
        }
      }
      // node branch: x_out = tanh(mask_x(MLP(cat V)))  (attention.py:292-293)
      {
        float *cur = hmc, *oth = hmc2;
        for (int p = threadIdx.x; p < mc_o1 * N4; p += blockDim.x) {
          const int o = p / N4;
          const float t = cur[p] + __ldg(W + mc.b[0] + o);
          cur[p] = mc.nl == 1 ? t : fast_elu(t);
        }
        __syncthreads();
        for (int li = 1; li < mc.nl; ++li) {
          const bool last = li == mc.nl - 1;
          const int O = last ? mc.dout : mc.dhid;
          dense_fm(cur, N4, li == 1 ? mc_o1 : mc.dhid, nullptr, 0, 0, W + mc.w[li], W + mc.b[li], O, oth, 1, N4, N,
                   last ? ACT_NONE : ACT_ELU);
          __syncthreads();
          float *t = cur; cur = oth; oth = t;
        }
        for (int p = threadIdx.x; p < nh * N4; p += blockDim.x) {
          const int i = p % N4;
          xnext[p] = i < N ? fast_tanh(cur[p] * flags[i]) : 0.f;
        }
      }
      // edge branch: M = MLP(cat[A_1..A_c, adj_1..adj_c]) ; adj_out = mask_adjs(M + M^T)  (attention.py:295-302)
      // (the head sums of the last channel are ordered before this by the __syncthreads above)
      mlp_fm(ly.mlp, W, att, ldp, ly.c_in, stack + ch_in * ldp, ldp, ly.c_in, NT, ehA, ehB, ldp, stack + ch_out * ldp, 1,
             ldp, ACT_ELU, ACT_NONE);
      for (int p = threadIdx.x; p < ly.c_out * NT; p += blockDim.x) {
        const int c = p / NT, t = p - c * NT;
        const int ij = pij[t];
        float *pl = stack + (ch_out + c) * ldp + t;
        *pl = 2.0f * *pl * flags[ij >> 8] * flags[ij & 255];
      }
      __syncthreads();
      ch_in = ch_out;
      ch_out += ly.c_out;
      xin = xnext;
      kin = nh;
      float *t = xnext; xnext = xother; xother = t;
    }
    int fd_have = ch_out;
    if (A.is_cc) {
      hodge_branch(P, a, sm, b, ch_out);
      fd_have += A.c_init + A.hodge[0].c_out + (A.num_layers_h == 2 ? A.hodge[1].c_out : 0);
    }
    // final per-edge MLP, (1 - I) mask, mask_adjs  (ScoreNetwork_A.py:529-539)
    {
      float *fA = scr + L.fh_a, *fB = scr + L.fh_b;
      const int RC = L.fin_rows;
      for (int r0 = 0; r0 < NT; r0 += RC) {
        const int R = (NT - r0 < RC) ? NT - r0 : RC;
        mlp_fm(A.fin, W, stack + r0, ldp, fd_have, nullptr, 0, 0, R, fA, fB, RC, sadj + r0, 1, 0, ACT_ELU, ACT_NONE);
      }
    }
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
      const int ij = pij[t], i = ij >> 8, j = ij & 255;
      sadj[t] = (i == j) ? 0.f : sadj[t] * flags[i] * flags[j];
    }
    __syncthreads();
  }

  // ================= epilogue =================
  const size_t gx = (size_t)b * N * F, ga = (size_t)b * NP;
  if (a.mode == MODE_EVAL) {
    if (a.which & 1)
      for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
        const int i = p / F, f = p - i * F;
        a.out_x[gx + p] = sx[f * N4 + i];
      }
    if (a.which & 2)
      for (int p = threadIdx.x; p < NP; p += blockDim.x) {
        const int i = p / N, j = p - i * N;
        a.out_adj[ga + p] = sadj[tri_index_any(i, j, N)];
      }
    return;
  }
  const ccsd_objcoef_t cx = P->sched[a.nz.step * 3 + 0], ca = P->sched[a.nz.step * 3 + 1];
  const unsigned long long gs = (unsigned long long)(a.nz.sample_offset + b);
  if (a.mode == MODE_SCORE) {
    // scaled scores + per-sample squared norms of score and (masked) noise (solver.py:693-699, 1299-1305)
    float s2 = 0.f, z2 = 0.f;
    for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
      const int i = p / F, f = p - i * F;
      const float s = cx.score_scale * sx[f * N4 + i];
      a.out_x[gx + p] = s;
      const float z = (a.noise_x ? a.noise_x[gx + p] : normal1(a.nz.seed, gs, draw_id(0, a.nz.step, a.slot), p)) *
                      flags[i];
      s2 += s * s;
      z2 += z * z;
    }
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(0 * d.B + b) * P->ntile_max) * 2;
      np[0] = s2; np[1] = z2;
    }
    s2 = 0.f; z2 = 0.f;
    for (int t = threadIdx.x; t < NT; t += blockDim.x) {
      const int ij = pij[t], i = ij >> 8, j = ij & 255;
      const float s = ca.score_scale * sadj[t];
      a.out_adj[ga + i * N + j] = s;
      if (i != j) {
        a.out_adj[ga + j * N + i] = s;
        const int q = i * N + j;
        const float z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gs, draw_id(1, a.nz.step, a.slot), q)) *
                        flags[i] * flags[j];
        s2 += 2.f * s * s;
        z2 += 2.f * z * z;
      }
    }
    s2 = block_sum(s2, red);
    z2 = block_sum(z2, red);
    if (threadIdx.x == 0) {
      float *np = a.norm_part + ((size_t)(1 * d.B + b) * P->ntile_max) * 2;
      np[0] = s2; np[1] = z2;
    }
    return;
  }
  // MODE_PRED: mean = pa*obj + pb*score ; new = mean + pc*z   (solver.py:230-244, 386-398; sde.py:200-235)
  for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
    const int i = p / F, f = p - i * F;
    const float s = cx.score_scale * sx[f * N4 + i];
    const float z = (a.noise_x ? a.noise_x[gx + p] : normal1(a.nz.seed, gs, draw_id(0, a.nz.step, a.slot), p)) *
                    flags[i];
    const float m = cx.pa * x0[f * N4 + i] + cx.pb * s;
    const float v = m + cx.pc * z;
    a.out_x[gx + p] = v;
    a.mean_x[gx + p] = m;
    if (a.traj_x && b == 0) a.traj_x[p] = a.denoise ? m : v;
  }
  for (int t = threadIdx.x; t < NT; t += blockDim.x) {
    const int ij = pij[t], i = ij >> 8, j = ij & 255;
    const float s = ca.score_scale * sadj[t];
    float z = 0.f;
    if (i != j) {
      const int q = i * N + j;
      z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gs, draw_id(1, a.nz.step, a.slot), q)) * flags[i] *
          flags[j];
    }
    const float m = ca.pa * stack[t] + ca.pb * s;
    const float v = m + ca.pc * z;
    a.out_adj[ga + i * N + j] = v;
    a.mean_adj[ga + i * N + j] = m;
    if (a.traj_adj && b == 0) a.traj_adj[i * N + j] = a.denoise ? m : v;
    if (i != j) {
      a.out_adj[ga + j * N + i] = v;
      a.mean_adj[ga + j * N + i] = m;
      if (a.traj_adj && b == 0) a.traj_adj[j * N + i] = a.denoise ? m : v;
    }
  }
}

}  // namespace ccsd
