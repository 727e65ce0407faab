// xa_kernel.cuh -- one CTA per graph: ScoreNetworkX + ScoreNetworkA / ScoreNetworkA_CC entirely in
// shared memory, with the sampler epilogue (score scaling, Langevin norms, predictor update,
// Philox / injected noise, masks) fused in.
//
// Reference: ScoreNetwork_X.py:102-133, ScoreNetwork_A.py:505-541, ScoreNetwork_A_CC.py:275-332,
// attention.py:84-132,270-304, hodge_attention.py:80-129,290-325, layers.py:115-158.
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

__device__ __forceinline__ int edge_index(int i, int j, int N) {  // i < j
  return i * N - (i * (i + 1)) / 2 + (j - i - 1);
}

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
  const int N = d.N, E = d.E, ldp = L.ldp;
  const float *W = P->W;
  float *stack = sm + L.stack;
  const float *flags = sm + L.flags;
  const float scale = 1.0f / sqrtf((float)d.K);  // HodgeAttention out_dim = K (hodge_attention.py:236-239)
  const int c0 = A.c_init;
  const ccsd_hodge_layer_t &h0 = A.hodge[0];
  const int ad0 = h0.attn_dim;
  const int PR0 = P->PR0;
  const float *P0 = a.P0 + (size_t)b * E * PR0;

  // channels [ch_hodge0, ch_hodge0 + c0): hodgedual_to_adj(adj_to_hodgedual(adjc)) = adjc with zero diagonal
  for (int p = threadIdx.x; p < c0 * N * N; p += blockDim.x) {
    const int c = p / (N * N), ij = p - c * N * N;
    const int i = ij / N, j = ij - i * N;
    stack[(ch_hodge0 + c) * ldp + ij] = (i == j) ? 0.f : stack[c * ldp + ij];
  }
  // zero the diagonals / whole planes of the hodge output channels (off-diagonals are filled per edge)
  {
    int nout = h0.c_out + (A.num_layers_h == 2 ? A.hodge[1].c_out : 0);
    for (int p = threadIdx.x; p < nout * N; p += blockDim.x) {
      const int c = p / N, i = p - c * N;
      stack[(ch_hodge0 + c0 + c) * ldp + i * N + i] = 0.f;
    }
  }
  __syncthreads();

  if (A.num_layers_h == 1) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const int i = P->edge_ij[2 * e], j = P->edge_ij[2 * e + 1];
      const float fe = flags[i] * flags[j];
      float att[CCSD_MAX_CH], q[SMALL_MAX], k[SMALL_MAX], out[SMALL_MAX];
      for (int c = 0; c < c0; ++c) {
        const float av = stack[c * ldp + i * N + j];
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
      for (int c = 0; c < h0.c_out; ++c) {
        const float v = 2.0f * tanhf(fe * fe * out[c]);
        stack[(ch_hodge0 + c0 + c) * ldp + i * N + j] = v;
        stack[(ch_hodge0 + c0 + c) * ldp + j * N + i] = v;
      }
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
    const float av = stack[c * ldp + i * N + j];
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
    const float v = row[e];
    stack[(ch_hodge0 + c0 + c) * ldp + i * N + j] = v;
    stack[(ch_hodge0 + c0 + c) * ldp + j * N + i] = v;
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
    for (int c = 0; c < h1.c_out; ++c) {
      const float v = 2.0f * tanhf(fe * fe * out[c]);
      stack[(ch_hodge0 + c0 + c1 + c) * ldp + i * N + j] = v;
      stack[(ch_hodge0 + c0 + c1 + c) * ldp + j * N + i] = v;
    }
  }
  __syncthreads();
}

// ---- the kernel -------------------------------------------------------------------------------
__global__ void __launch_bounds__(XA_THREADS, 1) xa_kernel(const DevPlan *__restrict__ P, XaArgs a) {
  CCSD_SMEM(sm);
  const ccsd_plan_desc_t &d = P->d;
  const XaLayout &L = P->xa;
  const int b = blockIdx.x;
  const int N = d.N, F = d.F, NP = N * N, ldn = L.ldn, ldp = L.ldp;
  const float *W = P->W;
  float *flags = sm + L.flags, *dvec = sm + L.dvec, *hcat = sm + L.hcat, *an = sm + L.an;
  float *stack = sm + L.stack, *sx = sm + L.sx, *sadj = sm + L.sadj, *red = sm + L.red;
  float *scr = sm + L.scratch;
  float *xw = scr + L.xw, *hA = scr + L.hA, *hB = scr + L.hB;
  const int ldxw = L.ldxw;

  // ---- load the graph tile ----
  for (int i = threadIdx.x; i < N; i += blockDim.x) flags[i] = a.flags[(size_t)b * N + i];
  for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
    const int i = p / F, f = p - i * F;
    hcat[f * ldn + i] = a.x[(size_t)b * N * F + p];
  }
  for (int p = threadIdx.x; p < NP; p += blockDim.x) stack[p] = a.adj[(size_t)b * NP + p];
  __syncthreads();

  // ================= ScoreNetworkX =================
  if (a.which & 1) {
    const ccsd_netx_t &X = d.netx;
    gcn_norm(stack, N, dvec, an, ldn);
    int row_in = 0, din = F;
    for (int k = 0; k < X.depth; ++k) {
      const ccsd_gcn_t &g = X.gcn[k];
      dense2(hcat + row_in * ldn, ldn, din, nullptr, 0, 0, W + g.w, nullptr, g.dout, xw, ldxw, 1, N, ACT_NONE);
      __syncthreads();
      const int row_out = row_in + din;
      gcn_aggregate(an, ldn, N, xw, ldxw, 0, g.dout, W + g.b, hcat + row_out * ldn, 1, ldn, ACT_TANH);
      __syncthreads();
      row_in = row_out;
      din = g.dout;
    }
    mlp_rows(X.fin, W, hcat, ldn, X.fdim, nullptr, 0, 0, N, hA, hB, ldn, sx, 1, ldn, ACT_ELU, ACT_NONE);
    for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
      const int f = p / N, i = p - f * N;
      sx[f * ldn + i] *= flags[i];
    }
    __syncthreads();
  }

  // ================= ScoreNetworkA / ScoreNetworkA_CC =================
  if (a.which & 2) {
    const ccsd_neta_t &A = d.neta;
    float *qn = scr + L.qn, *kf = scr + L.kf, *vcat = scr + L.vcat, *att = scr + L.att;
    const int ldq = L.ldq;
    // pow_tensor (graph_utils.py:274-292)
    for (int c = 1; c < A.c_init; ++c) {
      for (int p = threadIdx.x; p < NP; p += blockDim.x) {
        const int i = p / N, j = p - i * N;
        float s = 0.f;
        for (int k = 0; k < N; ++k) s += stack[(c - 1) * ldp + i * N + k] * stack[k * N + j];
        stack[c * ldp + p] = s;
      }
      __syncthreads();
    }
    const float *xin = hcat;  // layer 0 reads the raw node features (feature-major, rows 0..F-1)
    int kin = F;
    float *xcur = sm + L.xa, *xnext = sm + L.xb;
    int ch_in = 0, ch_out = A.c_init;
    for (int l = 0; l < A.num_layers; ++l) {
      const ccsd_attn_layer_t &ly = A.layer[l];
      const int ad = ly.attn_dim, nh = ly.conv_out;
      const int adp = round_up(ad, 4);
      const float scale = 1.0f / sqrtf((float)nh);  // / sqrt(out_dim)  (attention.py:125)
      for (int c = 0; c < ly.c_in; ++c) {
        gcn_norm(stack + (ch_in + c) * ldp, N, dvec, an, ldn);
        dense2(xin, ldn, kin, nullptr, 0, 0, W + ly.q[c].w, nullptr, ad, xw, ldxw, 1, N, ACT_NONE);
        dense2(xin, ldn, kin, nullptr, 0, 0, W + ly.k[c].w, nullptr, ad, xw + adp, ldxw, 1, N, ACT_NONE);
        dense2(xin, ldn, kin, nullptr, 0, 0, W + ly.v[c].w, nullptr, nh, xw + 2 * adp, ldxw, 1, N, ACT_NONE);
        __syncthreads();
        gcn_aggregate(an, ldn, N, xw, ldxw, 0, ad, W + ly.q[c].b, qn, ldq, 1, ACT_NONE);
        gcn_aggregate(an, ldn, N, xw, ldxw, adp, ad, W + ly.k[c].b, kf, 1, ldn, ACT_NONE);
        gcn_aggregate(an, ldn, N, xw, ldxw, 2 * adp, nh, W + ly.v[c].b, vcat + c * nh * ldn, 1, ldn, ACT_NONE);
        __syncthreads();
        attn_scores(qn, ldq, kf, ldn, N, ad, A.num_heads, scale, att + c * ldp);
        __syncthreads();
      }
      // node branch: x_out = tanh(mask_x(MLP(cat V)))  (attention.py:292-293)
      mlp_rows(ly.multi_channel, W, vcat, ldn, ly.c_in * nh, nullptr, 0, 0, N, hA, hB, ldn, xnext, 1, ldn,
               ACT_ELU, ACT_NONE);
      for (int p = threadIdx.x; p < nh * N; p += blockDim.x) {
        const int f = p / N, i = p - f * N;
        xnext[f * ldn + i] = tanhf(xnext[f * ldn + i] * flags[i]);
      }
      // edge branch: M = MLP(cat[A_1..A_c, adj_1..adj_c]) ; adj_out = mask_adjs(M + M^T)  (attention.py:295-302)
      mlp_rows(ly.mlp, W, att, ldp, ly.c_in, stack + ch_in * ldp, ldp, ly.c_in, NP, hA, hB, ldp,
               stack + ch_out * ldp, 1, ldp, ACT_ELU, ACT_NONE);
      for (int p = threadIdx.x; p < ly.c_out * (N * (N + 1) / 2); p += blockDim.x) {
        const int c = p / (N * (N + 1) / 2);
        int rem = p - c * (N * (N + 1) / 2), i = 0;
        while (rem >= N - i) { rem -= N - i; ++i; }
        const int j = i + rem;
        float *pl = stack + (ch_out + c) * ldp;
        const float v = (pl[i * N + j] + pl[j * N + i]) * flags[i] * flags[j];
        pl[i * N + j] = v;
        pl[j * N + i] = v;
      }
      __syncthreads();
      ch_in = ch_out;
      ch_out += ly.c_out;
      xin = xnext;
      kin = nh;
      float *t = xcur; xcur = xnext; xnext = t;
    }
    int fd_have = ch_out;
    if (A.is_cc) {
      hodge_branch(P, a, sm, b, ch_out);
      fd_have += A.c_init + A.hodge[0].c_out + (A.num_layers_h == 2 ? A.hodge[1].c_out : 0);
    }
    // final per-edge MLP, (1 - I) mask, mask_adjs  (ScoreNetwork_A.py:529-539)
    const int RC = L.fin_rows;
    for (int r0 = 0; r0 < NP; r0 += RC) {
      const int R = (NP - r0 < RC) ? NP - r0 : RC;
      mlp_rows(A.fin, W, stack + r0, ldp, fd_have, nullptr, 0, 0, R, hA, hB, L.fin_ld, sadj + r0, 1, 0, ACT_ELU,
               ACT_NONE);
    }
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
      const int i = p / N, j = p - i * N;
      sadj[p] = (i == j) ? 0.f : sadj[p] * flags[i] * flags[j];
    }
    __syncthreads();
  }

  // ================= epilogue =================
  const size_t gx = (size_t)b * N * F, ga = (size_t)b * NP;
  if (a.mode == MODE_EVAL) {
    if (a.which & 1)
      for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
        const int i = p / F, f = p - i * F;
        a.out_x[gx + p] = sx[f * ldn + i];
      }
    if (a.which & 2)
      for (int p = threadIdx.x; p < NP; p += blockDim.x) a.out_adj[ga + p] = sadj[p];
    return;
  }
  const ccsd_objcoef_t cx = P->sched[a.nz.step * 3 + 0], ca = P->sched[a.nz.step * 3 + 1];
  const unsigned long long gs = (unsigned long long)(a.nz.sample_offset + b);
  if (a.mode == MODE_SCORE) {
    // scaled scores + per-sample squared norms of score and (masked) noise (solver.py:693-699, 1299-1305)
    float s2 = 0.f, z2 = 0.f;
    for (int p = threadIdx.x; p < N * F; p += blockDim.x) {
      const int i = p / F, f = p - i * F;
      const float s = cx.score_scale * sx[f * ldn + i];
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
    for (int p = threadIdx.x; p < NP; p += blockDim.x) {
      const int i = p / N, j = p - i * N;
      const float s = ca.score_scale * sadj[p];
      a.out_adj[ga + p] = s;
      float z = 0.f;
      if (i != j) {
        const int q = (i < j) ? i * N + j : j * N + i;
        z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gs, draw_id(1, a.nz.step, a.slot), q)) *
            flags[i] * flags[j];
      }
      s2 += s * s;
      z2 += z * z;
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
    const float s = cx.score_scale * sx[f * ldn + i];
    const float z = (a.noise_x ? a.noise_x[gx + p] : normal1(a.nz.seed, gs, draw_id(0, a.nz.step, a.slot), p)) *
                    flags[i];
    const float m = cx.pa * hcat[f * ldn + i] + cx.pb * s;
    const float v = m + cx.pc * z;
    a.out_x[gx + p] = v;
    a.mean_x[gx + p] = m;
    if (a.traj_x && b == 0) a.traj_x[p] = a.denoise ? m : v;
  }
  for (int p = threadIdx.x; p < NP; p += blockDim.x) {
    const int i = p / N, j = p - i * N;
    const float s = ca.score_scale * sadj[p];
    float z = 0.f;
    if (i != j) {
      const int q = (i < j) ? i * N + j : j * N + i;
      z = (a.noise_adj ? a.noise_adj[ga + q] : normal1(a.nz.seed, gs, draw_id(1, a.nz.step, a.slot), q)) *
          flags[i] * flags[j];
    }
    const float m = ca.pa * stack[p] + ca.pb * s;
    const float v = m + ca.pc * z;
    a.out_adj[ga + p] = v;
    a.mean_adj[ga + p] = m;
    if (a.traj_adj && b == 0) a.traj_adj[p] = a.denoise ? m : v;
  }
}

}  // namespace ccsd
