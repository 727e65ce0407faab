// ccsd_b200.cu -- host side of the C ABI (include/ccsd_b200.h): plan validation, shared-memory and
// workspace layout, the per-step launch sequences of the PC and S4 samplers.
//
// Reference call structure being replaced: ccsd/src/solver.py:961-1003 / 1109-1174 (pc_sampler),
// :1267-1371 / 1409-1561 (s4_solver).  Launch sequence per PC step with the Langevin corrector:
//   corrector:  gram(F0) [proj1] xa(SCORE) apply(NORM) coef update(x, adj) apply(CORR)   -> (x1, adj1, F1)
//   predictor:  gram(F1) [proj1] xa(PRED)  apply(PRED)                                  -> (x2, adj2, F2)
// and per S4 step:  gram [proj1] xa(SCORE) apply(SCORE) coef update(S4 chain).
// "xa" is the x / adj network pipeline of xa_pipe.cuh (x_net, attn_channel + attn_finish per attention
// layer, hodge, afinal).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "r2_kernels.cuh"
#include "xa_pipe.cuh"
#include "big_pipe.cuh"
#ifndef CCSD_EMU
#include "tc_gram.cuh"
#include "tc_apply.cuh"
#include "tc_afinal.cuh"
#include "tc_agg.cuh"
#include "tc_attn.cuh"
#include "tc_xfin.cuh"
#include "tc_hnorm.cuh"
#include "tc_edge.cuh"
#include "tc_r2big.cuh"
#endif

#ifdef CCSD_EMU
thread_local ccsd_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local float *ccsd_emu_smem = nullptr;
#endif

using namespace ccsd;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#ifdef CCSD_EMU
static int dev_copy(void *dst, const void *src, size_t bytes, void *) {
  memcpy(dst, src, bytes);
  return 0;
}
static int dev_check(const char *) { return 0; }
#else
static int dev_copy(void *dst, const void *src, size_t bytes, void *stream) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(CCSD_ERR_CUDA, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(e));
  return 0;
}
static int dev_check(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CCSD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}
#endif

struct ccsd_plan {
  DevPlan hp;  // host copy; device copy lives at the start of the workspace
  std::vector<ccsd_objcoef_t> sched;
  std::vector<unsigned long long> cell_mask;
  std::vector<int> edge_ij, tri_ij;
  const float *weights = nullptr;
  size_t n_weights = 0;
  // workspace
  char *ws = nullptr;
  size_t ws_bytes = 0;
  DevPlan *dP = nullptr;
  float *flags = nullptr, *x = nullptr, *adj = nullptr, *r2 = nullptr;
  float *mx = nullptr, *madj = nullptr, *mr2 = nullptr;
  float *sx = nullptr, *sadj = nullptr, *sr2 = nullptr;
  float *H = nullptr, *P0 = nullptr, *P1 = nullptr, *norm_part = nullptr, *coef = nullptr;
  unsigned long long *zmask = nullptr, *zmask_eval = nullptr;
  float *g_stack = nullptr, *g_att = nullptr, *g_hmc = nullptr, *g_x0 = nullptr, *g_x1 = nullptr;
  float *g_big = nullptr;       // scratch of the large-graph pipeline [B][xp.big_total]
  uint8_t *gimg = nullptr;      // tc_gram: projection rows as operand chunks (tc_gram_prep_kernel)
  bool gimg_fresh = false;      // built from the CURRENT weight blob (reset by ccsd_plan_init / ccsd_score_eval: the caller may refresh the blob)
  StepDev *sd_dev = nullptr;    // device-resident step state of a graph replay (ccsd_plan_run)
  bool graph_capture = false;   // do_step is being captured: kernels read the step / diff_traj slots from sd_dev
  int use_graph = 0;            // graph-only plans: replay one captured step (latency path of small batches)
  float *g_hu = nullptr;        // [B][n1] hodge_u_kernel's sums for the plan's own flags (valid after ccsd_plan_init)
  bool hu_ready = false;
  float *traj_x = nullptr, *traj_adj = nullptr, *traj_r2 = nullptr;
  bool bound = false, inited = false;
  long long *trace = nullptr;   // debug timeline buffer for the tensor-core apply kernel
  unsigned long long seed = 0;
  long long sample_offset = 0;
  size_t apply_smem = 0;
  int64_t launches = 0;
  int use_tc = 0, use_tc_apply = 0, use_tc_fin = 0, use_tc_agg = 0;
#ifndef CCSD_EMU
  int use_tc_attn[CCSD_MAX_LAYERS] = {0};   // per attention layer: tcgen05 attention-channel kernel (tc_attn.cuh)
  TcAttnLayout tattn[CCSD_MAX_LAYERS];
  int use_tc_big = 0;                       // E > 192: K-chunked tcgen05 GEMMs for the Gram product and H . F (tc_r2big.cuh)
  int use_tc_edge[CCSD_MAX_LAYERS] = {0};   // per attention layer: per-edge MLP on tcgen05 (tc_edge.cuh)
  int use_hnorm = 0;                        // rank-2 Langevin norms from Gram quantities instead of a NORM pass (tc_hnorm.cuh)
  cudaStream_t side = nullptr;              // internal stream: the norm kernels run beside the x / adj pipeline
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_pending = false;
  // x / adj pipeline forks (launch_xa): [0] ScoreNetworkX's final MLP, [1] the hodge branch, [2] the node MLP of each attention
  // layer (beside its per-edge MLP) -- all independent of the attention chain on the caller's stream until the joins
  cudaGraph_t ggraph = nullptr;             // the captured step of the last graph replay (ccsd_plan_run) ...
  cudaGraphExec_t gexec = nullptr;          // ... kept until its replays have run (ev_graph)
  cudaEvent_t ev_graph = nullptr;
  cudaStream_t gs = nullptr;                // capture stream (the caller's may be the legacy default stream, which cannot capture)
  cudaStream_t xs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_x0 = nullptr, ev_xj[3] = {nullptr, nullptr, nullptr}, ev_xl = nullptr;
  float *Dg = nullptr, *Rs = nullptr;       // [B][E] diag(F F^T), F 1
  float *H2 = nullptr;                      // [B][E][Ep] H . H of large complexes (tc_r2big)
  int use_tc_xfin = 0;                      // ScoreNetworkX final MLP on tcgen05 (tc_xfin.cuh)
  TcXfinLayout txf;
  uint8_t *ximg = nullptr;                  // weight operand image of tc_xfin
#endif
  float *g_hcat = nullptr;                  // [B][fdimX x N4] node features x, h_1 .. h_D (tc_xfin / ScoreNetworkX_GMH)
  int apply_big = 0;   // E too large for the resident F column block: hf_gemm_kernel + r2_epi_kernel (scratch = sr2)
  // optional per-kernel timing (CUDA events on the launching stream)
  bool profiling = false;
  struct ProfRec { const char *name; void *e0, *e1; };
  std::vector<ProfRec> prof;
};

#ifdef CCSD_EMU
#define PROF_BEGIN(p, nm, stream) ((void)0)
#define PROF_END(p, stream) ((void)0)
#else
static void prof_begin(ccsd_plan *p, const char *name, void *stream) {
  if (!p->profiling) return;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, (cudaStream_t)stream);
  p->prof.push_back({name, (void *)a, (void *)b});
}
static void prof_end(ccsd_plan *p, void *stream) {
  if (!p->profiling) return;
  cudaEventRecord((cudaEvent_t)p->prof.back().e1, (cudaStream_t)stream);
}
#define PROF_BEGIN(p, nm, stream) prof_begin(p, nm, stream)
#define PROF_END(p, stream) prof_end(p, stream)
#endif

// ---------------------------------------------------------------------------------------------
static int a4(int v) { return (v + 3) & ~3; }
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

static int check_mlp(const ccsd_mlp_t &m, size_t nw, const char *name, int small) {
  if (m.nl < 1 || m.nl > CCSD_MAX_MLP) return fail(CCSD_ERR_UNSUPPORTED, std::string(name) + ": 1..4 linears supported");
  for (int l = 0; l < m.nl; ++l) {
    const int din = l == 0 ? m.din : m.dhid, dout = l == m.nl - 1 ? m.dout : m.dhid;
    const int opad = (dout + 7) / 8 * 8;
    if (m.w[l] < 0 || (m.w[l] & 3) || (size_t)m.w[l] + (size_t)din * opad > nw || m.b[l] < 0 || (m.b[l] & 3) ||
        (size_t)m.b[l] + opad > nw)
      return fail(CCSD_ERR_INVALID, std::string(name) + ": weight offsets out of range / unaligned");
    if (small && (din > SMALL_MAX || dout > SMALL_MAX))
      return fail(CCSD_ERR_UNSUPPORTED, std::string(name) + ": layer wider than 32 is not supported on this path");
  }
  return 0;
}
static int check_gcn(const ccsd_gcn_t &g, size_t nw, const char *name) {
  const int opad = (g.dout + 7) / 8 * 8;
  if (g.w < 0 || (g.w & 3) || (size_t)g.w + (size_t)g.din * opad > nw || g.b < 0 || (size_t)g.b + opad > nw)
    return fail(CCSD_ERR_INVALID, std::string(name) + ": weight offsets out of range / unaligned");
  return 0;
}

// a chain of AttentionLayers (ScoreNetworkA's trunk, ScoreNetworkX_GMH's layers): dims, weights, supported variants
static int check_attn_chain(const ccsd_attn_layer_t *layers, int n, int c_init, int F, int heads, size_t nw, int N, int *ch_total) {
  int ch = c_init, cin = c_init, kin = F;
  for (int l = 0; l < n; ++l) {
    const ccsd_attn_layer_t &ly = layers[l];
    if (ly.c_in != cin || ly.conv_in != kin) return fail(CCSD_ERR_INVALID, "AttentionLayer: layer chaining mismatch");
    if (ly.c_in > CCSD_MAX_CH || ly.c_out > CCSD_MAX_CH || ly.c_out < 1) return fail(CCSD_ERR_UNSUPPORTED, "AttentionLayer: more than 8 channels per layer");
    if (ly.attn_dim < heads || heads < 1) return fail(CCSD_ERR_INVALID, "AttentionLayer: attn_dim < num_heads");
    if (ly.conv_mlp && N > 64) return fail(CCSD_ERR_UNSUPPORTED, "conv == 'MLP' attention is not implemented on the large-graph pipeline (max_node_num > 64)");
    for (int c = 0; c < ly.c_in; ++c) {
      if (ly.conv_mlp) {
        if (int r = check_mlp(ly.qm[c], nw, "attn.gnn_q (MLP)", 0)) return r;
        if (int r = check_mlp(ly.km[c], nw, "attn.gnn_k (MLP)", 0)) return r;
        if (ly.qm[c].nl != 2 || ly.km[c].nl != 2 || ly.qm[c].din != kin || ly.qm[c].dout != ly.attn_dim || ly.km[c].dout != ly.attn_dim)
          return fail(CCSD_ERR_INVALID, "conv == 'MLP': Q / K must be 2-layer MLPs conv_in -> 2 attn_dim -> attn_dim");
      } else {
        if (int r = check_gcn(ly.q[c], nw, "attn.gnn_q")) return r;
        if (int r = check_gcn(ly.k[c], nw, "attn.gnn_k")) return r;
      }
      if (int r = check_gcn(ly.v[c], nw, "attn.gnn_v")) return r;
    }
    if (int r = check_mlp(ly.mlp, nw, "AttentionLayer.mlp", 0)) return r;
    if (int r = check_mlp(ly.multi_channel, nw, "AttentionLayer.multi_channel", 0)) return r;
    if (ly.mlp.din != 2 * ly.c_in || ly.mlp.dout != ly.c_out || ly.multi_channel.din != ly.c_in * ly.conv_out ||
        ly.multi_channel.dout != ly.conv_out)
      return fail(CCSD_ERR_INVALID, "AttentionLayer: MLP dims mismatch");
    ch += ly.c_out; cin = ly.c_out; kin = ly.conv_out;
  }
  *ch_total = ch;
  return 0;
}

static int validate(const ccsd_plan_desc_t &d, size_t nw) {
  if (d.B < 1 || d.N < 2 || d.F < 1) return fail(CCSD_ERR_INVALID, "B, N, F must be positive (N >= 2)");
  if (d.N > 64 && d.is_cc)
    return fail(CCSD_ERR_UNSUPPORTED, "max_node_num > 64 is only supported for graph-only plans (the large-graph pipeline has no hodge branch)");
  if (d.N > 1024) return fail(CCSD_ERR_UNSUPPORTED, "max_node_num > 1024 is not supported");
  if (d.sampler != CCSD_SAMPLER_PC && d.sampler != CCSD_SAMPLER_S4) return fail(CCSD_ERR_INVALID, "unknown sampler");
  if (d.n_lang_steps < 1 || d.n_lang_steps > 16) return fail(CCSD_ERR_UNSUPPORTED, "Langevin n_steps must be in 1..16");
  if (d.n_diff_steps < 1) return fail(CCSD_ERR_INVALID, "n_diff_steps must be >= 1");
  const ccsd_netx_t &X = d.netx;
  if (d.nets & 1) {
  if (X.nfeat != d.F || X.depth < 1 || X.depth > CCSD_MAX_LAYERS) return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkX: depth 1..8, nfeat == F");
  if (X.fdim != X.nfeat + X.depth * X.nhid) return fail(CCSD_ERR_INVALID, "ScoreNetworkX: fdim mismatch");
  if (X.gmh) {
    if (d.N > 64) return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkX_GMH is not implemented on the large-graph pipeline (max_node_num > 64)");
    if (X.gmh_c_init < 1 || X.gmh_c_init > CCSD_MAX_CH) return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkX_GMH: c_init 1..8");
    int chx = 0;
    if (int r = check_attn_chain(X.glayer, X.depth, X.gmh_c_init, d.F, X.gmh_heads, nw, d.N, &chx)) return r;
    for (int k = 0; k < X.depth; ++k)
      if (X.glayer[k].conv_out != X.nhid) return fail(CCSD_ERR_INVALID, "ScoreNetworkX_GMH: every layer must output nhid node features");
  } else
  for (int k = 0; k < X.depth; ++k)
    if (int r = check_gcn(X.gcn[k], nw, "ScoreNetworkX.layers")) return r;
  if (int r = check_mlp(X.fin, nw, "ScoreNetworkX.final", 0)) return r;
  if (X.fin.din != X.fdim || X.fin.dout != d.F) return fail(CCSD_ERR_INVALID, "ScoreNetworkX.final dims mismatch");
  }
  const ccsd_neta_t &A = d.neta;
  if (d.is_cc && (d.E != d.N * (d.N - 1) / 2 || d.K < 1)) return fail(CCSD_ERR_INVALID, "rank-2 dims mismatch");
  if (d.nets & 2) {
  if (A.num_layers < 1 || A.num_layers > CCSD_MAX_LAYERS || A.c_init < 1 || A.c_init > CCSD_MAX_CH)
    return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkA: num_layers 1..8, c_init 1..8");
  int ch = A.c_init;
  if (int r = check_attn_chain(A.layer, A.num_layers, A.c_init, d.F, A.num_heads, nw, d.N, &ch)) return r;
  if (A.is_cc && !d.is_cc) return fail(CCSD_ERR_INVALID, "ScoreNetworkA_CC needs a combinatorial-complex plan (is_cc)");
  if (A.is_cc && A.base_cc) {
    if (A.num_layers_h < 1 || A.num_layers_h > CCSD_MAX_HODGE_LAYERS)
      return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkA_Base_CC: num_layers_h must be 1 or 2");
    int hin = A.c_init;
    for (int l = 0; l < A.num_layers_h; ++l) {
      const ccsd_hbase_layer_t &h = A.hbase[l];
      if (h.c_in != hin || h.c_in > CCSD_MAX_CH || h.c_out > CCSD_MAX_CH || h.hid < 1 || h.hid > (l == 0 ? SMALL_MAX : 8))
        return fail(CCSD_ERR_UNSUPPORTED, "HodgeBaselineLayer: channels <= 8, hidden width <= 32 (first layer) / 8 (second layer)");
      if (int r = check_mlp(h.mlp_hodge, nw, "mlp_hodge", 1)) return r;
      if (h.mlp_hodge.din != h.c_in) return fail(CCSD_ERR_INVALID, "HodgeBaselineLayer: mlp_hodge input width mismatch");
      const int hp = (h.hid + 7) / 8 * 8;
      for (int c = 0; c < h.c_in; ++c)
        if (h.w1[c] < 0 || h.w2[c] < 0 || h.b1[c] < 0 || h.b2[c] < 0 || (size_t)h.w1[c] + (size_t)d.E * hp > nw ||
            (size_t)h.w2[c] + (size_t)d.E * hp > nw || (size_t)h.b2[c] + d.E > nw)
          return fail(CCSD_ERR_INVALID, "HodgeBaselineLayer: weight offsets out of range");
      ch += h.c_out; hin = h.c_out;
    }
    ch += A.c_init;
  } else if (A.is_cc) {
    if (A.num_layers_h < 1 || A.num_layers_h > CCSD_MAX_HODGE_LAYERS)
      return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkA_CC: num_layers_h must be 1 or 2 (the dense H @ rank2 value path of deeper stacks is not implemented)");
    int hin = A.c_init;
    for (int l = 0; l < A.num_layers_h; ++l) {
      const ccsd_hodge_layer_t &h = A.hodge[l];
      if (h.c_in != hin || h.c_in > CCSD_MAX_CH || h.c_out > CCSD_MAX_CH || h.attn_dim > SMALL_MAX || h.attn_dim < A.num_heads_h || A.num_heads_h < 1)
        return fail(CCSD_ERR_UNSUPPORTED, "HodgeAdjAttentionLayer: channels <= 8, attn_dim <= 32");
      if (A.n_proj_rows[l] != h.c_in * 2 * h.attn_dim) return fail(CCSD_ERR_INVALID, "hodge projection rows mismatch");
      if (int r = check_mlp(h.mlp_attention, nw, "mlp_attention", 1)) return r;
      if (int r = check_mlp(h.mlp_value, nw, "mlp_value", 1)) return r;
      ch += h.c_out; hin = h.c_out;
    }
    ch += A.c_init;
    const int Kw = a4(d.K);
    const size_t rows = (size_t)A.n_proj_rows[0] + (A.num_layers_h == 2 ? A.n_proj_rows[1] : 0);
    if (A.proj_w < 0 || (A.proj_w & 3) || (size_t)A.proj_w + rows * Kw > nw) return fail(CCSD_ERR_INVALID, "hodge projection weights out of range");
  }
  if (ch != A.fdim) return fail(CCSD_ERR_INVALID, "ScoreNetworkA: fdim does not match the channel stack (the reference would raise a shape error too)");
  if (int r = check_mlp(A.fin, nw, "ScoreNetworkA.final", 0)) return r;
  if (A.fin.din != A.fdim || A.fin.dout != 1) return fail(CCSD_ERR_INVALID, "ScoreNetworkA.final dims mismatch");
  }
  if (d.is_cc && (d.nets & 4)) {
    const ccsd_netf_t &Fn = d.netf;
    if (Fn.cnum != 2) return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkF: only cnum == 2 (F, H F) is implemented");
    if (Fn.num_layers < 1 || Fn.num_layers > CCSD_MAX_F_LAYERS) return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkF: num_layers 1..4");
    int fd = Fn.cnum, fin = Fn.cnum;
    for (int l = 0; l < Fn.num_layers; ++l) {
      if (int r = check_mlp(Fn.layer[l], nw, "HodgeNetworkLayer", 1)) return r;
      if (Fn.layer[l].din != fin) return fail(CCSD_ERR_INVALID, "ScoreNetworkF: layer chaining mismatch");
      fin = Fn.layer[l].dout; fd += fin;
    }
    if (fd != Fn.fdim || fd > SMALL_MAX) return fail(CCSD_ERR_UNSUPPORTED, "ScoreNetworkF: fdim mismatch or > 32");
    if (int r = check_mlp(Fn.fin, nw, "ScoreNetworkF.final", 1)) return r;
    if (Fn.fin.din != fd || Fn.fin.dout != 1) return fail(CCSD_ERR_INVALID, "ScoreNetworkF.final dims mismatch");
  }
  return 0;
}

static void make_layout(const ccsd_plan_desc_t &d, XpLayout &L) {
  memset(&L, 0, sizeof(L));
  const int N = d.N, F = d.F;
  const ccsd_netx_t &X = d.netx;
  const ccsd_neta_t &A = d.neta;
  const int N4 = a4(N), NT = N * (N + 1) / 2, ldp = a4(NT);
  L.N4 = N4; L.NT = NT; L.ldp = ldp;
  const int T = NT >= 96 ? 128 : 64;
  L.Tx = L.Th = L.Tm = T;
  L.Tc = L.Tf = 64;   // measured on B200 (community_small_CC): 64 threads beat 32 and 128 for the per-channel CTAs
  // tuning overrides (experiments): threads per CTA of the individual pipeline kernels
  auto envi = [](const char *nm, int dflt) { const char *e = getenv(nm); return e ? atoi(e) : dflt; };
  L.Tx = envi("CCSD_XP_TX", L.Tx); L.Tc = envi("CCSD_XP_TC", L.Tc); L.Tf = envi("CCSD_XP_TF", L.Tf);
  L.Tm = envi("CCSD_XP_TM", L.Tm);
  int nh_max = 1, ad_max = 1, cin_max = 1, kin_max = F, mc_hid = 1, mc_o1 = 1, eh = 1, eh_bufs = 1, nch_max = 1;
  // every AttentionLayer of the plan: ScoreNetworkA's trunk, then ScoreNetworkX_GMH's layers (they share the kernels)
  const int nA = (d.nets & 2) ? A.num_layers : 0, nG = ((d.nets & 1) && X.gmh) ? X.depth : 0;
  int gmh_ch = 0;
  for (int l = 0; l < nG; ++l) gmh_ch += X.glayer[l].c_out;
  for (int l = 0; l < nA + nG; ++l) {
    const ccsd_attn_layer_t &ly = l < nA ? A.layer[l] : X.glayer[l - nA];
    const int heads = imax(l < nA ? A.num_heads : X.gmh_heads, 1);
    nh_max = imax(nh_max, ly.conv_out);
    ad_max = imax(ad_max, ly.attn_dim);
    cin_max = imax(cin_max, ly.c_in);
    kin_max = imax(kin_max, ly.conv_in);
    mc_hid = imax(mc_hid, imax(ly.multi_channel.dhid, ly.multi_channel.dout));
    mc_o1 = imax(mc_o1, ly.multi_channel.nl == 1 ? ly.multi_channel.dout : ly.multi_channel.dhid);
    if (ly.mlp.nl > 1) eh = imax(eh, ly.mlp.dhid);
    if (ly.mlp.nl > 2) eh_bufs = 2;
    const int ds = imax(ly.attn_dim / heads, 1);   // torch.split: ceil(ad / (ad / heads)) chunks
    nch_max = imax(nch_max, (ly.attn_dim + ds - 1) / ds);
  }
  int o = 0;
  auto take = [&](int n) { int r = o; o += a4(n); return r; };
  // ---- x_net_kernel ----
  int din_max = F;
  for (int k = 0; k < X.depth; ++k) din_max = imax(din_max, X.gcn[k].din);
  o = 0;
  L.x_flags = take(N4); L.x_dvec = take(N4);
  L.x_adj = take(imax(imax(A.c_init, nG ? X.gmh_c_init : 1), 1) * ldp);
  L.x_an = take(N * N4);
  L.x_x0 = take(F * N4);
  L.x_sx = take(F * N4);
  L.x_red = take(40);
  L.x_hcat = take(imax(X.depth * X.nhid, 1) * N4);
  L.x_ax = take(din_max * N4);
  L.x_ha = take(imax(X.fin.dhid, 1) * N4);
  L.x_hb = X.fin.nl > 2 ? take(imax(X.fin.dhid, 1) * N4) : L.x_ha;
  L.x_total = o;
  // ---- attn_channel_kernel ----
  o = 0;
  L.c_dvec = take(N4);
  L.c_adj = take(ldp);
  L.c_an = take(N * N4);
  L.c_xin = take(kin_max * N4);
  L.c_ax = take(kin_max * N4);
  L.c_q = take(ad_max * N4);
  L.c_k = take(ad_max * N4);
  L.c_v = take(nh_max * N4);
  L.c_atp = take(nch_max * ldp);
  {
    int wmax = 0;
    for (int l = 0; l < nA + nG; ++l) {
      const ccsd_attn_layer_t &ly = l < nA ? A.layer[l] : X.glayer[l - nA];
      const int o1 = ly.multi_channel.nl == 1 ? ly.multi_channel.dout : ly.multi_channel.dhid;
      const int r8a = (ly.attn_dim + 7) / 8 * 8, r8n = (ly.conv_out + 7) / 8 * 8, r8o = (o1 + 7) / 8 * 8;
      wmax = imax(wmax, ly.conv_in * (2 * r8a + r8n) + ly.conv_out * r8o);
    }
#if XP_STAGE_W
    L.c_w = take(wmax);
#else
    L.c_w = o; (void)wmax;
#endif
  }
  L.c_mh = take(2 * ad_max * N4);   // hidden layer of the conv == "MLP" Q / K networks
  L.c_total = o;
  // ---- attn_finish_kernel ----
  o = 0;
  L.f_flags = take(N4);
  L.f_hs = take(mc_hid * N4);
  L.f_hs2 = take(mc_hid * N4);
  L.f_eha = take(eh * ldp);
  L.f_ehb = eh_bufs == 2 ? take(eh * ldp) : L.f_eha;
  L.f_total = o;
  // ---- hodge_kernel ----
  o = 0;
  L.h_flags = take(N4);
  if (d.is_cc && A.is_cc && A.num_layers_h == 2) {
    const int E = d.E;
    L.lde = E | 1;
    L.h_hq = take(A.c_init * E * A.hodge[0].attn_dim);
    L.h_hk = take(A.c_init * E * A.hodge[0].attn_dim);
    L.h_h1 = take(A.hodge[0].c_out * E * L.lde);
    L.h_hdeg = take(A.hodge[0].c_out * E);
    L.h_att1 = take(A.hodge[0].c_out * E);
    L.h_alpha = take(E);
    L.h_part = take(128);
  }
  L.h_total = o;
  // ---- hodge_base_kernel ----
  o = L.h_flags + a4(N4);
  if (d.is_cc && A.is_cc && A.base_cc) {
    L.hb_u0 = take(A.c_init * d.E * ((A.hbase[0].hid + 7) / 8 * 8));
    L.hb_fe = take(d.E);
    L.hb_tri = take(d.E);
  }
  L.hb_total = o;
  // ---- afinal_kernel: 64-row chunks of node pairs ----
  o = 0;
  const int fin_h = imax(A.fin.nl > 1 ? A.fin.dhid : 1, 1);
  L.m_rows = imin(64, ldp);
  L.m_nchunk = (NT + L.m_rows - 1) / L.m_rows;
  L.m_flags = take(N4);
  L.m_out = take(L.m_rows);
  L.m_red = take(40);
  L.m_fa = take(fin_h * L.m_rows);
  L.m_fb = A.fin.nl > 2 ? take(fin_h * L.m_rows) : L.m_fa;
  L.m_total = o;
  // ---- global scratch ----
  L.g_stack = imax(imax(A.fdim, nG ? X.gmh_c_init + gmh_ch : 1), 1) * ldp;
  L.g_att = imax(cin_max, nG ? X.gmh_c_init : 1) * ldp;
  L.mc_o1_max = mc_o1;
  L.g_hmc = cin_max * mc_o1 * N4;
  L.g_x = imax(F, nh_max) * N4;
  // ---- large-graph pipeline (big_pipe.cuh) ----
  {
    static const bool force_big = getenv("CCSD_B200_FORCE_BIG") != nullptr;   // A/B switch for tests: small graphs through the large-graph kernels
    L.big = (!d.is_cc && (N > 64 || force_big)) ? 1 : 0;
  }
  if (L.big) {
    const int Np = (N + 7) / 8 * 8;
    L.big_Np = Np; L.big_PS = N * Np;
    L.big_nrc = (N + BIG_RC - 1) / BIG_RC;
    L.big_nseg = (N + BIG_SEG - 1) / BIG_SEG;
    int adp_max = 8, nhp_max = 8, cout_max = 1, sm_node = 0, sm_edge = 0;
    for (int l = 0; l < A.num_layers; ++l) {
      const ccsd_attn_layer_t &ly = A.layer[l];
      adp_max = imax(adp_max, (ly.attn_dim + 7) / 8 * 8);
      nhp_max = imax(nhp_max, (ly.conv_out + 7) / 8 * 8);
      cout_max = imax(cout_max, ly.c_out);
      const ccsd_mlp_t &mc = ly.multi_channel, &me = ly.mlp;
      sm_node = imax(sm_node, a4(ly.conv_out) * BIG_RC + (mc.nl > 1 ? mc.dhid : 1) * BIG_RC * (mc.nl > 2 ? 2 : 1));
      sm_edge = imax(sm_edge, ly.c_out * BIG_SEG + (me.nl > 1 ? me.dhid : 1) * BIG_SEG * (me.nl > 2 ? 2 : 1));
    }
    int xdp = 8;
    for (int k = 0; k < X.depth; ++k) xdp = imax(xdp, (X.gcn[k].dout + 7) / 8 * 8);
    L.big_sm_node = sm_node + 8; L.big_sm_edge = sm_edge + 8;
    L.big_sm_fin = BIG_SEG + 40 + (A.fin.nl > 1 ? A.fin.dhid : 1) * BIG_SEG * (A.fin.nl > 2 ? 2 : 1) + 8;
    L.big_sm_xfin = a4(F) * BIG_RCX + 40 + (X.fin.nl > 1 ? X.fin.dhid : 1) * BIG_RCX * (X.fin.nl > 2 ? 2 : 1) + 8;
    int ob = 0;
    auto tk = [&](long long n) { int r = ob; ob += (int)((n + 7) / 8 * 8); return r; };
    const int cm = imax(cin_max, imax(A.c_init, 1));
    L.big_S = tk((long long)imax(A.fdim, 1) * L.big_PS);
    L.big_ATT = tk((long long)cm * L.big_PS);
    L.big_Y = tk((long long)imax(cm * N * (2 * adp_max + nhp_max), N * xdp));
    L.big_T_xw = imin(128, ((2 * adp_max + nhp_max) / 8 * (BIG_RC / 4) + 31) / 32 * 32);
    L.big_TQK = tk((long long)cm * 2 * adp_max * Np);
    L.big_TV = tk((long long)cm * nhp_max * Np);
    L.big_XF0 = tk((long long)imax(kin_max, nh_max) * Np);
    L.big_XF1 = tk((long long)imax(kin_max, nh_max) * Np);
    L.big_DV = tk((long long)cm * Np);
    L.big_HC = tk((long long)imax(X.fdim, F) * Np);
    L.big_total = ob;
    // the per-graph-tile scratch is not used
    L.g_stack = L.g_att = L.g_hmc = L.g_x = 8;
  }
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int ccsd_plan_desc_size(void) { return (int)sizeof(ccsd_plan_desc_t); }
int ccsd_objcoef_size(void) { return (int)sizeof(ccsd_objcoef_t); }
const char *ccsd_last_error(void) { return g_err.c_str(); }
const char *ccsd_version(void) {
#ifdef CCSD_EMU
  return "ccsd_b200 0.1 (HOST EMULATION BUILD - tests only)";
#else
  return "ccsd_b200 0.1 (sm_100a)";
#endif
}

static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

struct WsLayout {
  size_t plan, sched, cells, edges, zmask, zmask_eval, tri, gstack, gatt, ghmc, gx0, gx1, ghcat, ximg, dg, rs, h2, hu, sd, gimg, gbig, flags, x, adj, r2, mx, madj, mr2, sx, sadj, sr2, H, P0, P1, norm, coef, total;
};
static WsLayout ws_layout(const ccsd_plan *p) {
  const ccsd_plan_desc_t &d = p->hp.d;
  WsLayout w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return r; };
  const size_t B = d.B, N = d.N, F = d.F, E = d.E, K = d.K;
  w.plan = take(sizeof(DevPlan));
  w.sched = take(p->sched.size() * sizeof(ccsd_objcoef_t));
  w.cells = take(imax(1, d.K) * sizeof(unsigned long long));
  w.edges = take(imax(1, d.E) * 2 * sizeof(int));
  w.zmask = take(B * 8);
  w.zmask_eval = take(B * 8);
  w.tri = take((size_t)p->hp.xp.ldp * sizeof(int));
  w.gstack = take(B * (size_t)p->hp.xp.g_stack * 4);
  w.gatt = take(B * (size_t)p->hp.xp.g_att * 4);
  w.ghmc = take(B * (size_t)p->hp.xp.g_hmc * 4);
  w.gx0 = take(B * (size_t)p->hp.xp.g_x * 4);
  w.gx1 = take(B * (size_t)p->hp.xp.g_x * 4);
  const bool gmh = (d.nets & 1) && d.netx.gmh;
#ifndef CCSD_EMU
  w.dg = take(p->use_hnorm ? B * E * 4 : 16);
  w.rs = take(p->use_hnorm ? B * E * 4 : 16);
  w.h2 = take((p->use_hnorm && p->use_tc_big) ? B * E * (size_t)a4((int)E) * 4 + 64 : 16);
  w.ghcat = take((p->use_tc_xfin || gmh || ((d.nets & 1) && !p->hp.xp.big)) ? B * (size_t)d.netx.fdim * p->hp.xp.N4 * 4 : 16);
  w.ximg = take(p->use_tc_xfin ? (size_t)p->txf.img_bytes : 16);
#else
  w.ghcat = take(gmh ? B * (size_t)d.netx.fdim * p->hp.xp.N4 * 4 : 16); w.ximg = take(16); w.dg = take(16); w.rs = take(16); w.h2 = take(16);
#endif
  w.hu = take(p->hp.p1_fold ? B * (size_t)imax(1, d.neta.n_proj_rows[1]) * 4 : 16);
  w.sd = take(sizeof(StepDev));
#ifndef CCSD_EMU
  w.gimg = take(p->use_tc ? tc_gram_img_bytes((int)K, p->hp.PR0) : 16);
#else
  w.gimg = take(16);
#endif
  w.gbig = take(p->hp.xp.big ? B * (size_t)p->hp.xp.big_total * 4 : 16);
  w.flags = take(B * N * 4);
  w.x = take(B * N * F * 4); w.adj = take(B * N * N * 4); w.r2 = take(B * E * K * 4 + 16);
  w.mx = take(B * N * F * 4); w.madj = take(B * N * N * 4); w.mr2 = take(B * E * K * 4 + 16);
  w.sx = take(B * N * F * 4); w.sadj = take(B * N * N * 4); w.sr2 = take(B * E * K * 4 + 16);
  w.H = take(B * E * (size_t)a4((int)E) * 4 + 64);
  w.P0 = take(B * E * (size_t)imax(1, p->hp.PR0) * 4);
  w.P1 = take(B * E * (size_t)imax(1, p->hp.PR1) * 4);
  w.norm = take(3 * B * (size_t)p->hp.ntile_max * 2 * 4);
  w.coef = take(64);
  w.total = o;
  return w;
}

int ccsd_plan_create(const ccsd_plan_desc_t *desc, const ccsd_objcoef_t *schedule_host, const float *weights_dev,
                     size_t n_weights, ccsd_plan_t **out) {
  if (!desc || !schedule_host || !weights_dev || !out) return fail(CCSD_ERR_INVALID, "null argument");
  if (int r = validate(*desc, n_weights)) return r;
  ccsd_plan *p = new ccsd_plan();
  p->hp.d = *desc;
  const ccsd_plan_desc_t &d = p->hp.d;
  make_layout(d, p->hp.xp);
  p->weights = weights_dev;
  p->n_weights = n_weights;
  p->sched.assign(schedule_host, schedule_host + (size_t)d.n_diff_steps * 3);
  const bool hodge = d.is_cc && (d.nets & 2) && d.neta.is_cc && !d.neta.base_cc;   // the baseline network has no rank-2 projections
  p->hp.PR0h = hodge ? d.neta.n_proj_rows[0] : 0;
  p->hp.PR0 = p->hp.PR0h;
  p->hp.PR1 = (hodge && d.neta.num_layers_h == 2) ? d.neta.n_proj_rows[1] : 0;
  p->hp.p1_fold = 0;
  if (p->hp.PR1 > 0 && d.neta.hodge[0].mlp_value.nl == 1 && d.neta.hodge[0].mlp_value.dout == 1 &&
      p->hp.PR0h + p->hp.PR1 <= 64) {
    // the second hodge layer's projections become extra Gram columns (xa_pipe.cuh, hodge_kernel)
    p->hp.p1_fold = 1;
    p->hp.PR0 = p->hp.PR0h + p->hp.PR1;
    p->hp.PR1 = 0;
  }
  if (p->hp.p1_fold) {   // shared memory of hodge_kernel for the folded projections
    XpLayout &XL2 = p->hp.xp;
    const int n1 = d.neta.n_proj_rows[1];
    XL2.h_p1 = XL2.h_total; XL2.h_total += a4(d.E * n1);
    XL2.h_u = XL2.h_total; XL2.h_total += a4(n1);
  }
  p->hp.Kp = a4(d.K);
  p->hp.Ep = a4(d.E);
  p->hp.ntile_r2 = d.is_cc ? (d.K + APPLY_TN - 1) / APPLY_TN : 1;
  p->hp.ntile_adj = p->hp.xp.m_nchunk;
  p->hp.ntile_x = 1;
  if (p->hp.xp.big) { p->hp.ntile_adj = d.N * p->hp.xp.big_nseg; p->hp.ntile_x = (d.N + BIG_RCX - 1) / BIG_RCX; }
  p->hp.ntile_max = imax(imax(imax(1, p->hp.ntile_r2), p->hp.ntile_adj), p->hp.ntile_x);
  p->hp.f_mode = 0; p->hp.f_nlin = 0;
  p->hp.ap_group = d.is_cc ? imax(1, imin(8, 192 / imax(d.E, 1))) : 1;
  p->hp.gram_group = 1;
  if (d.is_cc && (d.nets & 4)) {
    const ccsd_netf_t &Fn = d.netf;
    bool w8 = Fn.fdim <= 40, w4 = true;
    int nlin = 0;
    for (int l = 0; l < Fn.num_layers; ++l) {
      const ccsd_mlp_t &M = Fn.layer[l];
      if (M.din > 8 || M.dout > 8 || (M.nl > 1 && M.dhid > 8)) w8 = false;
      if (M.din > 4 || M.dout > 4 || (M.nl > 1 && M.dhid > 4)) w4 = false;
      nlin += M.nl;
    }
    if (Fn.affine) p->hp.f_mode = 1;
    else if (w8 && Fn.fin.nl == 1) { p->hp.f_mode = w4 ? 3 : 2; p->hp.f_nlin = nlin; }   // 3: four entries at a time (tensor-core apply kernel)
    else if (w8 && Fn.fin.nl == 2 && Fn.num_layers <= NETF_F2_MAXL && netf_stage_floats(Fn, nlin, 4) <= 2048) {
      p->hp.f_mode = 4; p->hp.f_nlin = nlin;                                             // two-Linear final MLP (Base_CC checkpoints)
    }
  }
  p->apply_smem = d.is_cc ? ((size_t)d.E * (APPLY_TN + 4) + 16 * 68 + 40 + (size_t)netf_stage_floats(d.netf, p->hp.f_nlin, p->hp.f_mode) + 4) * 4 : 0;
  const XpLayout &XL = p->hp.xp;
  const size_t xp_max = XL.big ? (size_t)imax(imax(XL.big_sm_node, XL.big_sm_edge), imax(XL.big_sm_fin, XL.big_sm_xfin)) * 4
                               : (size_t)imax(imax(imax(XL.x_total, XL.c_total), imax(XL.f_total, imax(XL.h_total, XL.hb_total))), XL.m_total) * 4;
  if (p->apply_smem > 227 * 1024) {   // large E (grid_small_CC): GEMM into scratch + element-wise epilogue
    p->apply_big = 1;
    p->apply_smem = (size_t)(40 + netf_stage_floats(d.netf, p->hp.f_nlin, p->hp.f_mode) + 4) * 4;
    p->hp.ntile_r2 = R2EPI_CHUNKS;
    p->hp.ntile_max = imax(p->hp.ntile_max, p->hp.ntile_r2);
  }
  if (xp_max > 227 * 1024) {
    char buf[200];
    snprintf(buf, sizeof buf, "graph tile does not fit shared memory (x/adj pipeline %zu B, apply %zu B > 227 KB): N/E too large for the resident-tile kernels",
             xp_max, p->apply_smem);
    delete p;
    return fail(CCSD_ERR_UNSUPPORTED, buf);
  }
  // tables
  {
    const int N = d.N;
    p->tri_ij.assign((size_t)XL.ldp, 0);
    int t = 0;
    if (N <= 255)   // (i << 8) | j packing; the large-graph pipeline does not use the table
      for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) p->tri_ij[t++] = (i << 8) | j;
  }
  if (d.is_cc) {
    const int N = d.N;
    p->edge_ij.resize((size_t)d.E * 2);
    int e = 0;
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) { p->edge_ij[2 * e] = i; p->edge_ij[2 * e + 1] = j; ++e; }
    // cells: combinations(range(N), dd) for dd = d_min..d_max, lexicographic (cc_utils.py:72-76)
    p->cell_mask.reserve(d.K);
    for (int dd = d.d_min; dd <= d.d_max; ++dd) {
      if (dd < 1 || dd > N) continue;
      std::vector<int> c(dd);
      for (int i = 0; i < dd; ++i) c[i] = i;
      while (true) {
        unsigned long long m = 0;
        for (int i = 0; i < dd; ++i) m |= 1ull << c[i];
        p->cell_mask.push_back(m);
        int i = dd - 1;
        while (i >= 0 && c[i] == N - dd + i) --i;
        if (i < 0) break;
        ++c[i];
        for (int j = i + 1; j < dd; ++j) c[j] = c[j - 1] + 1;
      }
    }
    if ((int)p->cell_mask.size() != d.K) {
      delete p;
      return fail(CCSD_ERR_INVALID, "K does not equal sum_d C(N, d) for d_min..d_max");
    }
  }
#ifndef CCSD_EMU
  // the attribute is per (function, device), not per plan: only ever raise it (several plans may coexist)
  {
    static CcsdSmemAttr at_big[4], at_xp[6], at_apply[4];
    int bad = 0;
    if (XL.big) {
      bad |= ccsd_ensure_smem(big_node_kernel, xp_max, at_big[0]) | ccsd_ensure_smem(big_edge_kernel, xp_max, at_big[1]) |
             ccsd_ensure_smem(big_final_kernel, xp_max, at_big[2]) | ccsd_ensure_smem(big_xfin_kernel, xp_max, at_big[3]);
    } else {
      bad |= ccsd_ensure_smem(x_net_kernel, xp_max, at_xp[0]) | ccsd_ensure_smem(attn_channel_kernel, xp_max, at_xp[1]) |
             ccsd_ensure_smem(attn_finish_kernel, xp_max, at_xp[2]) | ccsd_ensure_smem(hodge_kernel, xp_max, at_xp[3]) |
             ccsd_ensure_smem(hodge_base_kernel, xp_max, at_xp[4]) | ccsd_ensure_smem(afinal_kernel, xp_max, at_xp[5]);
    }
    if (d.is_cc && !p->apply_big)
      bad |= ccsd_ensure_smem(apply_kernel<0>, p->apply_smem, at_apply[0]) | ccsd_ensure_smem(apply_kernel<1>, p->apply_smem, at_apply[1]) |
             ccsd_ensure_smem(apply_kernel<2>, p->apply_smem, at_apply[2]) | ccsd_ensure_smem(apply_kernel<4>, p->apply_smem, at_apply[3]);
    if (bad) {
      delete p;
      return fail(CCSD_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(cudaGetLastError()));
    }
  }
  p->use_tc = d.is_cc ? tc_gram_supported(d.E, d.K, p->hp.PR0) : 0;
  if (p->use_tc) {   // small complexes: stack G samples per Gram work unit
    int G = imin(8, 192 / imax(d.E, 1));
    while (G > 1 && tc_gram_ncols(G * d.E, p->hp.PR0) > 256) --G;
    if (getenv("CCSD_B200_NO_GRAM_GROUP")) G = 1;   // A/B switch for tests and profiling
    if (const char *e = getenv("CCSD_B200_GRAM_GROUP")) { const int g = atoi(e); if (g >= 1 && g <= G) G = g; }
    p->hp.gram_group = imax(G, 1);
  }
  p->use_tc_apply = (d.is_cc && (d.nets & 4) && !p->apply_big) ? tc_apply_supported(d.E, d.K) : 0;
  // Wider non-affine ScoreNetworkF (entry paths 0 / 2 / 4: the Base_CC checkpoints) is ill conditioned -- two fp32
  // evaluations that differ in summation order already disagree by 4.5e-6 and the 2^-16 product error of the bf16x3
  // contractions grows to 6e-5 per evaluation -- so its rank-2 contractions stay on the fp32 FMA kernels (flat 1e-4 bar).
  if (d.is_cc && (d.nets & 4) && p->hp.f_mode != 1 && p->hp.f_mode != 3) p->use_tc = p->use_tc_apply = 0;
  p->use_tc_fin = (d.nets & 2) ? tc_afinal_supported(d.neta, d.neta.fdim) : 0;
  p->use_tc_agg = (XL.big && (d.nets & 2)) ? 1 : 0;
  if (getenv("CCSD_B200_NO_TC_AGG")) p->use_tc_agg = 0;
  if ((d.nets & 1) && !getenv("CCSD_B200_NO_TC_XFIN")) p->use_tc_xfin = tc_xfin_layout(d, XL, p->txf);
  p->use_tc_big = (d.is_cc && !getenv("CCSD_B200_NO_TC_BIG")) ? tc_r2big_supported(d, p->hp.PR0) : 0;
  if (d.is_cc && (d.nets & 4) && p->hp.f_mode != 1 && p->hp.f_mode != 3) p->use_tc_big = 0;   // ill-conditioned non-affine networks stay in fp32
  if ((d.nets & 2) && !getenv("CCSD_B200_NO_TC_EDGE"))
    for (int l = 0; l < d.neta.num_layers; ++l) p->use_tc_edge[l] = tc_edge_supported(d, XL, d.neta.layer[l]);
  if ((d.nets & 2) && !getenv("CCSD_B200_NO_TC_ATTN"))
    for (int l = 0; l < d.neta.num_layers; ++l) p->use_tc_attn[l] = tc_attn_layout(d, XL, d.neta.layer[l], p->tattn[l]);
  if (const char *e = getenv("CCSD_B200_NO_TC")) if (e[0] == '1') { p->use_tc = p->use_tc_apply = p->use_tc_fin = p->use_tc_agg = p->use_tc_xfin = p->use_tc_big = 0; memset(p->use_tc_attn, 0, sizeof p->use_tc_attn); memset(p->use_tc_edge, 0, sizeof p->use_tc_edge); }  // A/B switch for tests and profiling
  // rank-2 Langevin norms from Gram quantities (affine ScoreNetworkF on the tensor-core Gram / apply kernels, PC + Langevin)
  p->use_hnorm = p->use_tc && p->use_tc_apply && d.sampler == CCSD_SAMPLER_PC && d.use_corrector && tc_hnorm_supported(d, p->hp.f_mode) &&
                 tc_gram_supported(d.E, d.K, p->hp.PR0 + 1) && !getenv("CCSD_B200_NO_HNORM");
  if (p->use_tc_big && d.sampler == CCSD_SAMPLER_PC && d.use_corrector && (d.nets & 4) && p->hp.f_mode == 1 && d.netf.use_hodge_mask &&
      !getenv("CCSD_B200_NO_HNORM"))
    p->use_hnorm = 1;   // large complexes: H . H by the K-chunked GEMM, reductions by hnorm_big_kernel
  if (p->use_hnorm && !getenv("CCSD_B200_NO_SIDE_STREAM")) {
    if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      delete p;
      return fail(CCSD_ERR_CUDA, "cannot create the internal stream / events");
    }
  }
#ifndef CCSD_EMU
  p->use_graph = !d.is_cc && !getenv("CCSD_B200_NO_GRAPH");   // (complexes: the rank-2 passes dominate and a replay measured 8 % slower)
  if (!XL.big && !getenv("CCSD_B200_NO_SIDE_STREAM")) {
    bool ok = cudaEventCreateWithFlags(&p->ev_x0, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&p->ev_xl, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && ok; ++i)
      ok = cudaStreamCreateWithFlags(&p->xs[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&p->ev_xj[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      delete p;
      return fail(CCSD_ERR_CUDA, "cannot create the internal streams / events of the x / adj pipeline");
    }
  }
#endif
  if (p->use_hnorm && p->hp.gram_group > 1)
    while (p->hp.gram_group > 1 && tc_gram_ncols(p->hp.gram_group * d.E, p->hp.PR0 + 1) > 256) --p->hp.gram_group;
  if (p->use_tc_fin) {   // norm partial slots = 128-row tiles per graph
    p->hp.ntile_adj = XL.big ? d.N * ((d.N + 127) / 128) : (p->hp.xp.NT + 127) / 128;
    p->hp.ntile_max = imax(p->hp.ntile_max, p->hp.ntile_adj);
  }
  if (p->use_tc_apply) {
    if (int r = tc_apply_prepare()) { delete p; return fail(CCSD_ERR_CUDA, "tc_apply_prepare failed"); }
  }
  if (p->use_tc) {
    if (int r = tc_gram_prepare()) { delete p; return fail(CCSD_ERR_CUDA, "tc_gram_prepare failed"); }
  }
#endif
  *out = p;
  return 0;
}

#ifndef CCSD_EMU
static void drop_graph(ccsd_plan *p) {
  if (p->gexec) {
    if (p->ev_graph) cudaEventSynchronize(p->ev_graph);   // its replays have run (normally long ago)
    cudaGraphExecDestroy(p->gexec);
    p->gexec = nullptr;
  }
  if (p->ggraph) { cudaGraphDestroy(p->ggraph); p->ggraph = nullptr; }
}
#endif

void ccsd_plan_destroy(ccsd_plan_t *plan) {
  if (!plan) return;
#ifndef CCSD_EMU
  for (auto &r : plan->prof) { cudaEventDestroy((cudaEvent_t)r.e0); cudaEventDestroy((cudaEvent_t)r.e1); }
  if (plan->side) { cudaStreamSynchronize(plan->side); cudaStreamDestroy(plan->side); }
  drop_graph(plan);
  if (plan->ev_graph) cudaEventDestroy(plan->ev_graph);
  if (plan->gs) cudaStreamDestroy(plan->gs);
  for (int i = 0; i < 3; ++i) {
    if (plan->xs[i]) { cudaStreamSynchronize(plan->xs[i]); cudaStreamDestroy(plan->xs[i]); }
    if (plan->ev_xj[i]) cudaEventDestroy(plan->ev_xj[i]);
  }
  if (plan->ev_x0) cudaEventDestroy(plan->ev_x0);
  if (plan->ev_xl) cudaEventDestroy(plan->ev_xl);
  if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
  if (plan->ev_join) cudaEventDestroy(plan->ev_join);
#endif
  delete plan;
}

size_t ccsd_plan_workspace_bytes(const ccsd_plan_t *plan) { return plan ? ws_layout(plan).total : 0; }

int ccsd_plan_bind(ccsd_plan_t *p, void *workspace_dev, size_t bytes, void *stream) {
  if (!p || !workspace_dev) return fail(CCSD_ERR_INVALID, "null argument");
  const WsLayout w = ws_layout(p);
  if (bytes < w.total) return fail(CCSD_ERR_INVALID, "workspace too small");
  if (((uintptr_t)workspace_dev) & 255) return fail(CCSD_ERR_INVALID, "workspace must be 256-byte aligned");
  char *ws = (char *)workspace_dev;
  p->ws = ws; p->ws_bytes = bytes;
  p->dP = (DevPlan *)(ws + w.plan);
  p->flags = (float *)(ws + w.flags);
  p->x = (float *)(ws + w.x); p->adj = (float *)(ws + w.adj); p->r2 = (float *)(ws + w.r2);
  p->mx = (float *)(ws + w.mx); p->madj = (float *)(ws + w.madj); p->mr2 = (float *)(ws + w.mr2);
  p->sx = (float *)(ws + w.sx); p->sadj = (float *)(ws + w.sadj); p->sr2 = (float *)(ws + w.sr2);
  p->H = (float *)(ws + w.H); p->P0 = (float *)(ws + w.P0); p->P1 = (float *)(ws + w.P1);
  p->norm_part = (float *)(ws + w.norm); p->coef = (float *)(ws + w.coef);
  p->g_stack = (float *)(ws + w.gstack); p->g_att = (float *)(ws + w.gatt); p->g_hmc = (float *)(ws + w.ghmc);
  p->g_x0 = (float *)(ws + w.gx0); p->g_x1 = (float *)(ws + w.gx1);
  p->g_big = (float *)(ws + w.gbig);
  p->g_hcat = (float *)(ws + w.ghcat);
#ifndef CCSD_EMU
  p->ximg = (uint8_t *)(ws + w.ximg);
  p->Dg = (float *)(ws + w.dg); p->Rs = (float *)(ws + w.rs); p->H2 = (float *)(ws + w.h2);
#endif
  p->g_hu = (float *)(ws + w.hu); p->hu_ready = false;
  p->sd_dev = (StepDev *)(ws + w.sd);
  p->gimg = (uint8_t *)(ws + w.gimg);
  if (p->hp.xp.big) {
    // pad columns / rows of the planes are read as don't-care operands: make them finite once
    const size_t nb = (size_t)p->hp.d.B * p->hp.xp.big_total * 4;
#ifdef CCSD_EMU
    memset(p->g_big, 0, nb);
#else
    if (cudaMemsetAsync(p->g_big, 0, nb, (cudaStream_t)stream) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaMemsetAsync (large-graph scratch)");
#endif
  }
  p->hp.tri_ij = (const int *)(ws + w.tri);
  p->zmask = (unsigned long long *)(ws + w.zmask); p->zmask_eval = (unsigned long long *)(ws + w.zmask_eval);
  p->hp.W = p->weights;
  p->hp.sched = (const ccsd_objcoef_t *)(ws + w.sched);
  p->hp.cell_mask = (const unsigned long long *)(ws + w.cells);
  p->hp.edge_ij = (const int *)(ws + w.edges);
  if (int r = dev_copy(ws + w.plan, &p->hp, sizeof(DevPlan), stream)) return r;
  if (int r = dev_copy(ws + w.sched, p->sched.data(), p->sched.size() * sizeof(ccsd_objcoef_t), stream)) return r;
  if (!p->cell_mask.empty())
    if (int r = dev_copy(ws + w.cells, p->cell_mask.data(), p->cell_mask.size() * 8, stream)) return r;
  if (!p->edge_ij.empty())
    if (int r = dev_copy(ws + w.edges, p->edge_ij.data(), p->edge_ij.size() * 4, stream)) return r;
  if (int r = dev_copy(ws + w.tri, p->tri_ij.data(), p->tri_ij.size() * 4, stream)) return r;
  p->bound = true;
  p->inited = false;
  return 0;
}

int ccsd_plan_set_traj(ccsd_plan_t *p, float *tx, float *tadj, float *tr2) {
  if (!p) return fail(CCSD_ERR_INVALID, "null plan");
  p->traj_x = tx; p->traj_adj = tadj; p->traj_r2 = tr2;
  return 0;
}

static int grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

int ccsd_plan_init(ccsd_plan_t *p, const float *flags_dev, const float *px, const float *padj, const float *pr2,
                   uint64_t seed, int64_t sample_offset, void *stream) {
  if (!p || !flags_dev) return fail(CCSD_ERR_INVALID, "null argument");
  if (!p->bound) return fail(CCSD_ERR_STATE, "ccsd_plan_bind must be called before ccsd_plan_init");
  const ccsd_plan_desc_t &d = p->hp.d;
  if (int r = dev_copy(p->flags, flags_dev, (size_t)d.B * d.N * 4, stream)) return r;
  p->seed = seed; p->sample_offset = sample_offset;
  p->gimg_fresh = false;
#ifndef CCSD_EMU
  // the caller may have refreshed the weight blob since the last run: rebuild the bf16 operand image of the X network's MLP
  if (p->use_tc_xfin)
    if (tc_xfin_prep(p->dP, p->txf, p->ximg, stream)) return fail(CCSD_ERR_CUDA, "tc_xfin_prep launch failed");
#endif
  CCSD_LAUNCH(zmask_kernel, dim3(grid_for(d.B), 1, 1), 256, 0, stream, p->flags, p->zmask, d.B, d.N);
  if (p->hp.p1_fold && !p->hp.xp.big) {   // flag-only sums of the hodge branch: once per run
    CCSD_LAUNCH(hodge_u_kernel, dim3(d.B, 1, 1), 128, 96 * 4, stream, p->dP, p->flags, p->g_hu, p->hp.xp.Th);
    p->hu_ready = true;
  }
  InitArgs a;
  a.flags = p->flags; a.px = px; a.padj = padj; a.pr2 = pr2; a.x = p->x; a.adj = p->adj; a.r2 = p->r2;
  a.nz.seed = seed; a.nz.sample_offset = sample_offset; a.nz.step = -1; a.nz.sd = nullptr;
  const size_t units = d.is_cc ? (size_t)d.B * d.E * (p->hp.Kp / 4) : (size_t)d.B * d.N * d.N;
  CCSD_LAUNCH(init_kernel, dim3(grid_for(units), d.is_cc ? 3 : 2, 1), 256, 0, stream, p->dP, a);
  p->launches++;
  // means start as the prior (returned if zero steps are run)
  if (int r = dev_copy(p->mx, p->x, (size_t)d.B * d.N * d.F * 4, stream)) return r;
  if (int r = dev_copy(p->madj, p->adj, (size_t)d.B * d.N * d.N * 4, stream)) return r;
  if (d.is_cc)
    if (int r = dev_copy(p->mr2, p->r2, (size_t)d.B * d.E * d.K * 4, stream)) return r;
  p->inited = true;
  return dev_check("init_kernel");
}

// Large-graph pipeline (big_pipe.cuh): same contract as launch_xa
static int launch_xa_big(ccsd_plan *p, const XaArgs &a, void *stream) {
  const ccsd_plan_desc_t &d = p->hp.d;
  const XpLayout &L = p->hp.xp;
  const ccsd_neta_t &A = d.neta;
  const ccsd_netx_t &X = d.netx;
  BigArgs g; memset(&g, 0, sizeof g);
  g.a = a; g.base = p->g_big;
  const int B = d.B, nrc = L.big_nrc, nrca = (d.N + BIG_RCA - 1) / BIG_RCA;
  // threads per CTA (<= 128, the kernels' launch bound); environment overrides are tuning experiments
  static const auto envt = [](const char *nm, int dflt) { const char *e = getenv(nm); const int v = e ? atoi(e) : dflt; return v >= 32 && v <= 128 ? v / 32 * 32 : dflt; };
  static const int T_pow = envt("CCSD_BIG_T_POW", 128), T_xw = envt("CCSD_BIG_T_XW", 0), T_agg = envt("CCSD_BIG_T_AGG", 0),
                   T_node = envt("CCSD_BIG_T_NODE", 64), T_edge = envt("CCSD_BIG_T_EDGE", 32), T_fin = envt("CCSD_BIG_T_FIN", 128),
                   T_xfin = envt("CCSD_BIG_T_XFIN", 64);
  static const bool edge_fast = getenv("CCSD_BIG_EDGE_GENERIC") == nullptr;   // A/B switch: row-tile edge kernel
#define BIG_LAUNCH(kern, grid, thr, smem) do { PROF_BEGIN(p, #kern, stream); CCSD_LAUNCH(kern, grid, thr, smem, stream, p->dP, g); PROF_END(p, stream); p->launches++; } while (0)
  BIG_LAUNCH(big_prep_kernel, dim3(imin(148 * 2, (d.N * L.big_Np + 255) / 256), B, 1), 256, 0);
  const int c0 = (a.which & 2) ? A.c_init : 1;
  for (int c = 1; c < c0; ++c) { g.c = c; BIG_LAUNCH(big_pow_kernel, dim3(nrc, 1, B), T_pow, 0); }
  g.ch_in = 0; g.nch = c0;
  BIG_LAUNCH(big_deg_kernel, dim3((d.N + 127) / 128, c0, B), 128, 0);
  if (a.which & 1) {
    g.xmode = 1;
    int in_row = 0, out_row = d.F;
    for (int k = 0; k < X.depth; ++k) {
      g.gk = k; g.in_row = in_row; g.out_row = out_row;
      BIG_LAUNCH(big_xw_kernel, dim3(nrc, 1, B), 32, 0);    // 4 x 8 items (nhid <= 32)
      BIG_LAUNCH(big_agg_kernel, dim3(nrca, 1, B), 32, 0);   // 4 x 8 items of 8 x 8 (nhid <= 32)
      in_row = out_row; out_row += X.gcn[k].dout;
    }
    BIG_LAUNCH(big_xfin_kernel, dim3((d.N + BIG_RCX - 1) / BIG_RCX, 1, B), T_xfin, (size_t)L.big_sm_xfin * 4);
  }
  if (!(a.which & 2)) return dev_check("large-graph x network");
  g.xmode = 0;
  int ch_in = 0, ch_out = A.c_init;
  float *xf0 = p->g_big + L.big_XF0, *xf1 = p->g_big + L.big_XF1;
  for (int l = 0; l < A.num_layers; ++l) {
    const ccsd_attn_layer_t &ly = A.layer[l];
    g.layer = l; g.ch_in = ch_in; g.ch_out = ch_out; g.xin = xf0; g.xout = xf1;
    if (l > 0) { g.nch = ly.c_in; BIG_LAUNCH(big_deg_kernel, dim3((d.N + 127) / 128, ly.c_in, B), 128, 0); }
    BIG_LAUNCH(big_xw_kernel, dim3(nrc, ly.c_in, B), T_xw ? T_xw : L.big_T_xw, 0);
#ifndef CCSD_EMU
    if (p->use_tc_agg && tc_agg_supported(ly)) {
      PROF_BEGIN(p, "tc_agg_kernel", stream);
      if (tc_agg_launch(p->dP, p->hp, g, stream)) return fail(CCSD_ERR_CUDA, "tc_agg launch failed");
      PROF_END(p, stream);
      p->launches++;
    } else
#endif
    BIG_LAUNCH(big_agg_kernel, dim3(nrca, ly.c_in, B), T_agg ? T_agg : L.big_T_xw, 0);
    const int nb = (d.N + 3) / 4, nblk = nb * (nb + 1) / 2;
    BIG_LAUNCH(big_attn_kernel, dim3((nblk + 127) / 128, ly.c_in, B), 128, 0);
    BIG_LAUNCH(big_node_kernel, dim3(nrc, 1, B), T_node, (size_t)L.big_sm_node * 4);
    if (edge_fast && big_edge_fast_ok(ly)) {
      BIG_LAUNCH(big_edge_pair_kernel, dim3(((d.N + BIG_EROWS - 1) / BIG_EROWS) * ((d.N + BIG_ESEG - 1) / BIG_ESEG), 1, B), BIG_ESEG, (size_t)4 * BIG_EW * 4);
      const int nt = (d.N + 31) / 32;
      BIG_LAUNCH(big_mirror_kernel, dim3(nt * (nt + 1) / 2, ly.c_out, B), 256, (size_t)32 * 33 * 4);
    }
    else
      BIG_LAUNCH(big_edge_kernel, dim3(d.N * L.big_nseg, 1, B), T_edge, (size_t)L.big_sm_edge * 4);
    ch_in = ch_out; ch_out += ly.c_out;
    float *t = xf0; xf0 = xf1; xf1 = t;
  }
  g.ch_out = ch_out;   // planes the final MLP reads
#ifndef CCSD_EMU
  if (p->use_tc_fin) {
    XaArgs af = a;
    af.g_stack = p->g_big + L.big_S;
    PROF_BEGIN(p, "tc_afinal_kernel", stream);
    if (tc_afinal_launch(p->dP, p->hp, af, ch_out, stream)) return fail(CCSD_ERR_CUDA, "tc_afinal launch failed");
    PROF_END(p, stream);
    p->launches++;
    return dev_check("large-graph x/adj network pipeline");
  }
#endif
  BIG_LAUNCH(big_final_kernel, dim3(d.N * L.big_nseg, 1, B), T_fin, (size_t)L.big_sm_fin * 4);
#undef BIG_LAUNCH
  return dev_check("large-graph x/adj network pipeline");
}

// ScoreNetworkX / ScoreNetworkA(_CC) pipeline (xa_pipe.cuh) for the networks selected by a.which
static int launch_xa(ccsd_plan *p, XaArgs a, void *stream) {
  if (p->hp.xp.big) return launch_xa_big(p, a, stream);
  const ccsd_plan_desc_t &d = p->hp.d;
  const XpLayout &L = p->hp.xp;
  const ccsd_neta_t &A = d.neta;
  a.g_stack = p->g_stack; a.g_att = p->g_att; a.g_hmc = p->g_hmc; a.g_x0 = p->g_x0; a.g_x1 = p->g_x1;
  if ((a.which & 1) && d.netx.gmh) {
    // ScoreNetworkX_GMH (ScoreNetwork_X.py:280-314): adjacency powers + x hand-over, the AttentionLayers (the same
    // kernels as ScoreNetworkA's trunk, reading netx.glayer), then the final MLP + x epilogue
    const ccsd_netx_t &X = d.netx;
    XaArgs g = a;
    g.which = 1; g.gmh = 1; g.gmh_phase = 1; g.g_hcat = p->g_hcat;
    PROF_BEGIN(p, "x_net_kernel", stream);
    CCSD_LAUNCH(x_net_kernel, dim3(d.B, 1, 1), L.Tx, (size_t)L.x_total * 4, stream, p->dP, g);
    PROF_END(p, stream);
    p->launches++;
    int gin = 0, gout = X.gmh_c_init;
    const float *gxin = p->g_x0;
    float *gxout = p->g_x1;
    for (int l = 0; l < X.depth; ++l) {
      g.layer = l; g.ch_in = gin; g.ch_out = gout; g.g_xin = gxin; g.g_xout = gxout; g.skip_edge = 0;
      PROF_BEGIN(p, "attn_channel_kernel", stream);
      CCSD_LAUNCH(attn_channel_kernel, dim3(X.glayer[l].c_in, d.B, 1), L.Tc, (size_t)L.c_total * 4, stream, p->dP, g);
      PROF_END(p, stream);
      PROF_BEGIN(p, "attn_finish_kernel", stream);
      CCSD_LAUNCH(attn_finish_kernel, dim3(d.B, 1, 1), L.Tf, (size_t)L.f_total * 4, stream, p->dP, g);
      PROF_END(p, stream);
      p->launches += 2;
      gin = gout; gout += X.glayer[l].c_out;
      const float *t = gxout; gxout = (float *)gxin; gxin = t;
    }
    bool done = false;
#ifndef CCSD_EMU
    if (p->use_tc_xfin) {
      g.gmh = 0;
      PROF_BEGIN(p, "tc_xfin_kernel", stream);
      if (tc_xfin_launch(p->dP, p->hp, g, p->txf, p->g_hcat, X.fdim * L.N4, p->ximg, stream)) return fail(CCSD_ERR_CUDA, "tc_xfin launch failed");
      PROF_END(p, stream);
      p->launches++;
      done = true;
    }
#endif
    if (!done) {
      g.gmh_phase = 2;
      PROF_BEGIN(p, "x_net_kernel", stream);
      CCSD_LAUNCH(x_net_kernel, dim3(d.B, 1, 1), L.Tx, (size_t)L.x_total * 4, stream, p->dP, g);
      PROF_END(p, stream);
      p->launches++;
    }
    a.which &= ~1;
    if (!a.which) return dev_check("ScoreNetworkX_GMH");
  }
  bool split_x = false;   // the fp32 final MLP of ScoreNetworkX as a second x_net_kernel launch on a forked stream
#ifndef CCSD_EMU
  if (p->use_tc_xfin && (a.which & 1)) a.g_hcat = p->g_hcat;
  // Networks the tensor-core final MLP does not cover (fdim > 128: ENZYMES_small(_CC), ego_small): x_net_kernel stops after
  // the GCN stack and hands [x, h_1 .. h_D] over in global memory, exactly like the tensor-core path; its final MLP + epilogue
  // (the gmh_phase == 2 entry) then runs beside the hodge branch and the attention chain instead of in front of them.
  if (!p->use_tc_xfin && !d.netx.gmh && (a.which & 3) == 3 && p->xs[0] && !p->profiling) { a.g_hcat = p->g_hcat; split_x = true; }
#endif
  PROF_BEGIN(p, "x_net_kernel", stream);
  CCSD_LAUNCH(x_net_kernel, dim3(d.B, 1, 1), L.Tx, (size_t)L.x_total * 4, stream, p->dP, a);
  PROF_END(p, stream);
  p->launches++;
  // Forks: everything below only needs x_net_kernel's outputs until the joins, so ScoreNetworkX's final MLP, the hodge branch and
  // the attention chain run on three streams (they are latency-bound kernels that fill each other's idle issue slots).  Serial
  // when profiling (the per-kernel events live on the caller's stream) or when only one network is evaluated.
  bool forked = false, join_x = false, join_h = false;
#ifndef CCSD_EMU
  forked = p->xs[0] && !p->profiling && (a.which & 2);
  if (forked && cudaEventRecord(p->ev_x0, (cudaStream_t)stream) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream fork failed");
  if (a.g_hcat) {
    void *sx = stream;
    if (forked) {
      if (cudaStreamWaitEvent(p->xs[0], p->ev_x0, 0) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream fork failed");
      sx = (void *)p->xs[0];
    }
    if (split_x) {
      XaArgs a2 = a;
      a2.which = 1; a2.gmh_phase = 2;
      CCSD_LAUNCH(x_net_kernel, dim3(d.B, 1, 1), L.Tx, (size_t)L.x_total * 4, sx, p->dP, a2);
    } else {
      PROF_BEGIN(p, "tc_xfin_kernel", stream);
      if (tc_xfin_launch(p->dP, p->hp, a, p->txf, p->g_hcat, d.netx.fdim * L.N4, p->ximg, sx)) return fail(CCSD_ERR_CUDA, "tc_xfin launch failed");
      PROF_END(p, stream);
    }
    p->launches++;
    if (forked) {
      if (cudaEventRecord(p->ev_xj[0], p->xs[0]) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream join failed");
      join_x = true;
    }
  }
#endif
  if (!(a.which & 2)) return dev_check("x_net_kernel");
  if (A.is_cc && !A.base_cc) {
    // The hodge branch (cc_utils.py:1503-1588, hodge_layers.py) reads the initial channels (adjacency powers) and the Gram
    // projections only and writes its own planes of the channel stack: independent of the attention layers
    int ch_h = A.c_init;
    for (int l = 0; l < A.num_layers; ++l) ch_h += A.layer[l].c_out;   // first hodge channel
    void *sh = stream;
#ifndef CCSD_EMU
    if (forked) {
      if (cudaStreamWaitEvent(p->xs[1], p->ev_x0, 0) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream fork failed");
      sh = (void *)p->xs[1];
    }
#endif
    if (p->hp.PR1 > 0) {
      // projections of hodge layer 1 (value MLP with a non-linearity: not foldable into the Gram product)
      Proj1Args q; q.r2 = a.r2; q.flags = a.flags; q.g_stack = p->g_stack; q.g_stack_stride = L.g_stack; q.ldp = L.ldp;
      q.P1 = p->P1;
      q.epc = imax(1, imin(8, (48 * 1024) / (p->hp.Kp * 4)));
      PROF_BEGIN(p, "proj1_kernel", stream);
      CCSD_LAUNCH(proj1_kernel, dim3((d.E + q.epc - 1) / q.epc, d.B, 1), 128, ((size_t)q.epc * p->hp.Kp + 2 * 72 + 8) * 4, sh, p->dP, q);
      PROF_END(p, stream);
      p->launches++;
    }
    XaArgs h = a;
    h.ch_in = ch_h;
    h.g_hu = (p->hu_ready && a.flags == p->flags) ? p->g_hu : nullptr;   // (score_eval with the caller's own flags: computed in the kernel)
    PROF_BEGIN(p, "hodge_kernel", stream);
    CCSD_LAUNCH(hodge_kernel, dim3(d.B, 1, 1), L.Th, (size_t)L.h_total * 4, sh, p->dP, h);
    PROF_END(p, stream);
    p->launches++;
#ifndef CCSD_EMU
    if (forked) {
      if (cudaEventRecord(p->ev_xj[1], p->xs[1]) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream join failed");
      join_h = true;
    }
#endif
  }
  int ch_in = 0, ch_out = A.c_init;
  const float *xin = p->g_x0;
  float *xout = p->g_x1;
  for (int l = 0; l < A.num_layers; ++l) {
    const ccsd_attn_layer_t &ly = A.layer[l];
    a.layer = l; a.ch_in = ch_in; a.ch_out = ch_out; a.g_xin = xin; a.g_xout = xout;
#ifndef CCSD_EMU
    if (p->use_tc_attn[l]) {
      PROF_BEGIN(p, "tc_attn_kernel", stream);
      a.trace = (l == 1 && a.mode == MODE_PRED && !getenv("CCSD_B200_TRACE_GRAM")) ? p->trace : nullptr;   // debug timeline: layer 1 of the predictor's evaluation
      if (tc_attn_launch(p->dP, p->hp, a, p->tattn[l], stream)) return fail(CCSD_ERR_CUDA, "tc_attn launch failed");
      PROF_END(p, stream);
    } else
#endif
    {
      PROF_BEGIN(p, "attn_channel_kernel", stream);
      CCSD_LAUNCH(attn_channel_kernel, dim3(ly.c_in, d.B, 1), L.Tc, (size_t)L.c_total * 4, stream, p->dP, a);
      PROF_END(p, stream);
    }
#ifndef CCSD_EMU
    a.skip_edge = p->use_tc_edge[l];
#endif
    void *sf = stream;
#ifndef CCSD_EMU
    const bool fork_l = forked && p->use_tc_edge[l];   // the node MLP (attn_finish) beside the per-edge MLP (tc_edge)
    if (fork_l) {
      if (cudaEventRecord(p->ev_xl, (cudaStream_t)stream) != cudaSuccess || cudaStreamWaitEvent(p->xs[2], p->ev_xl, 0) != cudaSuccess)
        return fail(CCSD_ERR_CUDA, "stream fork failed");
      sf = (void *)p->xs[2];
    }
#endif
    PROF_BEGIN(p, "attn_finish_kernel", stream);
    CCSD_LAUNCH(attn_finish_kernel, dim3(d.B, 1, 1), L.Tf, (size_t)L.f_total * 4, sf, p->dP, a);
    PROF_END(p, stream);
    p->launches += 2;
#ifndef CCSD_EMU
    if (p->use_tc_edge[l]) {
      PROF_BEGIN(p, "tc_edge_kernel", stream);
      if (tc_edge_launch(p->dP, p->hp, a, stream)) return fail(CCSD_ERR_CUDA, "tc_edge launch failed");
      PROF_END(p, stream);
      p->launches++;
    }
    if (fork_l) {
      if (cudaEventRecord(p->ev_xj[2], p->xs[2]) != cudaSuccess || cudaStreamWaitEvent((cudaStream_t)stream, p->ev_xj[2], 0) != cudaSuccess)
        return fail(CCSD_ERR_CUDA, "stream join failed");
    }
#endif
    ch_in = ch_out;
    ch_out += ly.c_out;
    const float *t = xout; xout = (float *)xin; xin = t;
  }
  int fd_have = ch_out;
  if (A.is_cc && A.base_cc) {
    a.ch_in = ch_out;   // first hodge channel
    PROF_BEGIN(p, "hodge_base_kernel", stream);
    CCSD_LAUNCH(hodge_base_kernel, dim3(d.B, 1, 1), L.Th, (size_t)L.hb_total * 4, stream, p->dP, a);
    PROF_END(p, stream);
    p->launches++;
    fd_have += A.c_init + A.hbase[0].c_out + (A.num_layers_h == 2 ? A.hbase[1].c_out : 0);
  }
  int fd_hodge = 0;
  if (A.is_cc && !A.base_cc) fd_hodge = A.c_init + A.hodge[0].c_out + (A.num_layers_h == 2 ? A.hodge[1].c_out : 0);
  fd_have += fd_hodge;
  a.ch_out = fd_have;   // channels the final MLP reads
  auto join = [&]() -> int {   // the forked streams rejoin the caller's stream
#ifndef CCSD_EMU
    if (join_h && cudaStreamWaitEvent((cudaStream_t)stream, p->ev_xj[1], 0) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream join failed");
    if (join_x && cudaStreamWaitEvent((cudaStream_t)stream, p->ev_xj[0], 0) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream join failed");
#endif
    join_h = join_x = false;
    return 0;
  };
  if (int r = join()) return r;
#ifndef CCSD_EMU
  if (p->use_tc_fin) {
    PROF_BEGIN(p, "tc_afinal_kernel", stream);
    if (tc_afinal_launch(p->dP, p->hp, a, fd_have, stream)) return fail(CCSD_ERR_CUDA, "tc_afinal launch failed");
    PROF_END(p, stream);
    p->launches++;
    return dev_check("x/adj network pipeline");
  }
#endif
  PROF_BEGIN(p, "afinal_kernel", stream);
  CCSD_LAUNCH(afinal_kernel, dim3(L.m_nchunk, d.B, 1), L.Tm, (size_t)L.m_total * 4, stream, p->dP, a);
  PROF_END(p, stream);
  p->launches++;
  return dev_check("x/adj network pipeline");
}

static void launch_apply(ccsd_plan *p, const ApplyArgs &q, void *stream) {
#ifndef CCSD_EMU
  if (p->use_tc_apply) { tc_apply_launch(p->dP, p->hp, q, stream); return; }
#endif
  if (p->apply_big) {
    const ccsd_plan_desc_t &d = p->hp.d;
#ifndef CCSD_EMU
    if (p->use_tc_big) {
      PROF_END(p, stream);                                   // (close the caller's "apply_kernel" record: it times nothing)
      PROF_BEGIN(p, "tc_r2big_kernel<hf>", stream);
      if (tc_r2big_hf(p->dP, p->hp, q.r2, q.H, p->sr2, stream)) { fail(CCSD_ERR_CUDA, "tc_r2big (H F) launch failed"); return; }
      PROF_END(p, stream);
      PROF_BEGIN(p, "r2_epi_kernel", stream);
    } else
#endif
    {
      HfArgs h; h.r2 = q.r2; h.H = q.H; h.hf = p->sr2;
      CCSD_LAUNCH(hf_gemm_kernel, dim3((d.K + GRAM_BN - 1) / GRAM_BN, (d.E + GRAM_BM - 1) / GRAM_BM, d.B), 256,
                  2 * GRAM_BK * (GRAM_BM + 4) * 4, stream, p->dP, h);
    }
    const dim3 ge(R2EPI_CHUNKS, d.B, 1);
    const float *hf = p->sr2;
    if (p->hp.f_mode == 1) CCSD_LAUNCH(r2_epi_kernel<1>, ge, 256, p->apply_smem, stream, p->dP, q, hf);
    else if (p->hp.f_mode == 4) CCSD_LAUNCH(r2_epi_kernel<4>, ge, 256, p->apply_smem, stream, p->dP, q, hf);
    else if (p->hp.f_mode >= 2) CCSD_LAUNCH(r2_epi_kernel<2>, ge, 256, p->apply_smem, stream, p->dP, q, hf);
    else CCSD_LAUNCH(r2_epi_kernel<0>, ge, 256, p->apply_smem, stream, p->dP, q, hf);
    p->launches++;
    return;
  }
  const dim3 grid(p->hp.ntile_r2, p->hp.d.B, 1);
  if (p->hp.f_mode == 1) CCSD_LAUNCH(apply_kernel<1>, grid, 256, p->apply_smem, stream, p->dP, q);
  else if (p->hp.f_mode == 4) CCSD_LAUNCH(apply_kernel<4>, grid, 256, p->apply_smem, stream, p->dP, q);
  else if (p->hp.f_mode >= 2) CCSD_LAUNCH(apply_kernel<2>, grid, 256, p->apply_smem, stream, p->dP, q);
  else CCSD_LAUNCH(apply_kernel<0>, grid, 256, p->apply_smem, stream, p->dP, q);
}

// H, P0 (and P1) of the rank-2 tensor `r2` (+ adj for P1)
static int launch_rank2_pre(ccsd_plan *p, const float *r2, const float *adj, const float *flags, void *stream) {
  const ccsd_plan_desc_t &d = p->hp.d;
#ifndef CCSD_EMU
  if (p->use_tc) {
    PROF_BEGIN(p, "tc_gram_kernel", stream);
    if (p->hp.PR0 > 0 && !p->gimg_fresh) {
      tc_gram_prep_kernel<<<(d.K + TG_BK - 1) / TG_BK, 128, 0, (cudaStream_t)stream>>>(p->dP, p->gimg);
      p->gimg_fresh = true;
      p->launches++;
    }
    if (int r = tc_gram_launch(p->dP, p->hp, r2, p->H, p->P0, p->use_hnorm ? p->Dg : nullptr, p->use_hnorm ? p->Rs : nullptr, stream, getenv("CCSD_B200_TRACE_GRAM") ? p->trace : nullptr, p->hp.PR0 > 0 ? p->gimg : nullptr)) return fail(CCSD_ERR_CUDA, "tc_gram launch failed");
    PROF_END(p, stream);
    p->launches++;
  } else
#endif
#ifndef CCSD_EMU
  if (p->use_tc_big) {
    PROF_BEGIN(p, "tc_r2big_kernel<gram>", stream);
    if (tc_r2big_gram(p->dP, p->hp, r2, p->H, p->P0, p->use_hnorm ? p->Dg : nullptr, p->use_hnorm ? p->Rs : nullptr, stream)) return fail(CCSD_ERR_CUDA, "tc_r2big (Gram) launch failed");
    PROF_END(p, stream);
    p->launches++;
  } else
#endif
  {
    GramArgs g; g.r2 = r2; g.H = p->H; g.P0 = p->P0;
    const int ncols = d.E + p->hp.PR0;
    PROF_BEGIN(p, "gram_kernel", stream);
    CCSD_LAUNCH(gram_kernel, dim3((ncols + GRAM_BN - 1) / GRAM_BN, (d.E + GRAM_BM - 1) / GRAM_BM, d.B), 256,
                2 * GRAM_BK * (GRAM_BM + 4) * 4, stream, p->dP, g);
    PROF_END(p, stream);
    p->launches++;
  }
  return dev_check("rank2 pre-pass");
}

static int do_step(ccsd_plan *p, int step, const float *nx, const float *nadj, const float *nr2, int write_mean,
                   void *stream) {
  const ccsd_plan_desc_t &d = p->hp.d;
  if ((d.nets & 3) != 3 || (d.is_cc && !(d.nets & 4)))
    return fail(CCSD_ERR_STATE, "sampling needs ScoreNetworkX, ScoreNetworkA and (for CC) ScoreNetworkF in the plan");
  const size_t sx = (size_t)d.B * d.N * d.F, sa = (size_t)d.B * d.N * d.N, sr = (size_t)d.B * d.E * d.K;
  const int s4 = d.sampler == CCSD_SAMPLER_S4;
  NoiseCtx nz; nz.seed = p->seed; nz.sample_offset = p->sample_offset; nz.step = step;
  nz.sd = p->graph_capture ? p->sd_dev : nullptr;   // captured step: index and diff_traj slots live on the device
  float *tx = (p->traj_x && !p->graph_capture) ? p->traj_x + (size_t)step * d.N * d.F : nullptr;
  float *ta = (p->traj_adj && !p->graph_capture) ? p->traj_adj + (size_t)step * d.N * d.N : nullptr;
  float *tr = (p->traj_r2 && d.is_cc) ? p->traj_r2 + (size_t)step * d.E * d.K : nullptr;

  // rank-2 apply pass: H F + ScoreNetworkF + the mode's epilogue
  auto apply_pass = [&](int mode, int slot) -> int {
    ApplyArgs q; memset(&q, 0, sizeof q);
    q.r2 = p->r2; q.H = p->H; q.flags = p->flags; q.mode = mode; q.slot = slot; q.denoise = d.denoise; q.nz = nz;
    q.norm_part = p->norm_part; q.coef = p->coef; q.zmask = p->zmask;
    q.trace = (mode == MODE_CORR || mode == MODE_SCORE) ? p->trace : nullptr;   // debug timeline of the Langevin-correction (S4: score) pass
    q.noise = nr2 ? nr2 + (size_t)slot * sr : nullptr;
    if (mode == MODE_SCORE) q.out = p->sr2;
    else if (mode == MODE_CORR) q.out = p->r2;
    else if (mode == MODE_PRED) { q.out = p->r2; q.mean = p->mr2; q.write_mean = write_mean; q.traj = tr; }
    PROF_BEGIN(p, p->use_tc_apply ? "tc_apply_kernel" : "apply_kernel", stream);
    launch_apply(p, q, stream);
    PROF_END(p, stream);
    p->launches++;
    return dev_check("apply pass");
  };
  // ||score||^2 and ||z||^2 of the rank-2 object for the Langevin step size: from Gram quantities + a Philox-only
  // reduction when ScoreNetworkF is affine (tc_hnorm.cuh), else a NORM pass over the state
  auto norm_pass = [&](int slot) -> int {
#ifndef CCSD_EMU
    if (p->use_hnorm) {
      // fork: everything enqueued on `stream` so far (the Gram kernel) precedes the norm kernels; they run on the internal
      // stream beside whatever the caller's stream gets next (the x / adj pipeline) and are joined before coef_kernel
      void *ns = stream;
      if (p->side) {
        if (cudaEventRecord(p->ev_fork, (cudaStream_t)stream) != cudaSuccess || cudaStreamWaitEvent(p->side, p->ev_fork, 0) != cudaSuccess)
          return fail(CCSD_ERR_CUDA, "stream fork failed");
        ns = (void *)p->side;
      }
      if (p->use_tc_big) {
        PROF_BEGIN(p, "tc_r2big_kernel<hh>", ns);
        if (tc_r2big_gram_x(p->dP, p->hp, p->H, d.E, p->hp.Ep, 0, 0, p->H2, nullptr, nullptr, nullptr, ns)) return fail(CCSD_ERR_CUDA, "tc_r2big (H H) launch failed");
        PROF_END(p, ns);
        HnormBigArgs hb; memset(&hb, 0, sizeof hb);
        hb.H = p->H; hb.H2 = p->H2; hb.Dg = p->Dg; hb.Rs = p->Rs; hb.flags = p->flags; hb.norm_part = p->norm_part; hb.step = step;
        PROF_BEGIN(p, "hnorm_big_kernel", ns);
        CCSD_LAUNCH(hnorm_big_kernel, dim3(d.B, 1, 1), 256, 64 * 4, ns, p->dP, hb);
        PROF_END(p, ns);
        p->launches++;
      } else {
        TcHnormArgs h; memset(&h, 0, sizeof h);
        h.H = p->H; h.Dg = p->Dg; h.Rs = p->Rs; h.flags = p->flags; h.norm_part = p->norm_part; h.step = step; h.G = p->hp.ap_group;
        PROF_BEGIN(p, "tc_hnorm_kernel", ns);
        if (tc_hnorm_launch(p->dP, p->hp, h, ns)) return fail(CCSD_ERR_CUDA, "tc_hnorm launch failed");
        PROF_END(p, ns);
      }
      ZnormArgs z; memset(&z, 0, sizeof z);
      z.flags = p->flags; z.noise = nr2 ? nr2 + (size_t)slot * sr : nullptr; z.zmask = p->zmask; z.norm_part = p->norm_part;
      z.slot = slot; z.nz = nz;
      PROF_BEGIN(p, "znorm_kernel", ns);
      CCSD_LAUNCH(znorm_kernel, dim3(p->hp.ntile_r2, d.B, 1), 256, (size_t)(40 + ((d.E + 3) & ~3) + APPLY_TN / 4 + 8) * 4, ns, p->dP, z);
      PROF_END(p, ns);
      p->launches += 2;
      if (p->side) {
        if (cudaEventRecord(p->ev_join, p->side) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream join failed");
        p->side_pending = true;
      }
      return dev_check("rank-2 norms");
    }
#endif
    return apply_pass(MODE_NORM, slot);
  };
  auto join_side = [&]() -> int {
#ifndef CCSD_EMU
    if (p->side_pending) {
      p->side_pending = false;
      if (cudaStreamWaitEvent((cudaStream_t)stream, p->ev_join, 0) != cudaSuccess) return fail(CCSD_ERR_CUDA, "stream join failed");
    }
#endif
    return 0;
  };
  // x / adj networks (+ the Gram pre-pass their hodge branch and the rank-2 network need)
  auto xa_phase = [&](int mode, int slot, int which, const float *xin, const float *adjin) -> int {
    XaArgs a; memset(&a, 0, sizeof a);
    a.x = xin; a.adj = adjin; a.flags = p->flags; a.P0 = p->P0; a.P1 = p->P1; a.r2 = p->r2;
    a.mode = mode; a.which = which; a.slot = slot; a.denoise = d.denoise; a.nz = nz;
    a.norm_part = p->norm_part;
    a.noise_x = nx ? nx + (size_t)slot * sx : nullptr;
    a.noise_adj = nadj ? nadj + (size_t)slot * sa : nullptr;
    if (mode == MODE_SCORE) { a.out_x = p->sx; a.out_adj = p->sadj; }
    else { a.out_x = p->x; a.out_adj = p->adj; a.mean_x = p->mx; a.mean_adj = p->madj; a.traj_x = tx; a.traj_adj = ta; }
    return launch_xa(p, a, stream);
  };
  auto score_phase = [&](int mode, int slot, int r2_mode) -> int {
    if (d.is_cc)
      if (int r = launch_rank2_pre(p, p->r2, p->adj, p->flags, stream)) return r;
    bool normed = false;
#ifndef CCSD_EMU
    if (d.is_cc && r2_mode == MODE_NORM && p->use_hnorm) {   // the norm kernels only need the Gram products: start them beside the x / adj pipeline
      if (int r = norm_pass(slot)) return r;
      normed = true;
    }
#endif
    if (int r = xa_phase(mode, slot, 3, p->x, p->adj)) return r;
    if (d.is_cc && !normed) return r2_mode == MODE_NORM ? norm_pass(slot) : apply_pass(r2_mode, slot);
    return 0;
  };
  // Langevin / S4 step sizes of the objects in obj_mask from the batch means of their norm partials
  auto coef_phase = [&](int obj_mask) -> int {
    if (int r = join_side()) return r;
    CoefArgs c; c.norm_part = p->norm_part; c.coef = p->coef; c.step = step; c.s4 = s4; c.obj_mask = obj_mask; c.sd = nz.sd;
    PROF_BEGIN(p, "coef_kernel", stream);
    CCSD_LAUNCH(coef_kernel, dim3(d.is_cc ? 3 : 2, 1, 1), d.B >= 2048 ? 1024 : 256, 64 * 4, stream, p->dP, c);
    PROF_END(p, stream);
    p->launches++;
    return dev_check("coef_kernel");
  };
  // the corrector update of objects [obj0, obj0 + nobj) with noise draw `slot` (PC: x and / or adj here, the rank-2
  // state in its own CORR apply pass) or the whole S4 chain for every object
  auto update_phase = [&](int obj0, int nobj, int slot) -> int {
    UpdateArgs u; memset(&u, 0, sizeof u);
    u.flags = p->flags; u.x = p->x; u.adj = p->adj; u.r2 = p->r2; u.sx = p->sx; u.sadj = p->sadj; u.sr2 = p->sr2;
    u.coef = p->coef; u.mx = p->mx; u.madj = p->madj; u.mr2 = p->mr2; u.nx = nx; u.nadj = nadj; u.nr2 = nr2;
    u.tx = tx; u.tadj = ta; u.tr2 = tr; u.s4 = s4; u.denoise = d.denoise; u.write_mean_r2 = write_mean; u.nz = nz;
    u.obj0 = obj0; u.slot0 = slot;
    const size_t units = obj0 + nobj == 3 ? (size_t)d.B * d.E * (p->hp.Kp / 4) : (obj0 + nobj == 2 ? sa : sx);
    PROF_BEGIN(p, "update_kernel", stream);
    CCSD_LAUNCH(update_kernel, dim3(grid_for(units), nobj, 1), 256, 0, stream, p->dP, u);
    PROF_END(p, stream);
    p->launches++;
    return dev_check("update phase");
  };

  if (s4) {
    if (int r = score_phase(MODE_SCORE, 0, MODE_SCORE)) return r;
    if (int r = coef_phase(7)) return r;
    return update_phase(0, d.is_cc ? 3 : 2, 0);
  }
  int slot = 0;
  if (d.use_corrector) {
    const int n = d.n_lang_steps;
    // corrector: the rank-2 score is never written to HBM -- a NORM pass produces the norms, and after
    // the step sizes are known a CORR pass recomputes the score and updates the state in place
    if (int r = score_phase(MODE_SCORE, 0, MODE_NORM)) return r;
    if (int r = coef_phase(7)) return r;
    if (n > 1) {
      // LangevinCorrector.update_fn loops n_steps times PER OBJECT, every object's loop seeing the other objects
      // at their PRE-corrector values (solver.py:692-701, 760-785; the three correctors are called with the same
      // x, adj, rank2, :1123-1140).  The mean buffers are free during the corrector: they keep x0 / adj0.
      if (int r = dev_copy(p->mx, p->x, sx * 4, stream)) return r;
      if (int r = dev_copy(p->madj, p->adj, sa * 4, stream)) return r;
    }
    if (int r = update_phase(0, 2, 0)) return r;
    for (int s = 1; s < n; ++s) {   // x: score_x(x_s, adj0)
      if (int r = xa_phase(MODE_SCORE, s, 1, p->x, p->madj)) return r;
      if (int r = coef_phase(1)) return r;
      if (int r = update_phase(0, 1, s)) return r;
    }
    for (int s = 1; s < n; ++s) {   // adj: score_adj(x0, adj_s, rank2_0); the Gram products of rank2_0 are still in H / P0
      if (int r = xa_phase(MODE_SCORE, s, 2, p->mx, p->adj)) return r;
      if (int r = coef_phase(2)) return r;
      if (int r = update_phase(1, 1, s)) return r;
    }
    if (d.is_cc) {
      if (int r = apply_pass(MODE_CORR, 0)) return r;
      for (int s = 1; s < n; ++s) {   // rank2: score_rank2(rank2_s) (ScoreNetworkF ignores x and adj)
        if (int r = launch_rank2_pre(p, p->r2, p->adj, p->flags, stream)) return r;
        if (int r = norm_pass(s)) return r;
        if (int r = coef_phase(4)) return r;
        if (int r = apply_pass(MODE_CORR, s)) return r;
      }
    }
    slot = n;
  }
  return score_phase(MODE_PRED, slot, MODE_PRED);
}

int ccsd_plan_step(ccsd_plan_t *p, int step, const float *nx, const float *nadj, const float *nr2, void *stream) {
  if (!p) return fail(CCSD_ERR_INVALID, "null plan");
  if (!p->inited) return fail(CCSD_ERR_STATE, "ccsd_plan_init must be called before stepping");
  if (step < 0 || step >= p->hp.d.n_diff_steps) return fail(CCSD_ERR_INVALID, "step out of range");
  return do_step(p, step, nx, nadj, nr2, 1, stream);
}

int ccsd_plan_run(ccsd_plan_t *p, int step_begin, int step_end, void *stream) {
  if (!p) return fail(CCSD_ERR_INVALID, "null plan");
  if (!p->inited) return fail(CCSD_ERR_STATE, "ccsd_plan_init must be called before running");
  if (step_begin < 0 || step_end > p->hp.d.n_diff_steps || step_begin > step_end) return fail(CCSD_ERR_INVALID, "step range out of bounds");
  int s = step_begin;
#ifndef CCSD_EMU
  // Latency path of small graph-only batches: a step is ~40 dependent launches of 8-30 us kernels, and the gaps between them
  // are a sixth of the step.  One step is captured as a CUDA graph whose kernels read the step index (schedule row, Philox
  // draw id) and the diff_traj slots from device memory (StepDev); step_advance_kernel closes the step; the graph is replayed
  // for every step but the first (eager: it also sets every kernel's shared-memory attribute) and the last (which returns means).
  if (p->use_graph && !p->profiling && step_end - step_begin >= 32) {   // (capture + instantiation cost ~1-15 ms: long runs only)
    const ccsd_plan_desc_t &d = p->hp.d;
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = do_step(p, s, nullptr, nullptr, nullptr, 0, stream)) return r;
    ++s;
    step_set_kernel<<<1, 1, 0, st>>>(p->sd_dev, s, p->traj_x ? p->traj_x + (size_t)s * d.N * d.F : nullptr,
                                      p->traj_adj ? p->traj_adj + (size_t)s * d.N * d.N : nullptr);
    drop_graph(p);
    if (!p->ev_graph && cudaEventCreateWithFlags(&p->ev_graph, cudaEventDisableTiming) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaEventCreate failed");
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const int64_t l0 = p->launches;
    if (!p->gs && cudaStreamCreateWithFlags(&p->gs, cudaStreamNonBlocking) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaStreamCreate failed");
    if (cudaStreamBeginCapture(p->gs, cudaStreamCaptureModeRelaxed) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaStreamBeginCapture failed");
    p->graph_capture = true;
    int rc = do_step(p, s, nullptr, nullptr, nullptr, 0, (void *)p->gs);
    p->graph_capture = false;
    step_advance_kernel<<<1, 1, 0, p->gs>>>(p->sd_dev, d.N * d.F, d.N * d.N);
    const cudaError_t ce = cudaStreamEndCapture(p->gs, &graph);
    const int64_t per_step = p->launches - l0 + 1;
    p->launches = l0;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || !graph) return fail(CCSD_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); return fail(CCSD_ERR_CUDA, "cudaGraphInstantiate failed"); }
    p->ggraph = graph; p->gexec = exec;   // destroyed by the next run / ccsd_plan_destroy, after ev_graph
    for (; s < step_end - 1; ++s) {
      if (cudaGraphLaunch(exec, st) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaGraphLaunch failed");
      p->launches += per_step;
    }
    if (cudaEventRecord(p->ev_graph, st) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaEventRecord failed");
  }
#endif
  for (; s < step_end; ++s)
    if (int r = do_step(p, s, nullptr, nullptr, nullptr, s == step_end - 1, stream)) return r;
  return 0;
}

int ccsd_plan_read(ccsd_plan_t *p, int want_mean, float *ox, float *oadj, float *or2, void *stream) {
  if (!p || !p->inited) return fail(CCSD_ERR_STATE, "plan not initialised");
  const ccsd_plan_desc_t &d = p->hp.d;
  if (ox) if (int r = dev_copy(ox, want_mean ? p->mx : p->x, (size_t)d.B * d.N * d.F * 4, stream)) return r;
  if (oadj) if (int r = dev_copy(oadj, want_mean ? p->madj : p->adj, (size_t)d.B * d.N * d.N * 4, stream)) return r;
  if (or2 && d.is_cc) if (int r = dev_copy(or2, want_mean ? p->mr2 : p->r2, (size_t)d.B * d.E * d.K * 4, stream)) return r;
  return 0;
}

int ccsd_score_eval(ccsd_plan_t *p, int which, const float *x, const float *adj, const float *r2, const float *flags,
                    float *out, void *stream) {
  if (!p || !x || !adj || !flags || !out) return fail(CCSD_ERR_INVALID, "null argument");
  if (!p->bound) return fail(CCSD_ERR_STATE, "ccsd_plan_bind must be called first");
  const ccsd_plan_desc_t &d = p->hp.d;
  if (d.is_cc && which != CCSD_NET_X && !r2) return fail(CCSD_ERR_INVALID, "rank2 is required for CC plans");
  if (which == CCSD_NET_RANK2 && !d.is_cc) return fail(CCSD_ERR_INVALID, "graph plans have no rank-2 network");
  if (which < 0 || which > 2 || !(d.nets & (1 << which))) return fail(CCSD_ERR_INVALID, "that network is not part of this plan");
  p->gimg_fresh = false;
  if (which == CCSD_NET_X || which == CCSD_NET_ADJ) {
    if (which == CCSD_NET_ADJ && d.is_cc)
      if (int r = launch_rank2_pre(p, r2, adj, flags, stream)) return r;
#ifndef CCSD_EMU
    if (p->use_tc_xfin && which == CCSD_NET_X)
      if (tc_xfin_prep(p->dP, p->txf, p->ximg, stream)) return fail(CCSD_ERR_CUDA, "tc_xfin_prep launch failed");
#endif
    XaArgs a; memset(&a, 0, sizeof a);
    a.x = x; a.adj = adj; a.flags = flags; a.P0 = p->P0; a.P1 = p->P1; a.r2 = r2; a.mode = MODE_EVAL;
    a.which = which == CCSD_NET_X ? 1 : 2;
    a.out_x = out; a.out_adj = out;
    return launch_xa(p, a, stream);
  }
  if (which == CCSD_NET_RANK2) {
    if (int r = launch_rank2_pre(p, r2, adj, flags, stream)) return r;
    CCSD_LAUNCH(zmask_kernel, dim3(grid_for(d.B), 1, 1), 256, 0, stream, flags, p->zmask_eval, d.B, d.N);
    ApplyArgs q; memset(&q, 0, sizeof q);
    q.r2 = r2; q.H = p->H; q.flags = flags; q.mode = MODE_EVAL; q.out = out; q.zmask = p->zmask_eval;
    launch_apply(p, q, stream);
    p->launches++;
    return dev_check("apply_kernel");
  }
  return fail(CCSD_ERR_INVALID, "unknown network id");
}

int ccsd_quantize(const float *in, uint8_t *out, size_t n, float thr, int mol, void *stream) {
  if (!in || !out) return fail(CCSD_ERR_INVALID, "null argument");
  if (n == 0) return 0;
  CCSD_LAUNCH(quantize_kernel, dim3(grid_for(n), 1, 1), 256, 0, stream, in, out, n, thr, mol);
  return dev_check("quantize_kernel");
}

int ccsd_mol_onehot(const float *x, const float *adj, int64_t *x_out, int64_t *adj_out, int B, int N, int F, void *stream) {
  if (!x || !adj || !x_out || !adj_out) return fail(CCSD_ERR_INVALID, "null argument");
  if (B < 1 || N < 1 || F < 1) return fail(CCSD_ERR_INVALID, "B, N, F must be positive");
  CCSD_LAUNCH(mol_onehot_kernel, dim3(grid_for((size_t)B * N * N), 1, 1), 256, 0, stream, x, adj, (long long *)x_out,
              (long long *)adj_out, B, N, F);
  return dev_check("mol_onehot_kernel");
}

int ccsd_cc_cells(const float *r2, uint8_t *present, int32_t *row, float *label, int B, int E, int K, void *stream) {
  if (!r2 || !present || !row || !label) return fail(CCSD_ERR_INVALID, "null argument");
  if (B < 1 || E < 1 || K < 1) return fail(CCSD_ERR_INVALID, "B, E, K must be positive");
  CCSD_LAUNCH(cc_cells_kernel, dim3(grid_for((size_t)B * K), 1, 1), 256, 0, stream, r2, present, (int *)row, label, B, E, K);
  return dev_check("cc_cells_kernel");
}

int64_t ccsd_plan_launch_count(const ccsd_plan_t *p) { return p ? p->launches : 0; }

int ccsd_debug_apply_trace(ccsd_plan_t *p, long long *trace_dev) {
  if (!p) return fail(CCSD_ERR_INVALID, "null plan");
  p->trace = trace_dev;
  return 0;
}

int ccsd_plan_info(const ccsd_plan_t *p, int what) {
  if (!p) return -1;
  switch (what) {
    case 0: return imax(imax(imax(p->hp.xp.x_total, p->hp.xp.c_total), imax(p->hp.xp.f_total, p->hp.xp.h_total)), p->hp.xp.m_total) * 4;
    case 1: return p->hp.xp.Tc;
    case 2: return (int)p->apply_smem;
    case 3: return p->use_tc;
    case 4: return p->use_tc_apply;
    case 5: return p->hp.f_mode;
    case 6: return p->hp.xp.m_rows;
    case 12: return p->hp.PR0;
    case 13: return p->use_tc_fin;
#ifndef CCSD_EMU
    case 15: return p->use_tc_xfin;
    case 16: return p->use_hnorm;
    case 18: return p->use_tc_big;
    case 19: return p->use_graph;
    case 17: { int n = 0; for (int l = 0; l < p->hp.d.neta.num_layers; ++l) n += p->use_tc_edge[l]; return n; }
    case 14: { int n = 0; for (int l = 0; l < p->hp.d.neta.num_layers; ++l) n += p->use_tc_attn[l]; return n; }   // layers on the tcgen05 attention kernel
#endif
    case 7: return p->hp.xp.x_total * 4;
    case 8: return p->hp.xp.c_total * 4;
    case 9: return p->hp.xp.f_total * 4;
    case 10: return p->hp.xp.h_total * 4;
    case 11: return p->hp.xp.m_total * 4;
    default: return -1;
  }
}

int ccsd_debug_gram(ccsd_plan_t *p, const float *r2, float *H_out, float *P0_out, int use_tc, void *stream) {
  if (!p || !r2 || !p->bound || !p->hp.d.is_cc) return fail(CCSD_ERR_INVALID, "ccsd_debug_gram: bound CC plan required");
  const ccsd_plan_desc_t &d = p->hp.d;
  const int saved = p->use_tc;
#ifndef CCSD_EMU
  const int saved_big = p->use_tc_big;
  const bool big = d.E > 192;
  if (use_tc && !(big ? tc_r2big_supported(d, p->hp.PR0) : tc_gram_supported(d.E, d.K, p->hp.PR0)))
    return fail(CCSD_ERR_UNSUPPORTED, "tensor-core Gram kernel does not cover this shape");
  if (use_tc && !big) tc_gram_prepare();
  p->use_tc_big = big ? use_tc : 0;
  p->use_tc = big ? 0 : use_tc;
#else
  p->use_tc = use_tc;
#endif
  int r = launch_rank2_pre(p, r2, p->adj, p->flags, stream);
  p->use_tc = saved;
#ifndef CCSD_EMU
  p->use_tc_big = saved_big;
#endif
  if (r) return r;
  if (H_out) {
    const size_t rows = (size_t)d.B * d.E, rb = (size_t)d.E * 4, pitch = (size_t)p->hp.Ep * 4;
#ifdef CCSD_EMU
    for (size_t r = 0; r < rows; ++r) memcpy((char *)H_out + r * rb, (const char *)p->H + r * pitch, rb);
#else
    if (cudaMemcpy2DAsync(H_out, rb, p->H, pitch, rb, rows, cudaMemcpyDefault, (cudaStream_t)stream) != cudaSuccess)
      return fail(CCSD_ERR_CUDA, "cudaMemcpy2DAsync failed");
#endif
  }
  if (P0_out && p->hp.PR0) if (int e = dev_copy(P0_out, p->P0, (size_t)d.B * d.E * p->hp.PR0 * 4, stream)) return e;
  return 0;
}

int ccsd_plan_set_profiling(ccsd_plan_t *p, int on) {
  if (!p) return fail(CCSD_ERR_INVALID, "null plan");
#ifndef CCSD_EMU
  for (auto &r : p->prof) { cudaEventDestroy((cudaEvent_t)r.e0); cudaEventDestroy((cudaEvent_t)r.e1); }
#endif
  p->prof.clear();
  p->profiling = on != 0;
  return 0;
}

int ccsd_plan_get_profile(ccsd_plan_t *p, int max_records, char *names, int name_stride, float *ms) {
  if (!p || !names || !ms) return fail(CCSD_ERR_INVALID, "null argument");
  int n = 0;
#ifndef CCSD_EMU
  for (auto &r : p->prof) {
    if (n >= max_records) break;
    if (cudaEventSynchronize((cudaEvent_t)r.e1) != cudaSuccess) return fail(CCSD_ERR_CUDA, "cudaEventSynchronize failed");
    float t = 0.f;
    cudaEventElapsedTime(&t, (cudaEvent_t)r.e0, (cudaEvent_t)r.e1);
    ms[n] = t;
    snprintf(names + (size_t)n * name_stride, name_stride, "%s", r.name);
    ++n;
  }
#endif
  return n;
}

}  // extern "C"
