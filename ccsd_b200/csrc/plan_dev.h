// plan_dev.h -- device-side view of a plan: descriptor + shared-memory layouts + table pointers.
#pragma once
#include "../../include/ccsd_b200.h"
#include "common.cuh"

namespace ccsd {

constexpr int SMALL_MAX = 32;     // widest layer of the per-entry / per-edge "small" MLPs
constexpr int GRAM_BM = 64, GRAM_BN = 64, GRAM_BK = 16;
constexpr int APPLY_TN = 64;      // cell columns per apply-kernel CTA

// Shared-memory layouts (float offsets) of the x / adj network pipeline (xa_pipe.cuh) and the per-graph
// strides of its global scratch.  Conventions in prims.cuh: feature-major buffers with ld = N4 (nodes) or
// ldp (node pairs i <= j), both multiples of 4.
struct XpLayout {
  int N4;              // N rounded up to 4
  int NT, ldp;         // node pairs i <= j: N(N+1)/2, rounded up to 4
  int Tx, Tc, Tf, Th, Tm;   // threads per CTA: x_net, attn_channel, attn_finish, hodge, afinal
  // x_net_kernel
  int x_flags, x_dvec, x_adj /*[c_init x ldp]*/, x_an /*[N x N4]*/, x_x0 /*[F x N4]*/, x_hcat /*[depth*nhid x N4]*/,
      x_ax, x_ha, x_hb /*[dhid x N4]*/, x_sx /*[F x N4]*/, x_red, x_total;
  // attn_channel_kernel
  int c_dvec, c_adj /*[ldp]*/, c_an, c_xin /*[kin x N4]*/, c_ax, c_q, c_k /*[adim x N4]*/, c_v /*[nhid x N4]*/,
      c_atp /*[heads x ldp]*/, c_w /*staged Q|K|V|W1 weights of the channel*/, c_mh /*[2 adim x N4] hidden layer of the conv == "MLP" Q / K networks*/, c_total;
  // attn_finish_kernel
  int f_flags, f_hs, f_hs2 /*[mc_hid x N4]*/, f_eha, f_ehb /*[hid x ldp]*/, f_total;
  // hodge_kernel (two hodge layers: hq, hk [c0 x E x ad0], h1 [c1 x E x lde], hdeg [c1 x E])
  int h_flags, h_hq, h_hk, h_h1, lde, h_hdeg, h_p1 /*[E x n1] folded projections*/, h_u /*[n1]*/,
      h_att1 /*[c1 x E] layer-1 attention diagonals*/, h_alpha /*[E]*/, h_part /*[threads] partial sums*/, h_total;
  // hodge_base_kernel (ScoreNetworkA_Base_CC; shares h_flags): u0 [c_init x E x hid_pad], edge flags [E], tri index [E]
  int hb_u0, hb_fe, hb_tri, hb_total;
  // afinal_kernel
  int m_flags, m_fa, m_fb /*[dhid x m_rows]*/, m_out /*[m_rows]*/, m_red, m_rows, m_nchunk, m_total;
  // global scratch, floats per graph
  int g_stack;         // [fdimA x ldp]  every adjacency channel the final MLP reads
  int g_att;           // [c_in x ldp]   symmetrised attention maps of the current layer
  int g_hmc;           // [c_in x mc_o1_max x N4]  per-channel shares of the node MLP's first Linear
  int g_x;             // [max(F, nhid) x N4]      node features (ping-pong)
  int mc_o1_max;
  // large-graph pipeline (big_pipe.cuh, max_node_num > 64): full N x Np planes in global memory
  int big;                      // 1: the plan runs the large-graph pipeline
  int big_Np, big_PS;           // N rounded up to 8; plane stride N * Np
  int big_nrc, big_nseg;        // row chunks per graph (BIG_RC rows); 64-column segments per row
  int big_total;                // floats of scratch per graph; offsets inside:
  int big_S /*[fdimA][N][Np]*/, big_ATT /*[c][N][Np]*/, big_Y /*[c][N][2 adp + nhp]*/,
      big_TQK /*[c][2 adp][Np]*/, big_TV /*[c nh][Np]*/, big_XF0, big_XF1 /*[kmax][Np]*/, big_DV /*[c][Np]*/,
      big_HC /*[fdimX][Np]*/;
  int big_T_xw;                 // threads per CTA of the xw / agg kernels (= items per CTA, <= 128)
  int big_sm_node, big_sm_edge, big_sm_fin, big_sm_xfin;   // dynamic shared memory (floats) of the MLP kernels
};

struct DevPlan {
  ccsd_plan_desc_t d;
  XpLayout xp;
  const float *W;                 // packed weights
  const ccsd_objcoef_t *sched;    // [n_diff_steps][3]
  const unsigned long long *cell_mask;  // [K] bit n set <=> node n in cell
  const int *edge_ij;             // [E][2]
  const int *tri_ij;              // [N(N+1)/2] (i << 8) | j of every node pair i <= j, row-major upper triangle
  int PR0, PR1;                   // columns of the Gram projection output P0 (row pitch) / of the P1 buffer
  int PR0h;                       // projection rows of hodge layer 0 proper (P0 columns [0, PR0h))
  int p1_fold;                    // hodge layer 0's value MLP is one Linear: its layer-1 projections are Gram
                                  // columns [PR0h, PR0) of F itself, rescaled per edge inside hodge_kernel
  int Kp;                         // K rounded up to 4 (Philox groups per rank-2 row = Kp/4)
  int Ep;                         // E rounded up to 4: row pitch of the H buffer [B][E][Ep]
  int ntile_r2;                   // apply-kernel column tiles per sample
  int ntile_adj;                  // afinal-kernel row chunks per sample (norm partial slots of the adjacency)
  int ntile_x;                    // norm partial slots of x (1 on the per-graph-tile path)
  int ntile_max;                  // stride of the per-object norm partials
  int f_mode;                     // ScoreNetworkF entry path: 0 generic, 1 affine fold, 2 <=8-wide unrolled, 3 <=4-wide x4 entries, 4 <=8-wide + 2-Linear final
  int f_nlin;                     // number of Linears staged for f_mode 2
  int gram_group;                 // tensor-core Gram kernel: samples per work unit (stacked rows <= 192, columns <= 256)
  int ap_group;                   // tensor-core apply kernel: samples per work group (min(8, 192 / E))
};

// modes of the score kernels
// EVAL: raw network output.  SCORE: scaled score to `out` + norm partials.  PRED: predictor update.
// NORM: norm partials only (nothing written).  CORR: Langevin correction with the step sizes in `coef`
// (rank-2 apply kernels only; x / adj are corrected by update_kernel).
enum { MODE_EVAL = 0, MODE_SCORE = 1, MODE_PRED = 2, MODE_NORM = 3, MODE_CORR = 4 };

// Device-resident step state of a CUDA-graph replay (ccsd_plan_run): the captured step reads the diffusion step index and the
// diff_traj destinations from here instead of from kernel arguments, and step_advance_kernel moves them on at its end, so ONE
// captured step serves every replay.  Eager launches pass sd = nullptr and use the argument values.
struct StepDev {
  int step, pad;
  float *tx, *ta;   // diff_traj destinations of this step (sample 0 of the shard) or nullptr
};

struct NoiseCtx {
  unsigned long long seed;
  long long sample_offset;
  int step;
  const StepDev *sd;
};
#if defined(__CUDACC__) || defined(CCSD_EMU)
__host__ __device__ inline int nz_step(const NoiseCtx &nz) { return nz.sd ? nz.sd->step : nz.step; }
#endif

}  // namespace ccsd
