/* ccsd_b200.h -- C ABI of the B200-native CCSD reverse-SDE sampler.
 *
 * The reference (AdrienC21/CCSD, pure Python/PyTorch) has no FFI; its boundary for this path is
 * the Python callable contract of ccsd/src/solver.py:856-875 (get_pc_sampler) and :1179-1198
 * (S4_solver), called from ccsd/src/utils/loader.py:337-458 (load_sampling_fn).  This header is
 * what a thin binding (ctypes stub in INTEGRATION.md; ccsd_b200/_native.py in this repo) binds to
 * replace that path.  Plain pointers and sizes only; no torch types.
 *
 * Ownership: every device pointer (weights, workspace, state, outputs, noise) is allocated and
 * owned by the caller (PyTorch); the library never allocates or frees device memory.  A plan owns
 * host metadata only.  All launches are asynchronous on the caller's stream.  A plan is used by
 * one host thread at a time; distinct plans are independent (no global mutable state except the
 * thread-local last-error string).
 *
 * Errors: 0 = OK; negative codes below; ccsd_last_error() gives the message.  Never aborts.
 */
#ifndef CCSD_B200_H
#define CCSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCSD_OK 0
#define CCSD_ERR_INVALID (-1)     /* bad argument / shape mismatch            -> ValueError          */
#define CCSD_ERR_UNSUPPORTED (-2) /* topology outside what the kernels cover   -> NotImplementedError */
#define CCSD_ERR_CUDA (-3)        /* CUDA runtime failure                      -> RuntimeError        */
#define CCSD_ERR_STATE (-4)       /* call order (e.g. step before bind/init)   -> RuntimeError        */

#define CCSD_MAX_LAYERS 8 /* attention layers (ScoreNetworkA.num_layers), GCN depth            */
#define CCSD_MAX_CH 8     /* channels per attention layer (c_init / c_hid / c_final)           */
#define CCSD_MAX_MLP 4    /* linears per MLP                                                    */
#define CCSD_MAX_HODGE_LAYERS 2
#define CCSD_MAX_F_LAYERS 4

/* samplers (solver.py:856 get_pc_sampler, :1179 S4_solver) */
#define CCSD_SAMPLER_PC 0
#define CCSD_SAMPLER_S4 1
/* which network for ccsd_score_eval */
#define CCSD_NET_X 0
#define CCSD_NET_ADJ 1
#define CCSD_NET_RANK2 2

/* One torch.nn.Linear chain (ccsd/src/models/layers.py:161-275, use_bn=False).  Weights live in
 * the blob TRANSPOSED to (in, out_pad) row-major with out_pad = round_up(out, 8) zero padded;
 * bias has out_pad entries.  w/b are float offsets into the blob (multiples of 4). */
typedef struct {
  int32_t nl;                  /* number of linears (>=1)                       */
  int32_t din, dhid, dout;     /* layer i maps d_i -> d_{i+1}                   */
  int32_t w[CCSD_MAX_MLP];
  int32_t b[CCSD_MAX_MLP];
} ccsd_mlp_t;

/* DenseGCNConv (layers.py:57-158): weight (in, out_pad) row-major, bias out_pad. */
typedef struct {
  int32_t din, dout;
  int32_t w, b;
} ccsd_gcn_t;

/* AttentionLayer (attention.py:186-304) with conv == "GCN" */
typedef struct {
  int32_t c_in, c_out, conv_in, attn_dim, conv_out;
  ccsd_gcn_t q[CCSD_MAX_CH], k[CCSD_MAX_CH], v[CCSD_MAX_CH];
  /* Optional (dout = 0 when absent): the value convolution of channel c folded with that channel's slice of
   * multi_channel's first Linear, W_vw = W_v . W1[c*conv_out:(c+1)*conv_out, :] ((conv_in, o1_pad) row-major,
   * o1 = width of that Linear) and b_vw = b_v . W1[...] -- the node MLP is linear in the channel concat
   * (attention.py:292), so the tensor-core attention kernel never materialises V.  Computed by the packer. */
  ccsd_gcn_t vw[CCSD_MAX_CH];
  /* conv == "MLP" (attention.py:170-180): Q and K are 2-layer tanh MLPs of x alone (in -> 2 attn_dim -> attn_dim), V stays
   * a DenseGCNConv; q[] / k[] are unused then */
  int32_t conv_mlp;
  ccsd_mlp_t qm[CCSD_MAX_CH], km[CCSD_MAX_CH];
  ccsd_mlp_t mlp;           /* per-edge MLP, in = 2*c_in                      */
  ccsd_mlp_t multi_channel; /* node MLP, in = c_in*conv_out                   */
} ccsd_attn_layer_t;

/* ScoreNetworkX (ScoreNetwork_X.py:22-133), or ScoreNetworkX_GMH (ScoreNetwork_X.py:156-341) when gmh = 1: `depth`
 * AttentionLayers on (x, A^1 .. A^gmh_c_init) instead of GCN layers (nhid = their conv_out), tanh of every layer's node
 * output, then the same concat + 3-layer MLP */
typedef struct {
  int32_t nfeat, depth, nhid, fdim;
  ccsd_gcn_t gcn[CCSD_MAX_LAYERS];
  ccsd_mlp_t fin;
  int32_t gmh, gmh_c_init, gmh_heads;
  ccsd_attn_layer_t glayer[CCSD_MAX_LAYERS];
} ccsd_netx_t;

/* HodgeAdjAttentionLayer (hodge_attention.py:185-325), conv == "HCN".  The q/k weights
 * (K x attn_dim) are stored TRANSPOSED as rows of length K ("projection rows"): row index
 * proj_row + (ch*2 + {0:q,1:k})*attn_dim + d in a [n_proj_rows x K_pad] matrix at blob offset
 * ccsd_neta_t.proj_w, so that rank2 @ W is extra columns of the F F^T Gram product. */
typedef struct {
  int32_t c_in, c_out, attn_dim;
  int32_t proj_row;             /* first projection row of this layer            */
  int32_t bq[CCSD_MAX_CH], bk[CCSD_MAX_CH]; /* bias offsets (attn_dim each)     */
  ccsd_mlp_t mlp_attention, mlp_value;
} ccsd_hodge_layer_t;

/* HodgeBaselineLayer (hodge_layers.py:287-416) of ScoreNetworkA_Base_CC.  BaselineBlock c (hodge_layers.py:202-284)
 * is a row-wise MLP E -> hid -> E on the Hodge adjacency: its first Linear is stored (E, hid_pad) row-major
 * (input-major, like every other Linear), its second Linear (hid -> E) as (E, hid_pad) row-major BY OUTPUT ROW,
 * bias b2 with E entries. */
typedef struct {
  int32_t c_in, c_out, hid;
  int32_t w1[CCSD_MAX_CH], b1[CCSD_MAX_CH], w2[CCSD_MAX_CH], b2[CCSD_MAX_CH];
  ccsd_mlp_t mlp_hodge;
} ccsd_hbase_layer_t;

/* ScoreNetworkA (ScoreNetwork_A.py:370-541) / ScoreNetworkA_CC (ScoreNetwork_A_CC.py:24-332) /
 * ScoreNetworkA_Base_CC (ScoreNetwork_A_Base_CC.py:24-323: is_cc = 1, base_cc = 1, hbase[] instead of hodge[]) */
typedef struct {
  int32_t is_cc;
  int32_t num_layers, c_init, num_heads, fdim;
  ccsd_attn_layer_t layer[CCSD_MAX_LAYERS];
  int32_t num_layers_h, num_heads_h;
  int32_t n_proj_rows[CCSD_MAX_HODGE_LAYERS]; /* projection rows per hodge layer */
  int32_t proj_w;                              /* blob offset of [sum rows x K_pad] */
  ccsd_hodge_layer_t hodge[CCSD_MAX_HODGE_LAYERS];
  int32_t base_cc;
  ccsd_hbase_layer_t hbase[CCSD_MAX_HODGE_LAYERS];
  ccsd_mlp_t fin;
} ccsd_neta_t;

/* ScoreNetworkF (ScoreNetwork_F.py:24-217) */
typedef struct {
  int32_t num_layers, cnum, fdim, use_hodge_mask;
  ccsd_mlp_t layer[CCSD_MAX_F_LAYERS]; /* HodgeNetworkLayer MLPs (hodge_layers.py:17-92) */
  ccsd_mlp_t fin;
  /* When every MLP of the network is a single Linear (num_linears == 1 and num_layers_mlp == 1, as in
   * the shipped community_small / ego_small / grid_small / QM9 CC checkpoints) there is no activation
   * anywhere and, the masks being {0,1}, the score is m * (aff[0]*f + aff[1]*(H f) + aff[2]).  The
   * packer folds the chain on the host (float64) and sets affine = 1. */
  int32_t affine;
  float aff[3];
} ccsd_netf_t;

/* Per-step, per-object scalars, computed on the host with the reference's own torch fp32
 * expressions (ccsd_b200/schedule.py; sde.py:345-786, solver.py:684-688, 752-756, 1291-1348). */
typedef struct {
  float score_scale; /* -1/std (VP, subVP) or 1 (VE): losses.py:67-70,159-162                  */
  float lg_alpha;    /* Langevin alpha: sde.alphas[timestep] (VP/subVP) or 1 (VE)              */
  float pa, pb, pc;  /* predictor: mean = pa*obj + pb*score ; new = mean + pc*z                */
  float s4_alpha;    /* S4 correction alpha (VPSDE only, else 1)                               */
  float s4_m1, s4_s1;/* first half transition: obj = m1*obj + s1*z                             */
  float s4_sd;       /* obj += sd*score,  sd = -g(t)^2 * dt                                    */
  float s4_m2, s4_s2;/* second half transition: mean = m2*obj ; new = mean + s2*z              */
  float pad_;
} ccsd_objcoef_t;

typedef struct {
  int32_t B, N, F;         /* batch on this device, max_node_num, max_feat_num                */
  int32_t is_cc, E, K;     /* rank-2 dims (cc_utils.py:268-283), 0 when !is_cc                */
  int32_t d_min, d_max;
  int32_t sampler;         /* CCSD_SAMPLER_*                                                  */
  int32_t use_corrector;   /* PC: 1 = Langevin, 0 = None (solver.py:833-853)                  */
  int32_t n_lang_steps;    /* Langevin n_steps (only 1 is implemented)                        */
  int32_t denoise;         /* return last means instead of last state                         */
  int32_t n_diff_steps;    /* sde_adj.N (solver.py:969,1119)                                  */
  int32_t nets;            /* bit0 ScoreNetworkX, bit1 ScoreNetworkA(_CC), bit2 ScoreNetworkF present;
                              stepping needs all of them, ccsd_score_eval only the one asked for */
  float snr, scale_eps;
  ccsd_netx_t netx;
  ccsd_neta_t neta;
  ccsd_netf_t netf;
} ccsd_plan_desc_t;

typedef struct ccsd_plan ccsd_plan_t;

/* sizeof checks for bindings */
int ccsd_plan_desc_size(void);
int ccsd_objcoef_size(void);

/* Validates the descriptor, copies it and the schedule (n_diff_steps x 3 ccsd_objcoef_t, order
 * x, adj, rank2; host memory).  weights_dev: packed fp32 blob on the device (n_weights floats). */
int ccsd_plan_create(const ccsd_plan_desc_t *desc, const ccsd_objcoef_t *schedule_host,
                     const float *weights_dev, size_t n_weights, ccsd_plan_t **out);
void ccsd_plan_destroy(ccsd_plan_t *plan);

/* Device scratch the caller must provide (state, scores, Gram matrices, tables). */
size_t ccsd_plan_workspace_bytes(const ccsd_plan_t *plan);
/* Binds the workspace and uploads the plan tables into it (async on stream). */
int ccsd_plan_bind(ccsd_plan_t *plan, void *workspace_dev, size_t bytes, void *stream);

/* Optional diff_traj recording (solver.py:987-995, 1149-1165): buffers of n_diff_steps x
 * (N*F | N*N | E*K) floats receiving sample `traj_sample`'s (mean if denoise else state) every step. */
int ccsd_plan_set_traj(ccsd_plan_t *plan, float *traj_x, float *traj_adj, float *traj_rank2);

/* Prior sampling + masking (solver.py:963-968, 1111-1118).  flags_dev: B x N fp32 {0,1}.
 * prior_*: RAW standard normals (B x N x F, B x N x N, B x E x K) or NULL for Philox keyed by
 * (seed, sample_offset + b, ...).  */
int ccsd_plan_init(ccsd_plan_t *plan, const float *flags_dev, const float *prior_x,
                   const float *prior_adj, const float *prior_rank2, uint64_t seed,
                   int64_t sample_offset, void *stream);

/* One sampler iteration i (solver.py:973-995 / 1123-1165 / 1280-1363 / 1424-1552).  noise_*:
 * RAW normals for this step, [n_draws x B x ...] per object in reference draw order (PC: corrector
 * then predictor; S4: correction, first half, second half), or NULL for Philox. */
int ccsd_plan_step(ccsd_plan_t *plan, int step, const float *noise_x, const float *noise_adj,
                   const float *noise_rank2, void *stream);

/* Steps [step_begin, step_end) with Philox noise. */
int ccsd_plan_run(ccsd_plan_t *plan, int step_begin, int step_end, void *stream);

/* Copies (last means if want_mean else current state) to caller buffers (any may be NULL). */
int ccsd_plan_read(ccsd_plan_t *plan, int want_mean, float *out_x, float *out_adj,
                   float *out_rank2, void *stream);

/* Parity seam: raw model output model(x, adj[, rank2], flags) of one network
 * (ScoreNetwork_X.py:102, ScoreNetwork_A.py:505, ScoreNetwork_A_CC.py:275, ScoreNetwork_F.py:175)
 * on caller-provided device inputs (B x ...).  rank2 may be NULL for graph plans. */
int ccsd_score_eval(ccsd_plan_t *plan, int which, const float *x, const float *adj,
                    const float *rank2, const float *flags, float *out, void *stream);

/* quantize (graph_utils.py:181-192): out[i] = in[i] < thr ? 0 : 1 (uint8); quantize_mol
 * (graph_utils.py:195-213): thresholds .5/1.5/2.5 -> {0,1,2,3}. */
int ccsd_quantize(const float *in_dev, uint8_t *out_dev, size_t n, float thr, int mol, void *stream);

/* Molecule post-processing of Sampler_mol.sample (ccsd/src/sampler.py:814-825) on the device, one pass:
 * adj_out [B,4,N,N] int64 = one_hot(quantize_mol(adj) - 1 with -1 -> 3) permuted to channels-first,
 * x_out [B,N,F+1] int64 = (x > 0.5) with the appended "no atom" column 1 - sum_f. */
int ccsd_mol_onehot(const float *x_dev, const float *adj_dev, int64_t *x_out_dev, int64_t *adj_out_dev, int B, int N, int F,
                    void *stream);

/* Batched core of cc_from_incidence (ccsd/src/utils/cc_utils.py:156-265, the rank-2 part :236-258): for every sample b
 * and candidate cell k of rank2 [B,E,K]: present[b,k] = any_e rank2[b,e,k] != 0, row[b,k] = argmax_e |rank2[b,e,k]|
 * (first maximum), label[b,k] = rank2[b,row,k].  The caller builds the complex from the present cells only. */
int ccsd_cc_cells(const float *rank2_dev, uint8_t *present_dev, int32_t *row_dev, float *label_dev, int B, int E, int K,
                  void *stream);

/* Number of kernel launches issued by this plan so far (bench.py's gpu_launches). */
int64_t ccsd_plan_launch_count(const ccsd_plan_t *plan);

/* Introspection for tests / DESIGN.md tables: what = 0 xa-kernel dynamic shared memory (bytes), 1 xa-kernel
 * threads per CTA, 2 fp32 apply-kernel shared memory, 3 / 4 whether the tcgen05 Gram / apply kernels are
 * selected, 5 ScoreNetworkF entry path (0 generic, 1 affine fold, 2 <= 8-wide unrolled, 3 <= 4-wide, four entries at a time), 6 final-MLP row chunk. */
int ccsd_plan_info(const ccsd_plan_t *plan, int what);

/* Per-kernel device timing for bench.py's roofline: when on, every launch of ccsd_plan_step /
 * ccsd_plan_run is bracketed by CUDA events on the launching stream.  ccsd_plan_get_profile waits
 * for them and returns the number of records written (name_stride bytes per name). */
int ccsd_plan_set_profiling(ccsd_plan_t *plan, int on);
int ccsd_plan_get_profile(ccsd_plan_t *plan, int max_records, char *names, int name_stride, float *ms);

/* Test seam: H = (F F^T)(1-I) and P0 = F Wp^T of `r2` ([B,E,K]) through the fp32 FMA kernel
 * (use_tc = 0) or the tcgen05 kernel (use_tc = 1); outputs [B,E,E] and [B,E,PR0] (either may be NULL). */
int ccsd_debug_gram(ccsd_plan_t *plan, const float *r2, float *H_out, float *P0_out, int use_tc, void *stream);

/* Debug: device buffer of 512 x 16 int64 that receives clock64 stamps of the pipeline roles of CTA 0 of every
 * sampler-step tensor-core apply pass (slots: 0 loader issue, 1 landed, 2 operand slot free, 3 full; 4 MMA
 * start, 5 MMA issued; 6-9 / 10-13 epilogue warp 0 / 8: waiting, accumulator ready, loaded, done); NULL = off. */
int ccsd_debug_apply_trace(ccsd_plan_t *plan, long long *trace_dev);

const char *ccsd_last_error(void);
const char *ccsd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CCSD_B200_H */
