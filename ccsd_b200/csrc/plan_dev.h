// plan_dev.h -- device-side view of a plan: descriptor + shared-memory layouts + table pointers.
#pragma once
#include "../../include/ccsd_b200.h"
#include "common.cuh"

namespace ccsd {

constexpr int XA_THREADS = 512;
constexpr int SMALL_MAX = 32;     // widest layer of the per-entry / per-edge "small" MLPs
constexpr int GRAM_BM = 64, GRAM_BN = 64, GRAM_BK = 16;
constexpr int APPLY_TN = 64;      // cell columns per apply-kernel CTA

// float offsets into the xa kernel's dynamic shared memory
struct XaLayout {
  int ldn, ldp;        // padded (odd) row counts: nodes, node pairs
  int flags, dvec;     // [N], [N]
  int hcat;            // [fdimX x ldn]  feature-major: x then every GCN layer output
  int an;              // [N x ldn]      normalised adjacency of the current channel
  int stack;           // [fdimA x ldp]  every adjacency channel the final MLP reads
  int xa, xb;          // [nhid x ldn]   node features ping-pong (A-net)
  int sx;              // [F x ldn]      ScoreNetworkX output (feature-major)
  int sadj;            // [N*N]          ScoreNetworkA output
  int red;             // [40]           reduction scratch
  int scratch;         // start of the phase-aliased region
  // scratch-relative offsets
  int xw, ldxw;        // [N x ldxw]     node-major x@W for q|k|v (also X-net x@W)
  int qn, ldq;         // [N x ldq]      node-major Q
  int kf;              // [adim x ldn]   feature-major K
  int vcat;            // [c_in*nhid x ldn]
  int att;             // [c_in x ldp]
  int hA, hB;          // hidden ping-pong for row MLPs (sized for the largest dhid x rows use)
  int fin_ld;          // padded row chunk of the final per-edge MLP
  // hodge (Lh == 2)
  int hq, hk;          // [c0 x E x ad0] each
  int h1, lde;         // [c1 x E x lde]
  int hdeg;            // [c1 x E]
  int fin_rows;        // row chunk of the final per-edge MLP
  int total;           // floats
};

struct DevPlan {
  ccsd_plan_desc_t d;
  XaLayout xa;
  const float *W;                 // packed weights
  const ccsd_objcoef_t *sched;    // [n_diff_steps][3]
  const unsigned long long *cell_mask;  // [K] bit n set <=> node n in cell
  const int *edge_ij;             // [E][2]
  int PR0, PR1;                   // projection rows of hodge layer 0 / 1
  int Kp;                         // K rounded up to 4 (Philox groups per rank-2 row = Kp/4)
  int ntile_r2;                   // apply-kernel column tiles per sample
  int ntile_max;                  // stride of the per-object norm partials
  int f_mode;                     // ScoreNetworkF entry path: 0 generic, 1 affine fold, 2 <=8-wide unrolled
  int f_nlin;                     // number of Linears staged for f_mode 2
};

// modes of the score kernels
enum { MODE_EVAL = 0, MODE_SCORE = 1, MODE_PRED = 2 };

struct NoiseCtx {
  unsigned long long seed;
  long long sample_offset;
  int step;
};

}  // namespace ccsd
