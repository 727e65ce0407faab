// plan_dev.h -- device-side view of a plan: descriptor + shared-memory layouts + table pointers.
#pragma once
#include "../../include/ccsd_b200.h"
#include "common.cuh"

namespace ccsd {

constexpr int XA_MAX_THREADS = 128;  // xa kernel: 64 or 128 threads per graph (XaLayout::T), several CTAs per SM
constexpr int XA_MIN_BLOCKS = 3;     // register budget: 65536 / (3 * 128) = 170 per thread
constexpr int SMALL_MAX = 32;     // widest layer of the per-entry / per-edge "small" MLPs
constexpr int GRAM_BM = 64, GRAM_BN = 64, GRAM_BK = 16;
constexpr int APPLY_TN = 64;      // cell columns per apply-kernel CTA

// float offsets into the xa kernel's dynamic shared memory (see xa_kernel.cuh / prims.cuh for the
// conventions: feature-major buffers with ld = N4 (nodes) or ldp (node pairs i <= j), both multiples of 4)
struct XaLayout {
  int T;               // threads per CTA
  int N4;              // N rounded up to 4
  int NT, ldp;         // node pairs i <= j: N(N+1)/2, rounded up to 4
  // persistent over the whole kernel
  int flags, dvec;     // [N4] each
  int pij;             // [ldp] ints: (i << 8) | j of every pair
  int an;              // [N x N4]       normalised adjacency of the current channel (symmetric)
  int x0;              // [F x N4]       input node features (also the predictor's base state)
  int xa, xb;          // [nhid x N4]    node features ping-pong (A-net)
  int sx;              // [F x N4]       ScoreNetworkX output
  int sadj;            // [ldp]          ScoreNetworkA output
  int red;             // [40]           reduction scratch
  int stack;           // [fdimA x ldp]  every adjacency channel the final MLP reads
  int scratch;         // start of the phase-aliased region
  // A-net attention layers (scratch-relative)
  int att;             // [c_in x ldp]   symmetrised attention per channel
  int ax;              // [kin x N4]     aggregated node features A x
  int hmc, hmc2;       // [mc_hid x N4]  multi_channel MLP hidden (accumulated channel by channel), ping-pong
  int q, k, v;         // [adim x N4] x2, [nhid x N4]
  int atp;             // [heads x ldp]  per-head attention partials
  int eh_a, eh_b;      // [hid x ldp]    per-edge MLP hidden ping-pong (aliases q/k/v/atp)
  // final per-edge MLP (scratch-relative)
  int fh_a, fh_b;      // [dhid x fin_rows]
  int fin_rows;        // row chunk (multiple of 4)
  // X-net (absolute; aliases stack channels >= 1 and the scratch: runs before the A-net)
  int xh_cat;          // [depth*nhid x N4] GCN layer outputs
  int xh_ax;           // [max din x N4]
  int xh_a, xh_b;      // [dhid x N4]    final MLP hidden ping-pong
  // hodge (Lh == 2, scratch-relative)
  int hq, hk;          // [c0 x E x ad0] each
  int h1, lde;         // [c1 x E x lde]
  int hdeg;            // [c1 x E]
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
  int Ep;                         // E rounded up to 4: row pitch of the H buffer [B][E][Ep]
  int ntile_r2;                   // apply-kernel column tiles per sample
  int ntile_max;                  // stride of the per-object norm partials
  int f_mode;                     // ScoreNetworkF entry path: 0 generic, 1 affine fold, 2 <=8-wide unrolled
  int f_nlin;                     // number of Linears staged for f_mode 2
};

// modes of the score kernels
// EVAL: raw network output.  SCORE: scaled score to `out` + norm partials.  PRED: predictor update.
// NORM: norm partials only (nothing written).  CORR: Langevin correction with the step sizes in `coef`
// (rank-2 apply kernels only; x / adj are corrected by update_kernel).
enum { MODE_EVAL = 0, MODE_SCORE = 1, MODE_PRED = 2, MODE_NORM = 3, MODE_CORR = 4 };

struct NoiseCtx {
  unsigned long long seed;
  long long sample_offset;
  int step;
};

}  // namespace ccsd
