// prims.cuh -- block-cooperative building blocks over shared memory (fp32 FMA path).
//
// Conventions
//   * "feature-major" buffers are buf[f*ld + row] with ld a multiple of 4 and every buffer starting
//     at a multiple of 4 floats, so that 4 consecutive rows of one feature are one 16-byte load.
//   * Rows are processed in groups of 4; rows past the valid count (up to the next multiple of 4)
//     may hold anything -- their results are computed and discarded, never stored.
//   * Work items are strided over the block (`for (it = threadIdx.x; it < items; it += blockDim.x)`)
//     and no primitive synchronises at its end unless its comment says so.
//   * Node pairs (i <= j) of an N-node graph are stored in upper-triangle row-major order
//     ("tri" index); every adjacency-shaped quantity of the score networks is symmetric
//     (ScoreNetwork_A.py:505-541: powers of a symmetric matrix, symmetrised attention, M + M^T).
#pragma once
#include "plan_dev.h"

namespace ccsd {

#ifdef CCSD_EMU
__device__ __forceinline__ float fast_tanh(float v) { return tanhf(v); }
__device__ __forceinline__ float fast_elu(float v) { return v > 0.f ? v : expm1f(v); }
#else
// tanh(v) = 1 - 2 / (exp(2v) + 1): two SFU ops; absolute error ~1e-7 (the libm path is ~25 instructions)
__device__ __forceinline__ float fast_tanh(float v) {
  const float e = __expf(2.0f * v);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
// elu: absolute error ~1e-7 for v < 0
__device__ __forceinline__ float fast_elu(float v) { return v > 0.f ? v : __expf(v) - 1.0f; }
#endif

__device__ __forceinline__ float act_fast(float v, int act) {
  if (act == ACT_ELU) return fast_elu(v);
  if (act == ACT_TANH) return fast_tanh(v);
  return v;
}

__device__ __forceinline__ int tri_index(int i, int j, int N) {  // i <= j
  return i * N - (i * (i - 1)) / 2 + (j - i);
}
__device__ __forceinline__ int tri_index_any(int i, int j, int N) { return i <= j ? tri_index(i, j, N) : tri_index(j, i, N); }

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// ---------------------------------------------------------------------------------------------
// One 4-row x 8-output register tile of a Linear:
//   acc[rr][j] = sum_k in(k, r0 + rr) * W[k*Opad + oc + j]
//   in(k, r) = k < K1 ? in1[k*ld1 + r] : in2[(k-K1)*ld2 + r]   (feature-major, r0 % 4 == 0)
// W in global memory (read-only path; shared by every CTA, so L1/L2 resident).  32 FMAs per three
// 16-byte loads; k is unrolled by two so that six loads are in flight per thread.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_fma(float acc[4][8], const float4 a, const float4 w0, const float4 w1) {
  const float av[4] = {a.x, a.y, a.z, a.w};
  const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[rr][j] += av[rr] * wv[j];
}

struct TileOps {   // operands of two consecutive k-steps
  float4 a0, a1, w00, w01, w10, w11;
};
// WSM: the weights were staged in shared memory (plain loads) instead of global memory (read-only path)
template <bool WSM>
__device__ __forceinline__ float4 ldw4(const float *p) {
  return WSM ? *reinterpret_cast<const float4 *>(p) : __ldg(reinterpret_cast<const float4 *>(p));
}
template <bool WSM>
__device__ __forceinline__ void tile_load2(TileOps &t, const float *in, int ld, const float *__restrict__ wp, int Opad) {
  t.a0 = ld4(in); t.a1 = ld4(in + ld);
  t.w00 = ldw4<WSM>(wp); t.w01 = ldw4<WSM>(wp + 4);
  t.w10 = ldw4<WSM>(wp + Opad); t.w11 = ldw4<WSM>(wp + Opad + 4);
}

// Software pipelined: the six loads of k-steps (k+2, k+3) are issued before the 64 FMAs of (k, k+1), so a
// weight load that misses L1 (the weights are shared by every CTA and live in L2) is covered by this warp's
// own FMAs plus the other resident warps'.
#ifndef CCSD_DENSE_PIPE
#define CCSD_DENSE_PIPE 1
#endif
template <bool WSM = false>
__device__ __forceinline__ void dense_tile(float acc[4][8], const float *in1, int ld1, int K1, const float *in2, int ld2,
                                           int K2, const float *__restrict__ W, int Opad, int r0, int oc) {
  const float *wp = W + oc;
#pragma unroll 1
  for (int seg = 0; seg < 2; ++seg) {
    const float *in = (seg ? in2 : in1) + r0;
    const int ld = seg ? ld2 : ld1, K = seg ? K2 : K1;
    const int npair = K >> 1;
#if CCSD_DENSE_PIPE
    if (npair > 0) {
      // two operand sets used alternately (no register copies): the loads of pair p+1 are in flight during
      // the 64 FMAs of pair p
      TileOps A, B;
      tile_load2<WSM>(A, in, ld, wp, Opad);
      int p = 1;
#pragma unroll 1
      for (; p + 1 < npair; p += 2) {
        tile_load2<WSM>(B, in + 2 * p * ld, ld, wp + 2 * p * Opad, Opad);
        tile_fma(acc, A.a0, A.w00, A.w01);
        tile_fma(acc, A.a1, A.w10, A.w11);
        tile_load2<WSM>(A, in + 2 * (p + 1) * ld, ld, wp + 2 * (p + 1) * Opad, Opad);
        tile_fma(acc, B.a0, B.w00, B.w01);
        tile_fma(acc, B.a1, B.w10, B.w11);
      }
      if (p < npair) {
        tile_load2<WSM>(B, in + 2 * p * ld, ld, wp + 2 * p * Opad, Opad);
        tile_fma(acc, A.a0, A.w00, A.w01);
        tile_fma(acc, A.a1, A.w10, A.w11);
        tile_fma(acc, B.a0, B.w00, B.w01);
        tile_fma(acc, B.a1, B.w10, B.w11);
      } else {
        tile_fma(acc, A.a0, A.w00, A.w01);
        tile_fma(acc, A.a1, A.w10, A.w11);
      }
    }
#else
    // lean variant (fewer registers -> more resident CTAs): no software pipelining, latency hidden by other warps
#pragma unroll 1
    for (int p = 0; p < npair; ++p) {
      TileOps cur;
      tile_load2<WSM>(cur, in + 2 * p * ld, ld, wp + 2 * p * Opad, Opad);
      tile_fma(acc, cur.a0, cur.w00, cur.w01);
      tile_fma(acc, cur.a1, cur.w10, cur.w11);
    }
#endif
    if (K & 1) {
      const int k = K - 1;
      const float4 a0 = ld4(in + k * ld);
      const float4 w00 = ldw4<WSM>(wp + k * Opad), w01 = ldw4<WSM>(wp + k * Opad + 4);
      tile_fma(acc, a0, w00, w01);
    }
    wp += K * Opad;
  }
}

// out(r, o) = act(bias[o] + sum_k in(k, r) W[k, o]),  r < R, o < O, stored at out[r*sro + o*soo].
// bias may be nullptr.  When `accum` the previous out(r, o) is added (and act / bias are the caller's
// business): used to fold a Linear over a channel-concatenated input channel by channel.
template <bool WSM = false>
__device__ __forceinline__ void dense_fm(const float *in1, int ld1, int K1, const float *in2, int ld2, int K2,
                                         const float *__restrict__ W, const float *__restrict__ bias, int O,
                                         float *out, int sro, int soo, int R, int act, bool accum = false,
                                         int tid0 = 0) {
  const int Opad = round_up(O, 8);
  const int nchunk = Opad >> 3, ngrp = (R + 3) >> 2;
  const int items = nchunk * ngrp;
  // tid0: the thread that takes item 0 (lets two primitives in one barrier window start on different warps)
  const int vt = ((int)threadIdx.x + (int)blockDim.x - tid0 % (int)blockDim.x) % (int)blockDim.x;
  for (int it = vt; it < items; it += blockDim.x) {
    const int chunk = it / ngrp, g = it - chunk * ngrp;
    const int oc = chunk << 3, r0 = g << 2;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bv = bias ? __ldg(bias + oc + j) : 0.f;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) acc[rr][j] = bv;
    }
    dense_tile<WSM>(acc, in1, ld1, K1, in2, ld2, K2, W, Opad, r0, oc);
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = r0 + rr;
      if (r < R) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (oc + j < O) {
            float *dst = out + r * sro + (oc + j) * soo;
            *dst = accum ? *dst + acc[rr][j] : act_fast(acc[rr][j], act);
          }
      }
    }
  }
}

// A whole MLP over R rows (feature-major in / hidden).  Hidden activations ping-pong between hA and hB
// ([dhid x ldh]).  __syncthreads() after every layer, including the last.
__device__ __forceinline__ void mlp_fm(const ccsd_mlp_t &m, const float *__restrict__ W, const float *in1, int ld1,
                                       int K1, const float *in2, int ld2, int K2, int R, float *hA, float *hB, int ldh,
                                       float *out, int sro, int soo, int hidden_act, int out_act) {
  const float *cur1 = in1, *cur2 = in2;
  int c_ld1 = ld1, c_K1 = K1, c_ld2 = ld2, c_K2 = K2;
  for (int l = 0; l < m.nl; ++l) {
    const bool last = (l == m.nl - 1);
    const int O = last ? m.dout : m.dhid;
    float *dst = last ? out : ((l & 1) ? hB : hA);
    dense_fm(cur1, c_ld1, c_K1, cur2, c_ld2, c_K2, W + m.w[l], W + m.b[l], O, dst, last ? sro : 1, last ? soo : ldh, R,
             last ? out_act : hidden_act);
    __syncthreads();
    cur1 = dst; c_ld1 = ldh; c_K1 = O; cur2 = nullptr; c_ld2 = 0; c_K2 = 0;
  }
}

// DenseGCNConv normalisation (layers.py:139-147) of one symmetric adjacency channel in tri storage:
// A^ = adj with unit diagonal, d = rowsum(A^).clamp(min=1)^-1/2, an[i][j] = d_i A^_ij d_j (full,
// symmetric, [N x N4]).  Two __syncthreads (dvec after the first, `an` after the second).
__device__ __forceinline__ void gcn_norm_tri(const float *adj_tri, int N, int N4, float *dvec, float *an) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += (j == i) ? 1.f : adj_tri[tri_index_any(i, j, N)];
    dvec[i] = 1.0f / sqrtf(fmaxf(s, 1.f));
  }
  __syncthreads();
  for (int p = threadIdx.x; p < N * N4; p += blockDim.x) {
    const int i = p / N4, j = p - i * N4;
    float v = 0.f;
    if (j < N) v = dvec[i] * ((i == j) ? 1.f : adj_tri[tri_index_any(i, j, N)]) * dvec[j];
    an[p] = v;
  }
  __syncthreads();
}

// ax(k, i) = sum_j an[i][j] * xin(k, j): the GCN aggregation applied BEFORE the feature transform
// ((A x) W = A (x W), layers.py:144-156; K_in <= K_out for every layer here so this order is cheaper
// and needs no node-major intermediate).  xin, ax feature-major [K x N4].  `an` symmetric.
__device__ __forceinline__ void gcn_aggregate_fm(const float *an, int N, int N4, const float *xin, int K, float *ax) {
  const int ngrp = N4 >> 2;
  for (int it = threadIdx.x; it < K * ngrp; it += blockDim.x) {
    const int k = it / ngrp, i0 = (it - k * ngrp) << 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *xr = xin + k * N4;
    for (int j = 0; j < N; ++j) {
      const float xv = xr[j];
      const float4 w = ld4(an + j * N4 + i0);   // an[j][i0..i0+3] = an[i0..i0+3][j]
      a0 += w.x * xv; a1 += w.y * xv; a2 += w.z * xv; a3 += w.w * xv;
    }
    float *dst = ax + k * N4 + i0;
    dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
  }
}

// Attention.forward score part (attention.py:111-130) for one channel.  Q, K feature-major
// [ad x N4].  Heads are chunks of ds = ad / heads features (torch.split: ceil(ad/ds) chunks).
// Item = (4x4 node block I <= J, head): both orientations q_i.k_j and q_j.k_i are accumulated in
// registers (32 FMAs per four 16-byte loads), tanh'ed, and the symmetrised value
// 0.5 (tanh(s_ij) + tanh(s_ji)) / n_heads goes to atp[head][tri(i, j)].  The caller sums atp over heads.
__device__ __forceinline__ void attn_scores_blk(const float *Q, const float *Kf, int N, int N4, int ad, int heads,
                                                float scale, float *atp, int ldp) {
  const int ds = ad / heads;
  const int nch = (ad + ds - 1) / ds;
  const float inv = 0.5f / (float)nch;
  const int nb = N4 >> 2, nblk = nb * (nb + 1) / 2;
  for (int it = threadIdx.x; it < nblk * nch; it += blockDim.x) {
    const int blk = it / nch, h = it - blk * nch;
    int I = 0, rem = blk;
    while (rem >= nb - I) { rem -= nb - I; ++I; }
    const int J = I + rem;
    const int i0 = I << 2, j0 = J << 2;
    float a[4][4], b[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) { a[u][v] = 0.f; b[u][v] = 0.f; }
    const int d0 = h * ds, d1 = (d0 + ds < ad) ? d0 + ds : ad;
    for (int dd = d0; dd < d1; ++dd) {
      const float4 qi = ld4(Q + dd * N4 + i0), kj = ld4(Kf + dd * N4 + j0);
      const float4 qj = ld4(Q + dd * N4 + j0), ki = ld4(Kf + dd * N4 + i0);
      const float qiv[4] = {qi.x, qi.y, qi.z, qi.w}, kjv[4] = {kj.x, kj.y, kj.z, kj.w};
      const float qjv[4] = {qj.x, qj.y, qj.z, qj.w}, kiv[4] = {ki.x, ki.y, ki.z, ki.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) { a[u][v] += qiv[u] * kjv[v]; b[u][v] += qjv[v] * kiv[u]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = i0 + u, j = j0 + v;
        if (i <= j && j < N) atp[h * ldp + tri_index(i, j, N)] = inv * (fast_tanh(a[u][v] * scale) + fast_tanh(b[u][v] * scale));
      }
  }
}

// Tiny MLP on a register/local vector (per-edge hodge MLPs, per-entry rank-2 MLPs).
// in/out may not alias.  All widths <= SMALL_MAX (validated at plan creation).
__device__ __forceinline__ void small_mlp(const ccsd_mlp_t &m, const float *__restrict__ W, const float *in,
                                          float *out, int hidden_act) {
  float a[SMALL_MAX], t[SMALL_MAX];
  int K = m.din;
  for (int k = 0; k < K; ++k) a[k] = in[k];
  for (int l = 0; l < m.nl; ++l) {
    const bool last = (l == m.nl - 1);
    const int O = last ? m.dout : m.dhid;
    const int Opad = round_up(O, 8);
    const float *w = W + m.w[l], *bb = W + m.b[l];
    for (int o = 0; o < O; ++o) {
      float s = __ldg(bb + o);
      for (int k = 0; k < K; ++k) s += a[k] * __ldg(w + k * Opad + o);
      t[o] = last ? s : act_apply(s, hidden_act);
    }
    for (int o = 0; o < O; ++o) a[o] = t[o];
    K = O;
  }
  for (int o = 0; o < K; ++o) out[o] = a[o];
}

}  // namespace ccsd
