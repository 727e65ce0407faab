// prims.cuh -- block-cooperative building blocks over shared memory (fp32 FMA path).
// Conventions: "feature-major" buffers are buf[f*ld + row]; "node-major" are buf[row*ld + f].
// Every primitive strides work items over the block and does NOT synchronise at the end.
#pragma once
#include "plan_dev.h"

namespace ccsd {

// out(r, o) = act(bias[o] + sum_k in(k, r) * W[k*Opad + o]),  Opad = round_up(O, 8)
// in(k, r) = k < K1 ? in1[k*ld1 + r] : in2[(k-K1)*ld2 + r]      (feature-major inputs)
// out(r, o) stored at out[r*sro + o*soo].  W, bias in global memory (read-only path).
__device__ __forceinline__ void dense2(const float *in1, int ld1, int K1, const float *in2, int ld2, int K2,
                                       const float *__restrict__ W, const float *__restrict__ bias, int O,
                                       float *out, int sro, int soo, int R, int act) {
  const int Opad = round_up(O, 8);
  const int nchunk = Opad >> 3;
  const int items = nchunk * R;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int chunk = it / R, r = it - chunk * R;
    const int oc = chunk << 3;
    float acc[8];
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = __ldg(bias + oc + j);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    const float *wp = W + oc;
    // the two input segments are walked with the same code; k is unrolled by 4 with every load of
    // the group issued before the FMAs (the weights come from L1/L2: latency, not bandwidth, bound)
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const float *in = seg ? in2 : in1;
      const int ld = seg ? ld2 : ld1, K = seg ? K2 : K1;
      int k = 0;
      for (; k + 4 <= K; k += 4) {
        float v[4];
        float4 w0[4], w1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = in[(k + u) * ld + r];
          w0[u] = __ldg(reinterpret_cast<const float4 *>(wp + u * Opad));
          w1[u] = __ldg(reinterpret_cast<const float4 *>(wp + u * Opad + 4));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[0] += v[u] * w0[u].x; acc[1] += v[u] * w0[u].y; acc[2] += v[u] * w0[u].z; acc[3] += v[u] * w0[u].w;
          acc[4] += v[u] * w1[u].x; acc[5] += v[u] * w1[u].y; acc[6] += v[u] * w1[u].z; acc[7] += v[u] * w1[u].w;
        }
        wp += 4 * Opad;
      }
      for (; k < K; ++k) {
        const float v = in[k * ld + r];
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wp));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(wp + 4));
        acc[0] += v * w0.x; acc[1] += v * w0.y; acc[2] += v * w0.z; acc[3] += v * w0.w;
        acc[4] += v * w1.x; acc[5] += v * w1.y; acc[6] += v * w1.z; acc[7] += v * w1.w;
        wp += Opad;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (oc + j < O) out[r * sro + (oc + j) * soo] = act_apply(acc[j], act);
  }
}

// A whole MLP over R rows.  Hidden activations ping-pong between hA/hB ([dhid x ldh], feature-major).
// Ends with a __syncthreads() after every layer (including the last).
__device__ __forceinline__ void mlp_rows(const ccsd_mlp_t &m, const float *__restrict__ W, const float *in1,
                                         int ld1, int K1, const float *in2, int ld2, int K2, int R, float *hA,
                                         float *hB, int ldh, float *out, int sro, int soo, int hidden_act,
                                         int out_act) {
  const float *cur1 = in1, *cur2 = in2;
  int c_ld1 = ld1, c_K1 = K1, c_ld2 = ld2, c_K2 = K2;
  for (int l = 0; l < m.nl; ++l) {
    const bool last = (l == m.nl - 1);
    const int O = last ? m.dout : m.dhid;
    float *dst = last ? out : ((l & 1) ? hB : hA);
    dense2(cur1, c_ld1, c_K1, cur2, c_ld2, c_K2, W + m.w[l], W + m.b[l], O, dst, last ? sro : 1,
           last ? soo : ldh, R, last ? out_act : hidden_act);
    __syncthreads();
    cur1 = dst; c_ld1 = ldh; c_K1 = O; cur2 = nullptr; c_ld2 = 0; c_K2 = 0;
  }
}

// DenseGCNConv normalisation (layers.py:139-147): A^ = adj with unit diagonal,
// d = rowsum(A^).clamp(min=1)^-1/2, an[i][j] = d_i * A^_ij * d_j.  adj: row-major N x N (ld = N).
// Contains two __syncthreads (dvec is ready after the first; `an` after the second).
__device__ __forceinline__ void gcn_norm(const float *adj, int N, float *dvec, float *an, int ldn) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < N; ++j) s += (j == i) ? 1.f : adj[i * N + j];
    dvec[i] = 1.0f / sqrtf(fmaxf(s, 1.f));
  }
  __syncthreads();
  for (int p = threadIdx.x; p < N * N; p += blockDim.x) {
    const int i = p / N, j = p - i * N;
    const float a = (i == j) ? 1.f : adj[p];
    an[i * ldn + j] = dvec[i] * a * dvec[j];
  }
  __syncthreads();
}

// out(i, f) = act(bias[f] + sum_j an[i][j] * xw[j*ldxw + f0 + f]),  f < nf, stored at out[i*sro + f*sfo].
// f0 and ldxw are multiples of 4 (float4 loads along f).
__device__ __forceinline__ void gcn_aggregate(const float *an, int ldn, int N, const float *xw, int ldxw, int f0,
                                              int nf, const float *__restrict__ bias, float *out, int sro,
                                              int sfo, int act) {
  const int ng = (nf + 3) >> 2;
  for (int it = threadIdx.x; it < ng * N; it += blockDim.x) {
    const int i = it / ng, g = it - i * ng;
    const int f = g << 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *row = an + i * ldn;
    const float *xp = xw + f0 + f;
    for (int j = 0; j < N; ++j) {
      const float w = row[j];
      const float4 v = *reinterpret_cast<const float4 *>(xp + j * ldxw);
      a0 += w * v.x; a1 += w * v.y; a2 += w * v.z; a3 += w * v.w;
    }
    const float r[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (f + q < nf) out[i * sro + (f + q) * sfo] = act_apply(r[q] + __ldg(bias + f + q), act);
  }
}

// Attention.forward score part (attention.py:111-130): heads are chunks of ds = ad / heads features
// (torch.split semantics: ceil(ad/ds) chunks), tanh(q.k * scale) averaged over chunks, symmetrised.
// qn: node-major [N x ldq]; kf: feature-major [ad x ldn]; att: row-major N x N.
__device__ __forceinline__ void attn_scores(const float *qn, int ldq, const float *kf, int ldn, int N, int ad,
                                            int heads, float scale, float *att) {
  const int ds = ad / heads;
  const int nch = (ad + ds - 1) / ds;
  const float inv = 1.0f / (float)nch;
  const int npair = N * (N + 1) / 2;
  for (int p = threadIdx.x; p < npair; p += blockDim.x) {
    // p -> (i <= j), row-major over the upper triangle
    int i = 0, rem = p;
    while (rem >= N - i) { rem -= N - i; ++i; }
    const int j = i + rem;
    float sij = 0.f, sji = 0.f;
    for (int c = 0; c < nch; ++c) {
      const int d0 = c * ds, d1 = (d0 + ds < ad) ? d0 + ds : ad;
      float a = 0.f, b2 = 0.f;
      for (int dd = d0; dd < d1; ++dd) {
        a += qn[i * ldq + dd] * kf[dd * ldn + j];
        b2 += qn[j * ldq + dd] * kf[dd * ldn + i];
      }
      sij += tanhf(a * scale);
      sji += tanhf(b2 * scale);
    }
    const float v = 0.5f * (sij * inv + sji * inv);
    att[i * N + j] = v;
    att[j * N + i] = v;
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
