// ts_probe.cu -- hardware probe (not product code): tcgen05.mma with the A operand in TENSOR MEMORY
// (".ts" form) and an MN-major, 128-byte-swizzled B operand in shared memory.  Checks the packing
// convention assumed by ccsd_b200/csrc/tc_apply.cuh: A[m][k] (bf16) lives at TMEM lane m, 32-bit
// column k/2, low half = even k.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/ts_probe tools/ts_probe.cu && /tmp/ts_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ccsd_b200/csrc/tc_common.cuh"

using namespace ccsd;

constexpr int M = 128, N = 32, K = 64;
constexpr int SLOT = 1;   // which 32-cell half of the 64-wide MN-major operand rows holds the tile

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__global__ void __launch_bounds__(160, 1) probe(const float *A, const float *Bm, float *D) {
  extern __shared__ uint8_t raw[];
  const uint32_t r0 = tc::smem_u32(raw);
  const uint32_t base = (r0 + 1023u) & ~1023u;
  uint8_t *gen = raw + (base - r0);
  const uint32_t sB = base;                   // [K rows][64 n] bf16 MN-major swizzled: K*128 bytes
  const uint32_t bars = base + K * 128;
  const uint32_t done = bars, tslot = bars + 16;
  uint32_t *tslot_gen = reinterpret_cast<uint32_t *>(gen + K * 128 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { tc::mbar_init(done, 1); tc::mbar_fence_init(); }
  if (warp == 4) tc::tmem_alloc(tslot, 128);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tslot_gen;
  if (warp < 4) {
    // A -> TMEM: lane m = 32*warp + lane, column j holds (A[m][2j], A[m][2j+1])
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat16 lo = __float2bfloat16_rn(A[m * K + 2 * (c0 + j)]), hi = __float2bfloat16_rn(A[m * K + 2 * (c0 + j) + 1]);
        r[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // B -> smem MN-major: element (k, n) at k*128 + (((n/8) ^ (k%8)) * 16) + (n%8)*2
    for (int t = threadIdx.x; t < K * N; t += 128) {
      const int k = t / N, n = t - k * N;
      const int nl = SLOT * 32 + n;   // logical column inside the 64-wide row
      const uint32_t off = (uint32_t)k * 128u + (uint32_t)(((nl >> 3) ^ (k & 7)) << 4) + (uint32_t)(nl & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16 *>(gen + off) = __float2bfloat16_rn(Bm[k * N + n]);
    }
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 1);
      for (int k4 = 0; k4 < K / 16; ++k4) {
        const uint64_t bd = tc::make_smem_desc(sB + (uint32_t)SLOT * 64u + (uint32_t)k4 * 2048u, 8192, 1024);
        umma_ts(tmem + 64, tmem + (uint32_t)k4 * 8u, bd, idesc, k4 != 0);
      }
      tc::umma_commit(done);
    }
    __syncwarp();
  } else {
    tc::mbar_wait(done, 0);
    tc::tc_fence_after_sync();
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 64u + (uint32_t)c0, v);
      for (int j = 0; j < 16; ++j) D[m * N + c0 + j] = v[j];
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 128);
}

int main() {
  std::vector<float> A(M * K), B(K * N), D(M * N), R(M * N, 0.f);
  srand(1);
  for (auto &v : A) v = (float)(rand() % 17 - 8);
  for (auto &v : B) v = (float)(rand() % 13 - 6);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[k * N + n];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const int smem = K * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 160, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double mx = 0; int bad = 0;
  for (int i = 0; i < M * N; ++i) { double d = fabs(D[i] - R[i]); if (d > mx) mx = d; if (d > 1e-3) ++bad; }
  printf("ts_probe: max |D - ref| = %g, mismatches = %d / %d  (D[0]=%g ref %g; D[65]=%g ref %g)\n", mx, bad, M * N, D[0], R[0], D[65], R[65]);
  return bad ? 2 : 0;
}
