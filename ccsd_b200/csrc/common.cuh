// common.cuh -- build shim + small device helpers shared by every kernel.
//
// Two builds of the same sources:
//   * product:   nvcc -gencode arch=compute_100a,code=sm_100a  -> ccsd_b200/_lib/libccsd_b200.so
//   * CCSD_EMU:  g++ -x c++ -DCCSD_EMU                         -> tests/_emu/libccsd_b200_emu.so
// The emulation build runs each kernel body on the host with ONE thread per block, blocks in
// sequence (so __syncthreads is a no-op).  It exists only so that `pytest -m "not gpu"` can check
// the kernels' index arithmetic, weight packing and masks against the oracle on a box without a
// GPU.  It is test infrastructure: nothing in ccsd_b200/ loads it, and it cannot be selected at
// run time -- the product library has no CPU path.
#pragma once

// Non-template kernels of the shared headers: a translation unit that only needs the device helpers (the
// per-FMODE tc_apply units) defines CCSD_AUX_TU, which gives them internal linkage so the linker sees one copy.
#ifdef CCSD_AUX_TU
#define CCSD_KERNEL static __global__
#else
#define CCSD_KERNEL __global__
#endif

#include <stddef.h>
#include <stdint.h>

#ifndef CCSD_EMU
#include <cuda_runtime.h>
// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, DEVICE): remember what was raised on each device so
// that plans on several devices of one process all get the limit they need (only ever raised, several plans coexist).
struct CcsdSmemAttr { size_t v[64] = {0}; };
template <class F> static inline int ccsd_ensure_smem(F func, size_t bytes, CcsdSmemAttr &rec) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (bytes <= rec.v[dev]) return 0;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return -1;
  rec.v[dev] = bytes;
  return 0;
}
#endif

#ifdef CCSD_EMU
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
struct ccsd_dim3 {
  unsigned x = 1, y = 1, z = 1;
  ccsd_dim3() {}
  ccsd_dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef ccsd_dim3 dim3;
extern thread_local ccsd_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local float *ccsd_emu_smem;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __syncthreads() ((void)0)
#define __syncwarp() ((void)0)
template <class T> static inline T __ldg(const T *p) { return *p; }
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __uint2float_rn(uint32_t v) { return (float)v; }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline float __shfl_sync(unsigned, float v, int) { return v; }
#define CCSD_FAST_LOGF(v) logf(v)
#define CCSD_FAST_RSQRTF(v) (1.0f / sqrtf(v))
#define CCSD_FAST_SINCOSF(v, s, c) sincosf(v, s, c)
#define CCSD_SMEM(name) float *name = ccsd_emu_smem
typedef void *cudaStream_t;
template <class Fn> static inline void ccsd_emu_launch(dim3 grid, size_t smem_bytes, Fn fn) {
  std::vector<float> buf(smem_bytes / 4 + 16);
  gridDim = grid;
  blockDim = ccsd_dim3(1, 1, 1);
  threadIdx = ccsd_dim3(0, 0, 0);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx = ccsd_dim3(x, y, z);
        // poison shared memory so that reads of unwritten entries surface as NaN
        for (auto &v : buf) v = NAN;
        ccsd_emu_smem = buf.data();
        fn();
      }
}
#define CCSD_LAUNCH(kern, grid, block, smem, stream, ...) \
  ccsd_emu_launch(dim3(grid), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define CCSD_FAST_LOGF(v) __logf(v)
#define CCSD_FAST_RSQRTF(v) rsqrtf(v)
#define CCSD_FAST_SINCOSF(v, s, c) __sincosf(v, s, c)
#define CCSD_SMEM(name) extern __shared__ __align__(16) float name[]
#define CCSD_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

namespace ccsd {

// Block-cooperative global -> shared copy of n floats (n % 4 == 0, both 16-byte aligned) that does not
// block: cp.async on the device (complete after stage_wait()), a plain loop in the host emulation.
__device__ __forceinline__ void stage_async(float *dst_smem, const float *__restrict__ src, int n) {
#ifdef CCSD_EMU
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst_smem[i] = src[i];
#else
  const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst_smem);
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 4u * i), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
// This thread's staged copies have landed (a __syncthreads makes every thread's visible to the block).
__device__ __forceinline__ void stage_wait() {
#ifndef CCSD_EMU
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

enum { ACT_NONE = 0, ACT_ELU = 1, ACT_TANH = 2 };

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_ELU) return v > 0.f ? v : expm1f(v);
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Sum over the whole block.  `red` is >= 33 floats of shared memory.  Every thread gets the sum.
__device__ __forceinline__ float block_sum(float v, float *red) {
#ifdef CCSD_EMU
  (void)red;
  return v;
#else
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // protect `red` from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarp ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
#endif
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-7 + Box-Muller: 4 standard normals per call.  Seven rounds is the smallest Philox4x32 variant that passes
// BigCrush (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11, table 2; ten rounds is their safety
// margin): the draw sits on the per-entry hot path of every rank-2 pass (ncu: ~40 % of tc_apply's instructions with
// ten rounds), and bit-identity with torch's generator is not a goal -- distributional parity is tested.
// counter = (group, draw_id, sample_lo, sample_hi), key = seed.  `group` is the element index / 4
// inside one sample (rank-2 rows are padded to a multiple of 4 cells so that a group never
// straddles rows); the sample index is GLOBAL (shard offset + local), so a sample's noise does
// not depend on how the batch is split across GPUs.
// ---------------------------------------------------------------------------------------------
#ifndef CCSD_PHILOX_ROUNDS
#define CCSD_PHILOX_ROUNDS 7
#endif
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < CCSD_PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 23-bit uniform in (0, 1) from a random word without an int->float conversion: the top 23 bits become the
// mantissa of a float in [1, 2); subtracting (1 - 2^-24) centres the 2^-23 grid inside (0, 1).
__device__ __forceinline__ float u01(uint32_t r) {
#ifdef CCSD_EMU
  union { uint32_t u; float f; } c;
  c.u = (r >> 9) | 0x3F800000u;
  return c.f - 0.99999994f;
#else
  return __uint_as_float((r >> 9) | 0x3F800000u) - 0.99999994f;
#endif
}

__device__ __forceinline__ void normal4(uint64_t seed, uint64_t sample, uint32_t draw_id, uint32_t group,
                                        float z[4]) {
  uint32_t r[4];
#ifdef CCSD_EXPERIMENT_NO_PHILOX
  r[0] = group; r[1] = draw_id; r[2] = (uint32_t)sample; r[3] = (uint32_t)seed;   // timing experiment only
#else
  philox4x32(group, draw_id, (uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)seed,
             (uint32_t)(seed >> 32), r);
#endif
  // Box-Muller on the SFU: radius = sqrt(t) = t * rsqrt(t) with t = -2 ln u > 0, angle through the fast
  // sin/cos (|err| ~1e-6 on a unit normal: irrelevant for noise).  The draw is on the per-entry hot path of
  // every sampler pass, so its instruction count matters more than its last bits.
  const float t0 = -2.0f * CCSD_FAST_LOGF(u01(r[0])), t2 = -2.0f * CCSD_FAST_LOGF(u01(r[2]));
  const float ra = t0 * CCSD_FAST_RSQRTF(t0), rb = t2 * CCSD_FAST_RSQRTF(t2);
  float s, c;
  CCSD_FAST_SINCOSF(6.283185307179586f * u01(r[1]), &s, &c);
  z[0] = ra * c; z[1] = ra * s;
  CCSD_FAST_SINCOSF(6.283185307179586f * u01(r[3]), &s, &c);
  z[2] = rb * c; z[3] = rb * s;
}

// One standard normal for flat element `idx` of a sample (used for the small x / adj tensors).
__device__ __forceinline__ float normal1(uint64_t seed, uint64_t sample, uint32_t draw_id, uint32_t idx) {
  float z[4];
  normal4(seed, sample, draw_id, idx >> 2, z);
  return z[idx & 3];
}

// draw ids: object in the top bits, (step+1)*32 + slot below; the prior uses step = -1.
__device__ __forceinline__ uint32_t draw_id(int obj, int step, int slot) {
  return ((uint32_t)obj << 28) | (uint32_t)((step + 1) * 32 + slot);   // slot < 32: Langevin inner steps + predictor
}

}  // namespace ccsd
