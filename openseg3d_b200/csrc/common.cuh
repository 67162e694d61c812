// common.cuh -- shared device helpers for libos3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/os3d.h"

#define OS3D_LAUNCH_CHECK()                      \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

#define OS3D_CUDA(call)                          \
  do {                                           \
    cudaError_t e__ = (call);                    \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

namespace os3d {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;  // elements per block of the 3-phase scan

__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// GELU, erf form (nn.GELU() default), with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7: at most one bf16 ulp
// in the result): erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z >= 0 -- one reciprocal, one ex2,
// six FMAs instead of erff()'s ~20 instructions.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  // in terms of u = |x| (z = u / sqrt 2): t = 1 / (1 + (p / sqrt 2) u), exp(-z^2) = 2^(-(log2 e / 2) u^2); the factor 1/2 of
  // Phi(-u) = erfc(z) / 2 is folded into the polynomial coefficients.  14 instructions, two of them MUFU.
  const float u = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, u, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((u * u) * -0.72134752044448170f));
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  const float half_erfc = (poly * t) * e;                             // Phi(-|x|)
  return x * (x >= 0.0f ? 1.0f - half_erfc : half_erfc);             // x * Phi(x)
}

// Exclusive scan of one int per thread across a 256-thread block. Returns exclusive prefix; total in *total.
__device__ __forceinline__ int block_excl_scan_256(int v, int *total) {
  __shared__ int warp_sums[8];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < 8 ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < 8) warp_sums[lane] = wi - w;  // exclusive warp offsets
    if (lane == 7) block_total = wi;
  }
  __syncthreads();
  int res = incl - v + warp_sums[wid];
  *total = block_total;
  __syncthreads();  // allow reuse of the shared arrays by a following call
  return res;
}

// Single-block exclusive scan of block_sums[0..n) in place; writes the grand total to block_sums[n].
static __global__ void scan_block_sums_kernel(int32_t *block_sums, int64_t n) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += kScanThreads) {
    int64_t i = base + threadIdx.x;
    int v = i < n ? block_sums[i] : 0;
    int total;
    int ex = block_excl_scan_256(v, &total);
    int carry = carry_s;
    if (i < n) block_sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[n] = carry_s;
}

// Same for `lanes` interleaved counters: element (i, l) at block_sums[i*lanes + l].
static __global__ void scan_block_sums_multi_kernel(int32_t *block_sums, int64_t n, int lanes) {
  __shared__ int carry_s;
  for (int l = 0; l < lanes; ++l) {
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += kScanThreads) {
      int64_t i = base + threadIdx.x;
      int v = i < n ? block_sums[i * lanes + l] : 0;
      int total;
      int ex = block_excl_scan_256(v, &total);
      int carry = carry_s;
      if (i < n) block_sums[i * lanes + l] = ex + carry;
      __syncthreads();
      if (threadIdx.x == 0) carry_s = carry + total;
      __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[n * lanes + l] = carry_s;
    __syncthreads();
  }
}

// ---- open addressing table -----------------------------------------------------------------------
__device__ __forceinline__ int32_t table_find(const os3d_slot_t *__restrict__ table, uint64_t mask, int64_t key) {
  uint64_t h = mix64((uint64_t)key) & mask;
  while (true) {
    // one 16-byte load per probe
    const int4 s = __ldg(reinterpret_cast<const int4 *>(table + h));
    const int64_t k = (int64_t)(((uint64_t)(uint32_t)s.y << 32) | (uint32_t)s.x);
    if (k == key) return s.z;
    if (k == -1) return -1;
    h = (h + 1) & mask;
  }
}

// Insert key if absent. Returns the slot index.
__device__ __forceinline__ int64_t table_insert(os3d_slot_t *table, uint64_t mask, int64_t key) {
  uint64_t h = mix64((uint64_t)key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(&table[h].key),
                                        (unsigned long long)(-1LL), (unsigned long long)key);
    if (prev == (unsigned long long)(-1LL) || prev == (unsigned long long)key) return (int64_t)h;
    h = (h + 1) & mask;
  }
}

}  // namespace os3d
