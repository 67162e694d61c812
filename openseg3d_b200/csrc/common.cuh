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

// Two GELUs at once on the two-wide FP32 instructions of sm_100 (fma / mul / add .f32x2): the same erf approximation,
// 9 instructions per element instead of 14 (the MLP epilogues are bound by their instruction count).  The result is put
// together as x Phi(x) = x / 2 + |x| (1/2 - Phi(-|x|)), which needs no select; its absolute error against x * Phi(x) is
// below |x| 2^-25 (for x < -5, where x Phi(x) itself is below 2e-6).  y = gelu(a + b) for both lanes.
__device__ __forceinline__ void gelu_erf_fast2(float a0, float a1, float b0, float b1, float &y0, float &y1) {
  uint64_t x2, u2, t2, e2, p2;
  asm("{\n\t"
      ".reg .b64 a, b;\n\t"
      "mov.b64 a, {%1, %2};\n\t"
      "mov.b64 b, {%3, %4};\n\t"
      "add.rn.f32x2 %0, a, b;\n\t"
      "}" : "=l"(x2) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
  asm("and.b64 %0, %1, 0x7fffffff7fffffff;" : "=l"(u2) : "l"(x2));                          // |x|
  float d0, d1, u0, u1, t0, t1, e0, e1;
  {
    uint64_t den2, sq2;
    const uint64_t kp = 0x3e6d33883e6d3388ull;      // {0.23164189f, 0.23164189f}
    const uint64_t k1 = 0x3f8000003f800000ull;      // {1, 1}
    const uint64_t ke = 0xbf38aa3bbf38aa3bull;      // {-0.7213475f, -0.7213475f} = -log2(e) / 2
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(den2) : "l"(u2), "l"(kp), "l"(k1));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sq2) : "l"(u2));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(sq2) : "l"(sq2), "l"(ke));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(den2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(u0), "=f"(u1) : "l"(sq2));
  }
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(u0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(u1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "f"(t0), "f"(t1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(e2) : "f"(e0), "f"(e1));
  {
    const uint64_t c5 = 0x3f07dc223f07dc22ull;      // 0.5307027145
    const uint64_t c4 = 0xbf3a00e3bf3a00e3ull;      // -0.7265760135
    const uint64_t c3 = 0x3f35f0e33f35f0e3ull;      // 0.7107068705
    const uint64_t c2 = 0xbe11a98ebe11a98eull;      // -0.142248368
    const uint64_t c1 = 0x3e0279063e027906ull;      // 0.127414796
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(c5), "l"(t2), "l"(c4));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(t2), "l"(c3));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(t2), "l"(c2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p2) : "l"(p2), "l"(t2), "l"(c1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p2) : "l"(p2), "l"(t2));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p2) : "l"(p2), "l"(e2));                           // Phi(-|x|)
    const uint64_t kh = 0x3f0000003f000000ull;      // {0.5, 0.5}
    const uint64_t kn = 0xbf800000bf800000ull;      // {-1, -1}
    uint64_t w2, hx2, y2;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(w2) : "l"(p2), "l"(kn), "l"(kh));               // 1/2 - Phi(-|x|)
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hx2) : "l"(x2), "l"(kh));                          // x / 2
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y2) : "l"(u2), "l"(w2), "l"(hx2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y2));
  }
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
