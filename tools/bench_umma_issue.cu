// Microbenchmark of the tcgen05.mma ISSUE path (what paced swformer_mlp_tc_kernel until it got two issuing warps, and
// what paces the per-key-block loop of window_attention_tc_kernel): one CTA per SM, W issuing warps, each issuing
// `groups` groups of G MMAs (M = 128, N, K = 16, bf16, operands = zeroed shared memory, no data dependence) followed
// by one tcgen05.commit; optionally waiting on that commit's mbarrier before the next group (the hand-off an epilogue
// would need).  Prints cycles per MMA as seen by the issuing thread and the MMA's own execution time N / 2 cycles
// for comparison.
//   mode bit 0: wait for each group's commit (serialised hand-off) instead of free-running
//   mode bit 1: issue under `if (lane == 0)` instead of warp-uniform + elect.sync
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/bench_umma_issue tools/bench_umma_issue.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../openseg3d_b200/csrc/tc_ptx.cuh"
using namespace os3d::ptx;

constexpr int kMaxWarps = 4;

__global__ void __launch_bounds__(32 * kMaxWarps, 1) k(int W, int G, int groups, int n, int mode, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[kMaxWarps];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (16384u + 32768u) / 16; i += blockDim.x)       // A: 128 x 64, B: 256 x 64 (SW128 blocks)
    reinterpret_cast<uint4 *>(smem_raw + (base - raw))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int w = 0; w < kMaxWarps; ++w) mbar_init(smem_u32(&bars[w]), 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long dt = 0;
  if (warp < W) {
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t a_lo = (uint32_t)make_kmajor_sw128_desc(base), b_lo = (uint32_t)make_kmajor_sw128_desc(base + 16384u);
    const uint32_t idesc = make_idesc_bf16(128, n);
    const uint32_t d = tmem + (uint32_t)warp * 128u;                 // each issuing warp has its own accumulator (n <= 128)
    const uint32_t bar = smem_u32(&bars[warp]);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (mode & 2) {
        if (lane == 0) {
          for (int i = 0; i < G; ++i) umma_bf16_lo(d, a_lo + 2 * (i & 3), b_lo + 2 * (i & 3), desc_hi, idesc, i > 0);
          umma_commit(bar);
        }
        __syncwarp();
      } else {
        if (elect_one()) {
          for (int i = 0; i < G; ++i) umma_bf16_lo(d, a_lo + 2 * (i & 3), b_lo + 2 * (i & 3), desc_hi, idesc, i > 0);
          umma_commit(bar);
        }
        __syncwarp();
      }
      if (mode & 1) {
        mbar_wait(bar, (uint32_t)g & 1u);
        tc_fence_after();
      }
    }
    if (!(mode & 1)) {                                               // approximate drain: nobody watched the intermediate
      // phases, so this parity may already have completed for an earlier group; the issuer-side time is what is measured
      mbar_wait(bar, (uint32_t)(groups - 1) & 1u);
    }
    dt = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
  if (blockIdx.x == 0 && lane == 0 && warp < W) out[warp] = dt;
}

int main() {
  long long *d, h[kMaxWarps];
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int groups = 4000;
  for (int n : {16, 64, 128})
    for (int W : {1, 2, 4})
      for (int G : {1, 4, 12})
        for (int mode = 0; mode < 4; ++mode) {
          k<<<148, 32 * kMaxWarps, 16384 + 32768 + 1024>>>(W, G, groups, n, mode, d);
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          printf("N=%3d warps=%d group=%2d wait_each=%d lane0_issue=%d : %7.1f cycles/MMA at the issuer (MMA itself ~%d), %7.1f cycles/group\n",
                 n, W, G, mode & 1, (mode >> 1) & 1, (double)h[0] / ((double)groups * G), n / 2, (double)h[0] / groups);
        }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
