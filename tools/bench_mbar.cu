// Microbenchmark of the mbarrier ring handshake on one SM-resident CTA per SM (no data movement):
//   P producer warps and one consumer warp exchange `slots` ring slots of depth D through full/empty mbarriers.
//   mode bit 0: producers poll with every lane (else lane 0 + __syncwarp)
//   mode bit 1: producers arrive with every lane (barrier count 32 P) (else lane 0, count P)
//   mode bit 2: consumer frees the slot with tcgen05.commit (else mbarrier.arrive)
// Prints cycles per slot.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/bench_mbar tools/bench_mbar.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../openseg3d_b200/csrc/tc_ptx.cuh"
using namespace os3d::ptx;

__global__ void __launch_bounds__(320, 1) k(int P, int D, int slots, int mode, long long *out) {
  __shared__ uint64_t bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full = smem_u32(bars), empty = smem_u32(bars + 32);
  if (threadIdx.x == 0) {
    for (int s = 0; s < D; ++s) { mbar_init(full + 8 * s, (mode & 2) ? 32 * P : P); mbar_init(empty + 8 * s, 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < P) {
    for (int q = 0; q < slots; ++q) {
      const int s = q % D; const uint32_t ph = ((q / D) & 1) ^ 1;
      if (mode & 1) mbar_wait(empty + 8 * s, ph);
      else { if (lane == 0) mbar_wait(empty + 8 * s, ph); __syncwarp(); }
      if (mode & 2) mbar_arrive(full + 8 * s);
      else { __syncwarp(); if (lane == 0) mbar_arrive(full + 8 * s); }
    }
  } else if (warp == 8) {
    for (int q = 0; q < slots; ++q) {
      const int s = q % D; const uint32_t ph = (q / D) & 1;
      mbar_wait(full + 8 * s, ph);
      if (elect_one()) { if (mode & 4) umma_commit(empty + 8 * s); else mbar_arrive(empty + 8 * s); }
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  long long *d; cudaMalloc(&d, 8);
  const int slots = 20000;
  for (int P : {8, 4, 1}) for (int D : {4, 8, 12}) for (int mode = 0; mode < 8; ++mode) {
    k<<<148, 320>>>(P, D, slots, mode, d);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("P=%d D=%2d poll_all=%d arrive_all=%d commit=%d : %7.1f cycles/slot\n", P, D, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1,
           (double)h / slots);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
