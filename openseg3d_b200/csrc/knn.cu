// knn.cu -- k nearest neighbours of each query point among the points of its own batch segment (brute force, exact).
// replaces: knn_query (seg3d/ops/knn_query/knn_query.py:8-26, src/knn_query_cuda.cu:23-112), used on the loss side to
//           give every coarse (level-4) voxel the label of its nearest fine voxel (tools/train.py:86-104).
//
// Same result as the reference's kernel, element for element: a max-heap of the k best squared distances per query
// (root = current worst; a key replaces the root only when STRICTLY closer, so on ties the lower index stays), keys
// visited in ascending index order, heap-sorted ascending at the end (knn_query_cuda.cu:23-49,83-111).  Squared
// distances are computed without FMA contraction ((dx*dx + dy*dy) + dz*dz, each op rounded), which is what the CPU
// oracle restates.  What differs is the data movement: a CTA of 256 queries walks the union of its queries' segments
// once, staging 2048 keys at a time in shared memory (structure-of-arrays, broadcast reads), instead of every thread
// streaming its whole segment from global memory; k = 1 (the only value the reference model uses) keeps its best
// candidate in registers.
#include "common.cuh"

namespace os3d {

constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 2048;
constexpr int kKnnMaxK = 100;      // the reference's local arrays hold 100 candidates (knn_query_cuda.cu:90-91)

__device__ __forceinline__ void knn_sift_down(float *dist, int32_t *idx, int k) {   // knn_query_cuda.cu:23-38
  int root = 0, child = 1;
  while (child < k) {
    if (child + 1 < k && dist[child + 1] > dist[child]) ++child;
    if (dist[root] > dist[child]) return;
    const float td = dist[root]; dist[root] = dist[child]; dist[child] = td;
    const int32_t ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
    root = child;
    child = 2 * root + 1;
  }
}

__device__ __forceinline__ int knn_segment(int64_t i, const int32_t *__restrict__ offset, int n_seg) {
  int b = 0;
  while (b < n_seg - 1 && i >= __ldg(offset + b)) ++b;    // get_bt_idx, knn_query_cuda.cu:52-63 (bounded)
  return b;
}

template <bool kOne>
__global__ void __launch_bounds__(kKnnThreads) knn_query_kernel(int64_t m, int nsample, const float *__restrict__ xyz,
                                                                const float *__restrict__ new_xyz,
                                                                const int32_t *__restrict__ offset,
                                                                const int32_t *__restrict__ new_offset, int n_seg,
                                                                int32_t *__restrict__ idx, float *__restrict__ dist2) {
  __shared__ float kx[kKnnTile], ky[kKnnTile], kz[kKnnTile];
  __shared__ int lo_s, hi_s;
  const int64_t q = (int64_t)blockIdx.x * kKnnThreads + threadIdx.x;
  const bool active = q < m;
  int start = 0, end = 0;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const int b = knn_segment(q, new_offset, n_seg);
    start = b == 0 ? 0 : __ldg(offset + b - 1);
    end = __ldg(offset + b);
    qx = __ldg(new_xyz + q * 3);
    qy = __ldg(new_xyz + q * 3 + 1);
    qz = __ldg(new_xyz + q * 3 + 2);
  }
  if (threadIdx.x == 0) { lo_s = 0x7fffffff; hi_s = 0; }
  __syncthreads();
  if (active && end > start) { atomicMin(&lo_s, start); atomicMax(&hi_s, end); }
  __syncthreads();
  const int lo = lo_s, hi = hi_s;

  float best_d[kOne ? 1 : kKnnMaxK];
  int32_t best_i[kOne ? 1 : kKnnMaxK];
  const int k = kOne ? 1 : nsample;
  for (int i = 0; i < k; ++i) { best_d[i] = 1e10f; best_i[i] = start; }     // knn_query_cuda.cu:92-95

  for (int t0 = lo; t0 < hi; t0 += kKnnTile) {
    const int tn = min(kKnnTile, hi - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += kKnnThreads) {
      kx[i] = __ldg(xyz + (int64_t)(t0 + i) * 3);
      ky[i] = __ldg(xyz + (int64_t)(t0 + i) * 3 + 1);
      kz[i] = __ldg(xyz + (int64_t)(t0 + i) * 3 + 2);
    }
    __syncthreads();
    const int a = max(start, t0) - t0, b = min(end, t0 + tn) - t0;
    for (int i = a; i < b; ++i) {
      const float dx = __fsub_rn(qx, kx[i]), dy = __fsub_rn(qy, ky[i]), dz = __fsub_rn(qz, kz[i]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d2 < best_d[0]) {
        best_d[0] = d2;
        best_i[0] = t0 + i;
        if (!kOne) knn_sift_down(best_d, best_i, k);
      }
    }
  }
  if (!active) return;
  if (!kOne) {
    for (int i = k - 1; i > 0; --i) {                  // heap_sort, knn_query_cuda.cu:41-49
      const float td = best_d[0]; best_d[0] = best_d[i]; best_d[i] = td;
      const int32_t ti = best_i[0]; best_i[0] = best_i[i]; best_i[i] = ti;
      knn_sift_down(best_d, best_i, i);
    }
  }
  for (int i = 0; i < k; ++i) {
    idx[q * nsample + i] = best_i[i];
    dist2[q * nsample + i] = best_d[i];
  }
}

}  // namespace os3d

using namespace os3d;

extern "C" int os3d_knn_query(const float *xyz, const float *new_xyz, int64_t m, int nsample, const int32_t *offset,
                              const int32_t *new_offset, int n_seg, int32_t *idx, float *dist2, void *stream) {
  if (nsample < 1 || nsample > kKnnMaxK || n_seg < 1 || m < 0) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  const unsigned g = (unsigned)cdiv(m, kKnnThreads);
  if (nsample == 1)
    knn_query_kernel<true><<<g, kKnnThreads, 0, (cudaStream_t)stream>>>(m, nsample, xyz, new_xyz, offset, new_offset, n_seg,
                                                                      idx, dist2);
  else
    knn_query_kernel<false><<<g, kKnnThreads, 0, (cudaStream_t)stream>>>(m, nsample, xyz, new_xyz, offset, new_offset, n_seg,
                                                                       idx, dist2);
  OS3D_LAUNCH_CHECK();
  return 0;
}
