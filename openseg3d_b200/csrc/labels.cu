// labels.cu -- loss-side label preparation: per-voxel majority labels.
// replaces WaymoDataset.prepare_voxel_labels (seg3d/datasets/waymo_dataset.py:213-246): a Python dict of 256-bin counters
// filled point by point in DataLoader workers, then np.argmax per voxel (ties -> the LOWEST label; voxels without a
// labelled point keep ignore_index).  Here: one RED.ADD per point into a dense [m, 32] histogram (labels 0..30 in bins
// 0..30, ignore_index in bin 31 -- the highest label, so the bin order is the label order and ties resolve as np.argmax
// does), then one thread per voxel picks the first maximal bin.  Integer work, bit-exact by construction.
#include "common.cuh"

namespace os3d {

constexpr int kLabelBins = 32;

__global__ void label_hist_kernel(const int64_t *__restrict__ pvid, const uint8_t *__restrict__ labels, int64_t n,
                                  int64_t m, int ignore, int32_t *__restrict__ hist, int32_t *__restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = __ldg(pvid + i);
  if (v < 0 || v >= m) return;                      // voxel_id == -1: point outside the range
  const int lab = labels[i];
  int bin;
  if (lab == ignore) bin = kLabelBins - 1;
  else if (lab < kLabelBins - 1 && lab < ignore) bin = lab;
  else { atomicOr(bad, 1); return; }                // a label this layout cannot order: reported to the caller
  atomicAdd(hist + v * kLabelBins + bin, 1);
}

__global__ void label_argmax_kernel(const int32_t *__restrict__ hist, int64_t m, int ignore, uint8_t *__restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m) return;
  const int4 *h = reinterpret_cast<const int4 *>(hist + v * kLabelBins);
  int best = 0, best_bin = -1;
#pragma unroll
  for (int q = 0; q < kLabelBins / 4; ++q) {
    const int4 c = __ldg(h + q);
    const int cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cc[j] > best) { best = cc[j]; best_bin = 4 * q + j; }       // strict >: the first maximum wins (np.argmax)
  }
  out[v] = best_bin < 0 ? (uint8_t)ignore : (best_bin == kLabelBins - 1 ? (uint8_t)ignore : (uint8_t)best_bin);
}

// Predicted label per point: argmax over the class logits of each row (tools/test.py:58, torch.argmax(point_out, dim=1)),
// written as uint8.  Ties -> the lowest class index; NaN never wins (a row of NaNs gives 0).  The rows of a block are one
// contiguous, 16-byte aligned span of memory (256 rows x c elements): it is staged in shared memory with 16-byte loads
// and each thread then scans its own row there -- the 46-byte rows of the 23-class head would otherwise be read as 23
// scattered 2-byte loads per thread.
constexpr int kArgRows = 256;
template <typename T>
__device__ __forceinline__ float logit_to_float(T v);
template <>
__device__ __forceinline__ float logit_to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float logit_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(kArgRows) argmax_rows_kernel(const T *__restrict__ x, int64_t n, int c, int rows_per_block,
                                                               uint8_t *__restrict__ out) {
  extern __shared__ uint4 stage4[];
  T *stage = reinterpret_cast<T *>(stage4);
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int rows = (int)min((int64_t)rows_per_block, n - row0);
  const int64_t elems = (int64_t)rows * c;
  const T *src = x + row0 * c;                                  // 16-byte aligned: rows_per_block is a multiple of 8
  constexpr int kPer16 = 16 / (int)sizeof(T);
  const int n16 = (int)(elems / kPer16);
  for (int i = threadIdx.x; i < n16; i += kArgRows) stage4[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
  for (int i = n16 * kPer16 + threadIdx.x; i < elems; i += kArgRows) stage[i] = src[i];
  __syncthreads();
  if ((int)threadIdx.x >= rows) return;
  const T *r = stage + (int)threadIdx.x * c;
  float best = -INFINITY;
  int arg = 0;
  for (int j = 0; j < c; ++j) {
    const float v = logit_to_float<T>(r[j]);
    if (v > best) { best = v; arg = j; }
  }
  out[row0 + threadIdx.x] = (uint8_t)arg;
}

}  // namespace os3d

using namespace os3d;

extern "C" int os3d_voxel_majority_labels(const int64_t *pvid, const uint8_t *labels, int64_t n, int64_t m, int ignore,
                                          int32_t *hist, int32_t *bad, uint8_t *out, void *stream) {
  if (n < 0 || m < 0 || ignore < kLabelBins - 1 || ignore > 255) return OS3D_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) return 0;
  OS3D_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * (size_t)m * kLabelBins, st));
  OS3D_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
  if (n > 0) label_hist_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(pvid, labels, n, m, ignore, hist, bad);
  label_argmax_kernel<<<(unsigned)cdiv(m, 256), 256, 0, st>>>(hist, m, ignore, out);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_argmax_rows(const void *x, int64_t n, int c, int elem_size, uint8_t *out, void *stream) {
  if (n < 0 || c < 1 || c > 256 || (elem_size != 2 && elem_size != 4)) return OS3D_ERR_BAD_ARG;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rows = kArgRows;                                               // rows per block: a multiple of 8 within 48 KB
  while ((size_t)rows * c * elem_size > 48 * 1024) rows -= 8;
  const size_t smem = (size_t)rows * c * elem_size;
  const unsigned grid = (unsigned)cdiv(n, rows);
  if (elem_size == 2) argmax_rows_kernel<__nv_bfloat16><<<grid, kArgRows, smem, st>>>((const __nv_bfloat16 *)x, n, c, rows, out);
  else argmax_rows_kernel<float><<<grid, kArgRows, smem, st>>>((const float *)x, n, c, rows, out);
  OS3D_LAUNCH_CHECK();
  return 0;
}
