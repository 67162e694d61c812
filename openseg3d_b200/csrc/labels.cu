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
