// voxelize.cu -- stage 1a: dynamic voxelization with first-occurrence voxel ids, bit-exact with the reference's
// serial numba loop (seg3d/core/voxel/voxel_generator.py:99-153) and collate_batch's id offsets
// (seg3d/datasets/waymo_dataset.py:347-365).
//
// The serial loop hands out voxel ids in order of first appearance.  In parallel that is:
//   1. insert  : key = linear (b,z,y,x) index into an open-addressing table, atomicMin(first point index)
//   2. flag    : point i is the "first" of its voxel iff table[slot].first == i ; block sums of the flags
//   3. scan    : exclusive scan of the block sums (single block)
//   4. assign  : voxel id of a first point = number of first points before it; emit coors row, pvid[i]
//   5. spread  : every other point copies the id of its voxel's first point
// All five kernels are HBM/L2-streaming passes over n points (28 B/point read once through shared-memory
// staging, 12 B/point written) -- see DESIGN.md "Kernels" for the byte counts.
#include "common.cuh"

namespace os3d {

struct VoxGeom {
  float lo[3], vs[3];
  int grid[3];  // X, Y, Z
};

constexpr int kVoxThreads = 256;
constexpr int kMaxStride = 16;

// 1. insert.  One thread per point; the block's rows are staged through shared memory so the global read is a
//    single fully coalesced sweep regardless of the row stride.
__global__ void __launch_bounds__(kVoxThreads) vox_insert_kernel(const float *__restrict__ points, int64_t n,
                                                                  int stride, int has_batch, VoxGeom g,
                                                                  os3d_slot_t *table, uint64_t mask,
                                                                  int32_t *__restrict__ slot_of) {
  __shared__ float tile[kVoxThreads * kMaxStride];
  const int64_t row0 = (int64_t)blockIdx.x * kVoxThreads;
  const int rows = (int)min((int64_t)kVoxThreads, n - row0);
  const float *src = points + row0 * stride;
  for (int t = threadIdx.x; t < rows * stride; t += kVoxThreads) tile[t] = __ldg(src + t);
  __syncthreads();
  if ((int)threadIdx.x >= rows) return;
  const float *p = tile + threadIdx.x * stride;
  const int off = has_batch ? 1 : 0;
  int c[3];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    // IEEE f32 subtract, divide, floor -- no FMA contraction, no reciprocal (voxel_generator.py:139)
    const float q = floorf(__fdiv_rn(__fsub_rn(p[off + j], g.lo[j]), g.vs[j]));
    ok = ok && (q >= 0.0f) && (q < (float)g.grid[j]);
    c[j] = (int)q;
  }
  const int64_t i = row0 + threadIdx.x;
  if (!ok) {
    slot_of[i] = -1;
    return;
  }
  const int64_t b = has_batch ? (int64_t)p[0] : 0;
  const int64_t key = ((b * g.grid[2] + c[2]) * g.grid[1] + c[1]) * (int64_t)g.grid[0] + c[0];
  const int64_t s = table_insert(table, mask, key);
  atomicMin(reinterpret_cast<unsigned int *>(&table[s].val), (unsigned int)i);
  slot_of[i] = (int32_t)s;
}

__device__ __forceinline__ int vox_is_first(const os3d_slot_t *table, const int32_t *slot_of, int64_t i, int64_t n) {
  if (i >= n) return 0;
  const int32_t s = slot_of[i];
  return (s >= 0 && table[s].val == (int32_t)i) ? 1 : 0;
}

// 2. flag + block sums
__global__ void __launch_bounds__(kScanThreads) vox_flag_kernel(const os3d_slot_t *__restrict__ table,
                                                                 const int32_t *__restrict__ slot_of, int64_t n,
                                                                 int32_t *__restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int cnt = 0;
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) cnt += vox_is_first(table, slot_of, base + t, n);
  int total;
  block_excl_scan_256(cnt, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// 4. assign ids to first points, emit coordinates
__global__ void __launch_bounds__(kScanThreads) vox_assign_kernel(const os3d_slot_t *__restrict__ table,
                                                                   const int32_t *__restrict__ slot_of, int64_t n,
                                                                   const int32_t *__restrict__ block_sums,
                                                                   int64_t n_blocks, VoxGeom g,
                                                                   int32_t *__restrict__ coors,
                                                                   int64_t *__restrict__ pvid,
                                                                   int32_t *__restrict__ num_voxels) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int f[kScanItems];
  int cnt = 0;
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    f[t] = vox_is_first(table, slot_of, base + t, n);
    cnt += f[t];
  }
  int total;
  int ex = block_excl_scan_256(cnt, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    const int64_t i = base + t;
    if (i >= n) break;
    if (f[t]) {
      int64_t key = table[slot_of[i]].key;
      const int x = (int)(key % g.grid[0]); key /= g.grid[0];
      const int y = (int)(key % g.grid[1]); key /= g.grid[1];
      const int z = (int)(key % g.grid[2]); key /= g.grid[2];
      reinterpret_cast<int4 *>(coors)[ex] = make_int4((int)key, z, y, x);
      pvid[i] = ex;
      ++ex;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *num_voxels = block_sums[n_blocks];
}

// 5. spread the id of the voxel's first point to every other point
__global__ void vox_spread_kernel(const os3d_slot_t *__restrict__ table, const int32_t *__restrict__ slot_of,
                                  int64_t n, int64_t *__restrict__ pvid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t s = slot_of[i];
  if (s < 0) {
    pvid[i] = -1;
    return;
  }
  const int32_t first = table[s].val;
  if (first != (int32_t)i) pvid[i] = pvid[first];
}

__global__ void cart2polar_kernel(const float *__restrict__ in, int64_t n, int in_stride, int has_batch,
                                  float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = in + i * in_stride;
  float *o = out + i * (in_stride + 2);
  const int off = has_batch ? 1 : 0;
  if (has_batch) o[0] = p[0];
  const float x = p[off], y = p[off + 1], z = p[off + 2];
  o[off] = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
  o[off + 1] = atan2f(y, x);
  o[off + 2] = z;
  o[off + 3] = x;
  o[off + 4] = y;
  for (int j = off + 3; j < in_stride; ++j) o[j + 2] = p[j];
}

}  // namespace os3d

using namespace os3d;

extern "C" int os3d_voxelize_scratch(int64_t n, int64_t *hash_cap, int64_t *n_blocks) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  *hash_cap = cap;
  *n_blocks = cdiv(n > 0 ? n : 1, kScanTile);
  return 0;
}

extern "C" int os3d_voxelize(const float *points, int64_t n, int stride, int has_batch, float lo_x, float lo_y,
                             float lo_z, float vs_x, float vs_y, float vs_z, int grid_x, int grid_y, int grid_z,
                             os3d_slot_t *table, int64_t hash_cap, int32_t *slot_of, int32_t *block_sums,
                             int64_t n_blocks, int32_t *coors, int64_t *pvid, int32_t *num_voxels, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (stride < 3 + (has_batch ? 1 : 0) || stride > kMaxStride || (hash_cap & (hash_cap - 1)) || hash_cap < 2 * n ||
      n_blocks != cdiv(n > 0 ? n : 1, kScanTile))
    return OS3D_ERR_BAD_ARG;
  if (n == 0) {
    OS3D_CUDA(cudaMemsetAsync(num_voxels, 0, sizeof(int32_t), st));
    return 0;
  }
  VoxGeom g{{lo_x, lo_y, lo_z}, {vs_x, vs_y, vs_z}, {grid_x, grid_y, grid_z}};
  OS3D_CUDA(cudaMemsetAsync(table, 0xff, sizeof(os3d_slot_t) * (size_t)hash_cap, st));
  vox_insert_kernel<<<(unsigned)cdiv(n, kVoxThreads), kVoxThreads, 0, st>>>(points, n, stride, has_batch, g, table,
                                                                             (uint64_t)hash_cap - 1, slot_of);
  vox_flag_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(table, slot_of, n, block_sums);
  scan_block_sums_kernel<<<1, kScanThreads, 0, st>>>(block_sums, n_blocks);
  vox_assign_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(table, slot_of, n, block_sums, n_blocks, g, coors, pvid,
                                                                 num_voxels);
  vox_spread_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(table, slot_of, n, pvid);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_cart2polar_rows(const float *in, int64_t n, int in_stride, int has_batch, float *out, void *stream) {
  if (n == 0) return 0;
  cart2polar_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, in_stride, has_batch, out);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" const char *os3d_error_string(int code) {
  if (code == 0) return "ok";
  if (code == OS3D_ERR_BAD_ARG) return "os3d: bad argument";
  return cudaGetErrorString((cudaError_t)code);
}

extern "C" int os3d_version(void) { return 1; }
