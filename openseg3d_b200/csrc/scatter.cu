// scatter.cu -- stage 1b/1c: point->voxel scatter_max / scatter_mean and voxel->point gather.
// replaces torch_scatter.scatter as called by VFE.forward (seg3d/models/voxel_encoders/vfe.py:24-25), the in-tree
// voxel_pooling kernels (seg3d/ops/voxel_pooling/src/voxel_pooling_cuda.cu:11-41, launched <<<N, C>>>) and
// voxel_to_point (seg3d/ops/voxel_to_point/voxel_to_point.py:5-17).
//
// All kernels are HBM streaming kernels: a thread owns one float4 of one row, a warp covers 512 contiguous bytes,
// index loads are one per row (broadcast within the row's lanes), reductions go to L2 as fire-and-forget RED ops
// (vector red.global.add.v4.f32 for the mean; signed-max / unsigned-min integer REDs for the max, which makes the
// max bit-exact and order independent).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace os3d {

__device__ __forceinline__ void red_max_f32(float *addr, float v) {
  // order-preserving integer trick: non-negative floats compare like signed ints, negative floats compare
  // reversed like unsigned ints.  Works against a -inf (0xff800000) initialised destination.
  if (!signbit(v)) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void fill_f32_kernel(float *__restrict__ p, int64_t n4, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) reinterpret_cast<float4 *>(p)[i] = make_float4(v, v, v, v);
  if (i == 0) for (int64_t t = n4 * 4; t < n; ++t) p[t] = v;
}

// c4 = c/4 when VEC, else c.  One thread per (row, vector).
template <bool VEC>
__global__ void scatter_max_kernel(const float *__restrict__ feats, const int64_t *__restrict__ ids, int64_t n, int c,
                                   int cv, float *__restrict__ out, int64_t m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * cv) return;
  const int64_t row = t / cv;
  const int col = (int)(t - row * cv);
  const int64_t id = __ldg(ids + row);
  if (id < 0 || id >= m) return;
  if (VEC) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(feats + row * c) + col);
    float *o = out + id * c + col * 4;
    red_max_f32(o, v.x); red_max_f32(o + 1, v.y); red_max_f32(o + 2, v.z); red_max_f32(o + 3, v.w);
  } else {
    red_max_f32(out + id * c + col, __ldg(feats + row * c + col));
  }
}

__global__ void fix_empty_kernel(float *__restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && __float_as_uint(out[i]) == 0xff800000u) out[i] = 0.0f;
}

// bf16 features (the point encoder's output in bf16 inference): read as they are -- no fp32 copy of the point features --
// and reduced into an fp32 scratch with the same order-preserving REDs; finalize_max_bf16_kernel then writes the voxel
// features as bf16 (exact: every maximum IS one of the bf16 inputs) and zeroes the rows nothing was scattered to.
// One thread per (row, 8 channels).
template <typename IdT>
__global__ void scatter_max_bf16_kernel(const uint4 *__restrict__ feats, const IdT *__restrict__ ids, int64_t n, int c,
                                        int cv, float *__restrict__ acc, int64_t m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * cv) return;
  const int64_t row = t / cv;
  const int col = (int)(t - row * cv);
  const int64_t id = (int64_t)__ldg(ids + row);
  if (id < 0 || id >= m) return;
  const uint4 v = __ldg(feats + t);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float *o = acc + id * c + col * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red_max_f32(o + 2 * i, __uint_as_float(w[i] << 16));
    red_max_f32(o + 2 * i + 1, __uint_as_float(w[i] & 0xffff0000u));
  }
}

__global__ void finalize_max_bf16_kernel(const float4 *__restrict__ acc, uint2 *__restrict__ out, int64_t n4, int fix_empty) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = acc[i];
  uint32_t u[4] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)};
  if (fix_empty) {
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = u[k] == 0xff800000u ? 0u : u[k];
  }
  out[i] = make_uint2((u[0] >> 16) | (u[1] & 0xffff0000u), (u[2] >> 16) | (u[3] & 0xffff0000u));   // values are bf16 already
}

// ---- sort-based maximum (bf16): no atomics -------------------------------------------------------------------------
// A lidar voxel holds ~1.5 points, so the RED formulation above issues one 4-byte atomic per (point, channel) -- 93 M of
// them per 8-frame batch, 0.4 ms -- to combine next to nothing.  Here the point rows are sorted by voxel id (one cub radix
// sort over ceil(log2 m) key bits, values = point rows); every run of equal ids is then reduced by the 8-lane group that
// finds its head and written once as a bf16 row: 128-byte reads, 128-byte writes, nothing else.
__global__ void voxel_keys_kernel(const int64_t *__restrict__ ids64, const int32_t *__restrict__ ids32, int64_t n, int64_t m,
                                  uint32_t *__restrict__ keys, int32_t *__restrict__ rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids64 ? __ldg(ids64 + i) : (int64_t)__ldg(ids32 + i);
  keys[i] = (id >= 0 && id < m) ? (uint32_t)id : (uint32_t)m;      // skipped points sort behind every voxel
  rows[i] = (int32_t)i;
}

// cv = c / 8 lanes per run (one 16-byte chunk each); a block of 256 threads covers 256 / cv sorted positions
__global__ void __launch_bounds__(256) run_max_bf16_kernel(const uint4 *__restrict__ feats, const uint32_t *__restrict__ keys,
                                                          const int32_t *__restrict__ rows, int64_t n, int cv, int64_t m,
                                                          uint4 *__restrict__ out) {
  const int per_block = 256 / cv;
  const int g = threadIdx.x / cv, col = threadIdx.x - g * cv;
  if (g >= per_block) return;
  const int64_t i = (int64_t)blockIdx.x * per_block + g;
  if (i >= n) return;
  const uint32_t key = __ldg(keys + i);
  if (key >= (uint32_t)m || (i > 0 && __ldg(keys + i - 1) == key)) return;        // not a run head
  float best[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) best[k] = -INFINITY;
  for (int64_t j = i; j < n && __ldg(keys + j) == key; ++j) {
    const uint4 v = __ldg(feats + (int64_t)__ldg(rows + j) * cv + col);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      best[2 * k] = fmaxf(best[2 * k], __uint_as_float(w[k] << 16));
      best[2 * k + 1] = fmaxf(best[2 * k + 1], __uint_as_float(w[k] & 0xffff0000u));
    }
  }
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = (__float_as_uint(best[2 * k]) >> 16) | (__float_as_uint(best[2 * k + 1]) & 0xffff0000u);
  out[(int64_t)key * cv + col] = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void fill_u32_kernel(uint4 *__restrict__ p, int64_t n4, uint32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) p[i] = make_uint4(v, v, v, v);
}

template <bool VEC>
__global__ void scatter_add_kernel(const float *__restrict__ feats, const int64_t *__restrict__ ids, int64_t n, int c,
                                   int cv, float *__restrict__ out, int32_t *__restrict__ counts,
                                   const int32_t *__restrict__ counts_in, int64_t m) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * cv) return;
  const int64_t row = t / cv;
  const int col = (int)(t - row * cv);
  const int64_t id = __ldg(ids + row);
  if (id < 0 || id >= m) return;
  if (col == 0 && counts) atomicAdd(counts + id, 1);
  // avg pooling with caller counts divides before accumulating, like voxel_pooling_cuda.cu:23
  const float s = counts_in ? (float)__ldg(counts_in + id) : 1.0f;
  if (VEC) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(feats + row * c) + col);
    if (counts_in) { v.x = __fdiv_rn(v.x, s); v.y = __fdiv_rn(v.y, s); v.z = __fdiv_rn(v.z, s); v.w = __fdiv_rn(v.w, s); }
    red_add_v4(out + id * c + col * 4, v);
  } else {
    atomicAdd(out + id * c + col, __fdiv_rn(__ldg(feats + row * c + col), s));
  }
}

// Few destination rows (the SE layer pools 1.45 M points into `batch` rows): per-address atomics would serialise, so
// every thread keeps a private float4 running sum over a strided set of rows, flushing to L2 only when the id changes
// (ids arrive sorted by frame) and once at the end.  kRowsPerBlock rows per CTA.
constexpr int kSmallMRows = 4096;
__global__ void __launch_bounds__(256) scatter_add_runs_kernel(const float *__restrict__ feats,
                                                              const int64_t *__restrict__ ids, int64_t n, int c, int cv,
                                                              float *__restrict__ out, int32_t *__restrict__ counts,
                                                              int64_t m) {
  const int col = threadIdx.x % cv;                 // float4 column owned by this thread
  const int lane_row = threadIdx.x / cv, row_step = 256 / cv;
  if (lane_row >= row_step) return;
  const int64_t row_end = min(n, (int64_t)(blockIdx.x + 1) * kSmallMRows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t cur = -1;
  int cnt = 0;
  for (int64_t row = (int64_t)blockIdx.x * kSmallMRows + lane_row; row < row_end; row += row_step) {
    const int64_t id = __ldg(ids + row);
    if (id < 0 || id >= m) continue;
    if (id != cur) {
      if (cur >= 0) {
        red_add_v4(out + cur * c + col * 4, acc);
        if (col == 0 && counts) atomicAdd(counts + cur, cnt);
      }
      cur = id; acc = make_float4(0.f, 0.f, 0.f, 0.f); cnt = 0;
    }
    const float4 v = __ldg(reinterpret_cast<const float4 *>(feats + row * c) + col);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    ++cnt;
  }
  if (cur >= 0) {
    red_add_v4(out + cur * c + col * 4, acc);
    if (col == 0 && counts) atomicAdd(counts + cur, cnt);
  }
}

// the same for bf16 rows read as they are (the SE layer pools the bf16 fused point features): a thread owns 8 channels
template <typename IdT>
__global__ void __launch_bounds__(256) scatter_add_runs_bf16_kernel(const uint4 *__restrict__ feats, const IdT *__restrict__ ids,
                                                                   int64_t n, int c, int cv, float *__restrict__ out,
                                                                   int32_t *__restrict__ counts, int64_t m) {
  const int col = threadIdx.x % cv;                 // 16-byte chunk (8 channels) owned by this thread
  const int lane_row = threadIdx.x / cv, row_step = 256 / cv;
  if (lane_row >= row_step) return;
  const int64_t row_end = min(n, (int64_t)(blockIdx.x + 1) * kSmallMRows);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
  int64_t cur = -1;
  int cnt = 0;
  auto flush = [&]() {
    float *o = out + cur * c + col * 8;
    red_add_v4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
    red_add_v4(o + 4, make_float4(acc[4], acc[5], acc[6], acc[7]));
    if (col == 0 && counts) atomicAdd(counts + cur, cnt);
  };
  // four rows in flight per thread: the loop is a chain of dependent L2 / HBM reads otherwise
  for (int64_t row = (int64_t)blockIdx.x * kSmallMRows + lane_row; row < row_end; row += 4 * (int64_t)row_step) {
    int64_t id4[4];
    uint4 v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = row + (int64_t)u * row_step;
      id4[u] = r < row_end ? (int64_t)__ldg(ids + r) : -1;
      if (id4[u] >= 0 && id4[u] < m) v4[u] = __ldg(feats + r * cv + col);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t id = id4[u];
      if (id < 0 || id >= m) continue;
      if (id != cur) {
        if (cur >= 0) flush();
        cur = id; cnt = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
      }
      const uint32_t w[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] += __uint_as_float(w[k] << 16);
        acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
      }
      ++cnt;
    }
  }
  if (cur >= 0) flush();
}

__global__ void mean_normalize_kernel(float *__restrict__ out, const int32_t *__restrict__ counts, int64_t m, int c) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * c) return;
  const int cnt = __ldg(counts + t / c);
  out[t] = out[t] / (float)max(cnt, 1);
}

// Backward of scatter_max with torch_scatter's semantics: the gradient of out[v, c] goes to ONE argmax row.  Ties are not
// exotic (bf16-rounded features, duplicated returns); among tied rows the lowest point index is taken (torch_scatter's
// choice among ties is whichever thread wrote last -- any single tied row is a valid member).  Pass 1: arg[v, c] =
// min{ i : feats[i, c] == out[v, c] } by atomicMin; pass 2 routes the gradient.
__global__ void scatter_argmax_kernel(const float *__restrict__ feats, const float *__restrict__ out,
                                      const int64_t *__restrict__ ids, int64_t n, int c, int64_t m, int32_t *__restrict__ arg) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  const int64_t row = t / c;
  const int col = (int)(t - row * c);
  const int64_t id = __ldg(ids + row);
  if (id >= 0 && id < m && feats[t] == out[id * c + col]) atomicMin(arg + id * c + col, (int32_t)row);
}

__global__ void scatter_max_bwd_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ arg,
                                       const int64_t *__restrict__ ids, int64_t n, int c, int64_t m,
                                       float *__restrict__ grad_in) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  const int64_t row = t / c;
  const int col = (int)(t - row * c);
  const int64_t id = __ldg(ids + row);
  float g = 0.0f;
  if (id >= 0 && id < m && __ldg(arg + id * c + col) == (int32_t)row) g = grad_out[id * c + col];
  grad_in[t] = g;
}

__global__ void scatter_mean_bwd_kernel(const float *__restrict__ grad_out, const int64_t *__restrict__ ids,
                                        const int32_t *__restrict__ counts, int64_t n, int c, int64_t m,
                                        float *__restrict__ grad_in) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  const int64_t row = t / c;
  const int col = (int)(t - row * c);
  const int64_t id = __ldg(ids + row);
  grad_in[t] = (id >= 0 && id < m) ? grad_out[id * c + col] / (float)max(__ldg(counts + id), 1) : 0.0f;
}

// out[i] = feats[ids[i]] (or 0).  16-byte vectors; one thread per (row, vector).
__global__ void gather_rows_kernel(const uint4 *__restrict__ feats, const int64_t *__restrict__ ids, int64_t n, int vec,
                                   uint4 *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * vec) return;
  const int64_t row = t / vec;
  const int col = (int)(t - row * vec);
  const int64_t id = __ldg(ids + row);
  out[t] = id >= 0 ? __ldg(feats + id * vec + col) : make_uint4(0, 0, 0, 0);
}

__global__ void gather_rows_scalar_kernel(const uint16_t *__restrict__ feats, const int64_t *__restrict__ ids,
                                          int64_t n, int halves, uint16_t *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * halves) return;
  const int64_t row = t / halves;
  const int col = (int)(t - row * halves);
  const int64_t id = __ldg(ids + row);
  out[t] = id >= 0 ? feats[id * halves + col] : (uint16_t)0;
}

}  // namespace os3d

using namespace os3d;

static inline unsigned grid_for(int64_t work, int threads) { return (unsigned)cdiv(work > 0 ? work : 1, threads); }

extern "C" int os3d_scatter_max_f32(const float *feats, const int64_t *ids, int64_t n, int c, float *out, int64_t m,
                                    int fix_empty, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (c <= 0) return OS3D_ERR_BAD_ARG;
  const int64_t total = m * c;
  if (total > 0) fill_f32_kernel<<<grid_for(total / 4 + 1, 256), 256, 0, st>>>(out, total / 4, total, -INFINITY);
  if (n > 0 && m > 0) {
    if (c % 4 == 0) scatter_max_kernel<true><<<grid_for(n * (c / 4), 256), 256, 0, st>>>(feats, ids, n, c, c / 4, out, m);
    else scatter_max_kernel<false><<<grid_for(n * c, 256), 256, 0, st>>>(feats, ids, n, c, c, out, m);
  }
  if (fix_empty && total > 0) fix_empty_kernel<<<grid_for(total, 256), 256, 0, st>>>(out, total);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_max_bf16(const void *feats, const void *ids, int ids_are_i64, int64_t n, int c, float *acc,
                                    void *out, int64_t m, int fix_empty, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (c <= 0 || c % 8) return OS3D_ERR_BAD_ARG;
  const int64_t total = m * c;
  if (total == 0) return 0;
  fill_f32_kernel<<<grid_for(total / 4 + 1, 256), 256, 0, st>>>(acc, total / 4, total, -INFINITY);
  if (n > 0) {
    if (ids_are_i64)
      scatter_max_bf16_kernel<int64_t><<<grid_for(n * (c / 8), 256), 256, 0, st>>>((const uint4 *)feats, (const int64_t *)ids, n,
                                                                                c, c / 8, acc, m);
    else
      scatter_max_bf16_kernel<int32_t><<<grid_for(n * (c / 8), 256), 256, 0, st>>>((const uint4 *)feats, (const int32_t *)ids, n,
                                                                                c, c / 8, acc, m);
  }
  finalize_max_bf16_kernel<<<grid_for(total / 4, 256), 256, 0, st>>>((const float4 *)acc, (uint2 *)out, total / 4, fix_empty);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_max_sorted_scratch(int64_t n, int64_t m, int64_t *temp_bytes) {
  size_t bytes = 0;
  int bits = 1;
  while ((1ll << bits) <= m) ++bits;                                // keys 0..m
  OS3D_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (const int32_t *)nullptr, (int32_t *)nullptr, (int)n, 0, bits));
  *temp_bytes = (int64_t)bytes;
  return 0;
}

extern "C" int os3d_scatter_max_sorted_bf16(const void *feats, const void *ids, int ids_are_i64, int64_t n, int c,
                                           uint32_t *keys, uint32_t *keys_sorted, int32_t *rows, int32_t *rows_sorted,
                                           void *temp, int64_t temp_bytes, void *out, int64_t m, int fix_empty,
                                           void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (c <= 0 || c % 8 || c > 2048 || n > 0x7fffffff || m >= 0x7fffffff) return OS3D_ERR_BAD_ARG;
  const int64_t total = m * c;
  if (total == 0) return 0;
  // rows nothing is scattered to: 0, or bf16 -inf (0xff80) when the caller wants torch_scatter's raw maximum
  if (fix_empty) OS3D_CUDA(cudaMemsetAsync(out, 0, (size_t)total * 2, st));
  else fill_u32_kernel<<<grid_for(total / 8, 256), 256, 0, st>>>((uint4 *)out, total / 8, 0xff80ff80u);
  if (n > 0) {
    int bits = 1;
    while ((1ll << bits) <= m) ++bits;
    voxel_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(ids_are_i64 ? (const int64_t *)ids : nullptr,
                                                        ids_are_i64 ? nullptr : (const int32_t *)ids, n, m, keys, rows);
    size_t bytes = (size_t)temp_bytes;
    OS3D_CUDA(cub::DeviceRadixSort::SortPairs(temp, bytes, keys, keys_sorted, rows, rows_sorted, (int)n, 0, bits, st));
    const int cv = c / 8, per_block = 256 / cv;
    run_max_bf16_kernel<<<grid_for(n, per_block), 256, 0, st>>>((const uint4 *)feats, keys_sorted, rows_sorted, n, cv, m,
                                                               (uint4 *)out);
  }
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_mean_f32(const float *feats, const int64_t *ids, int64_t n, int c, float *out,
                                     int32_t *counts, const int32_t *counts_in, int64_t m, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (c <= 0 || (!counts && !counts_in)) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  OS3D_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(m * c), st));
  if (counts) OS3D_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)m, st));
  if (n > 0) {
    if (c % 4 == 0 && c / 4 <= 256 && m <= 64 && !counts_in)
      scatter_add_runs_kernel<<<grid_for(n, kSmallMRows), 256, 0, st>>>(feats, ids, n, c, c / 4, out, counts, m);
    else if (c % 4 == 0)
      scatter_add_kernel<true><<<grid_for(n * (c / 4), 256), 256, 0, st>>>(feats, ids, n, c, c / 4, out, counts, counts_in, m);
    else
      scatter_add_kernel<false><<<grid_for(n * c, 256), 256, 0, st>>>(feats, ids, n, c, c, out, counts, counts_in, m);
  }
  if (!counts_in) mean_normalize_kernel<<<grid_for(m * c, 256), 256, 0, st>>>(out, counts, m, c);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_mean_small_bf16(const void *feats, const void *ids, int ids_are_i64, int64_t n, int c, float *out,
                                           int32_t *counts, int64_t m, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (c <= 0 || c % 8 || c / 8 > 256 || !counts || m > 64) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  OS3D_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(m * c), st));
  OS3D_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)m, st));
  if (n > 0) {
    if (ids_are_i64)
      scatter_add_runs_bf16_kernel<int64_t><<<grid_for(n, kSmallMRows), 256, 0, st>>>((const uint4 *)feats, (const int64_t *)ids,
                                                                                   n, c, c / 8, out, counts, m);
    else
      scatter_add_runs_bf16_kernel<int32_t><<<grid_for(n, kSmallMRows), 256, 0, st>>>((const uint4 *)feats, (const int32_t *)ids,
                                                                                   n, c, c / 8, out, counts, m);
  }
  mean_normalize_kernel<<<grid_for(m * c, 256), 256, 0, st>>>(out, counts, m, c);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_max_bwd_f32(const float *grad_out, const float *feats, const float *out, const int64_t *ids,
                                        int64_t n, int c, int64_t m, int32_t *arg, float *grad_in, void *stream) {
  if (n == 0) return 0;
  if (n > 0x7fffffff || !arg) return OS3D_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  OS3D_CUDA(cudaMemsetAsync(arg, 0x7f, sizeof(int32_t) * (size_t)m * c, st));        // 0x7f7f7f7f > any row index
  scatter_argmax_kernel<<<grid_for(n * c, 256), 256, 0, st>>>(feats, out, ids, n, c, m, arg);
  scatter_max_bwd_kernel<<<grid_for(n * c, 256), 256, 0, st>>>(grad_out, arg, ids, n, c, m, grad_in);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_mean_bwd_f32(const float *grad_out, const int64_t *ids, const int32_t *counts, int64_t n,
                                         int c, int64_t m, float *grad_in, void *stream) {
  if (n == 0) return 0;
  scatter_mean_bwd_kernel<<<grid_for(n * c, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, ids, counts, n, c, m,
                                                                                   grad_in);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_gather_rows(const void *feats, const int64_t *ids, int64_t n, int c, int elem_size, void *out,
                                void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return 0;
  const int64_t row_bytes = (int64_t)c * elem_size;
  if (row_bytes % 16 == 0) {
    const int vec = (int)(row_bytes / 16);
    gather_rows_kernel<<<grid_for(n * vec, 256), 256, 0, st>>>((const uint4 *)feats, ids, n, vec, (uint4 *)out);
  } else if (row_bytes % 2 == 0) {
    const int halves = (int)(row_bytes / 2);
    gather_rows_scalar_kernel<<<grid_for(n * halves, 256), 256, 0, st>>>((const uint16_t *)feats, ids, n, halves,
                                                                         (uint16_t *)out);
  } else {
    return OS3D_ERR_BAD_ARG;
  }
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_scatter_add_rows_f32(const float *grad_out, const int64_t *ids, int64_t n, int c, float *grad_feats,
                                         int64_t m, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) return 0;
  OS3D_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(float) * (size_t)(m * c), st));
  if (n > 0) {
    if (c % 4 == 0)
      scatter_add_kernel<true><<<grid_for(n * (c / 4), 256), 256, 0, st>>>(grad_out, ids, n, c, c / 4, grad_feats, nullptr, nullptr, m);
    else
      scatter_add_kernel<false><<<grid_for(n * c, 256), 256, 0, st>>>(grad_out, ids, n, c, c, grad_feats, nullptr, nullptr, m);
  }
  OS3D_LAUNCH_CHECK();
  return 0;
}
