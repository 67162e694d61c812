// scatter.cu -- stage 1b/1c: point->voxel scatter_max / scatter_mean and voxel->point gather.
// replaces torch_scatter.scatter as called by VFE.forward (seg3d/models/voxel_encoders/vfe.py:24-25), the in-tree
// voxel_pooling kernels (seg3d/ops/voxel_pooling/src/voxel_pooling_cuda.cu:11-41, launched <<<N, C>>>) and
// voxel_to_point (seg3d/ops/voxel_to_point/voxel_to_point.py:5-17).
//
// All kernels are HBM streaming kernels: a thread owns one float4 of one row, a warp covers 512 contiguous bytes,
// index loads are one per row (broadcast within the row's lanes), reductions go to L2 as fire-and-forget RED ops
// (vector red.global.add.v4.f32 for the mean; signed-max / unsigned-min integer REDs for the max, which makes the
// max bit-exact and order independent).
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
