// norm.cu -- stage 4 glue: fused residual + LayerNorm for the SWFormer encoder layer (post-norm):
//   out = resid + LayerNorm(x) * w + b          (seg3d/models/layers/point_transformer_layer.py:288-298:
//   x = shortcut + drop_path(norm1(attn(x)));  x = x + drop_path(norm2(mlp(x))))
// Rows are short (C = 48..384), so a row is owned by a sub-warp group of G lanes (G = 8/16/32), every lane moves
// 16-byte vectors, mean / variance are reduced with shuffles inside the group: one pass over HBM, 3 tensors touched
// (x, resid, out) instead of the 5 passes of layer_norm + add, and no one-CTA-per-row launch.
#include "common.cuh"

namespace os3d {

template <typename T> struct Vec8;   // 8 elements = one 16-byte (bf16) or two 16-byte (f32) accesses
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float *v) {
    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float *p, float *v) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float *p, const float *v) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// G lanes per row, each lane owns chunks lane, lane+G, ... (<= kMaxChunks of 8 elements)
template <typename T, int G, int kMaxChunks>
__global__ void __launch_bounds__(256) layernorm_residual_kernel(const T *__restrict__ x, const T *__restrict__ resid,
                                                                  const float *__restrict__ w, const float *__restrict__ b,
                                                                  int64_t m, int c, float eps, T *__restrict__ out) {
  const int lane = threadIdx.x % G;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int chunks = c / 8;
  const bool row_ok = row < m;
  float v[kMaxChunks][8];
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ch = lane + k * G;
    if (row_ok && ch < chunks) {
      Vec8<T>::load(x + row * c + ch * 8, v[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[k][i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k][i] = 0.0f;
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)c;
  float var = 0.0f;
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ch = lane + k * G;
    if (ch < chunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[k][i] - mean; var = fmaf(d, d, var); }
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)c + eps);
  if (!row_ok) return;
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ch = lane + k * G;
    if (ch < chunks) {
      float r[8], y[8];
      if (resid) Vec8<T>::load(resid + row * c + ch * 8, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = ch * 8 + i;
        y[i] = (v[k][i] - mean) * rstd * __ldg(w + col) + __ldg(b + col);
        if (resid) y[i] += r[i];
      }
      Vec8<T>::store(out + row * c + ch * 8, y);
    }
  }
}

template <typename T>
static int launch_ln(const T *x, const T *resid, const float *w, const float *b, int64_t m, int c, float eps, T *out,
                     cudaStream_t st) {
  const int chunks = c / 8;
#define OS3D_LN(G, K)                                                                                                \
  layernorm_residual_kernel<T, G, K><<<(unsigned)cdiv(m * G, 256), 256, 0, st>>>(x, resid, w, b, m, c, eps, out)
  if (chunks <= 8) OS3D_LN(8, 1);
  else if (chunks <= 16) OS3D_LN(16, 1);
  else if (chunks <= 32) OS3D_LN(32, 1);
  else if (chunks <= 64) OS3D_LN(32, 2);
  else if (chunks <= 128) OS3D_LN(32, 4);
  else return OS3D_ERR_BAD_ARG;
#undef OS3D_LN
  return 0;
}

}  // namespace os3d

using namespace os3d;

// out[r, :] = x[r, :] + table[idx[r], :]   (q = k = x + pos: the position embedding takes only window-volume many values,
// so it is a small L2-resident table indexed by the in-window position instead of an [M, C] tensor in HBM)
template <typename T>
__global__ void __launch_bounds__(256) add_table_rows_kernel(const T *__restrict__ x, const T *__restrict__ table,
                                                              const int32_t *__restrict__ idx, int64_t m, int chunks,
                                                              T *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * chunks) return;
  const int64_t r = t / chunks;
  const int c = (int)(t - r * chunks) * 8;
  const int64_t c_total = (int64_t)chunks * 8;
  float a[8], b[8];
  Vec8<T>::load(x + r * c_total + c, a);
  Vec8<T>::load(table + (int64_t)__ldg(idx + r) * c_total + c, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] += b[i];
  Vec8<T>::store(out + r * c_total + c, a);
}

// GELU (erf form, nn.GELU() default) on bf16 rows, 16 bytes per thread.  erff() costs ~20 instructions per element, which
// makes torch's kernel compute-bound on B200 (3.3 TB/s); here erf comes from Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7: at most one bf16 ulp in the result, 98.7 % bit-identical): erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),
// t = 1 / (1 + p z), z >= 0 -- one reciprocal, one ex2, six FMAs.
__global__ void __launch_bounds__(256) gelu_bf16_kernel(const __nv_bfloat16 *__restrict__ x, int64_t n8,
                                                        __nv_bfloat16 *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n8) return;
  float v[8];
  Vec8<__nv_bfloat16>::load(x + t * 8, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = gelu_erf_fast(v[i]);
  Vec8<__nv_bfloat16>::store(out + t * 8, v);
}

extern "C" int os3d_gelu_bf16(const void *x, int64_t n, void *out, void *stream) {
  if (n < 0 || n % 8) return OS3D_ERR_BAD_ARG;
  if (n == 0) return 0;
  gelu_bf16_kernel<<<(unsigned)cdiv(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, n / 8,
                                                                                (__nv_bfloat16 *)out);
  OS3D_LAUNCH_CHECK();
  return 0;
}

// out[r, :] = x[r, :] * (bias + table[idx[r], :])   -- the squeeze-excite gate of FlattenSELayer applied per point:
// x * gate[batch_idx] (bias 0) or, with the residual the segmentor adds, x + x * gate[batch_idx] (bias 1)
// (se_layer.py:24-30, segformer.py:134).  idx: int64 batch index per row; the table has one row per frame.
template <typename T>
__global__ void __launch_bounds__(256) scale_rows_by_table_kernel(const T *__restrict__ x, const float *__restrict__ table,
                                                                   const int64_t *__restrict__ idx, int64_t m, int chunks,
                                                                   float bias, T *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * chunks) return;
  const int64_t r = t / chunks;
  const int c = (int)(t - r * chunks) * 8;
  const int64_t c_total = (int64_t)chunks * 8;
  float a[8], g[8];
  Vec8<T>::load(x + r * c_total + c, a);
  Vec8<float>::load(table + __ldg(idx + r) * c_total + c, g);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] *= bias + g[i];
  Vec8<T>::store(out + r * c_total + c, a);
}

extern "C" int os3d_scale_rows_by_table(const void *x, const float *table, const int64_t *idx, int64_t m, int c, float bias,
                                        int elem_size, void *out, void *stream) {
  if (c <= 0 || c % 8 || (elem_size != 2 && elem_size != 4)) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  const int chunks = c / 8;
  const unsigned g = (unsigned)cdiv(m * chunks, 256);
  if (elem_size == 2)
    scale_rows_by_table_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)x, table, idx, m, chunks, bias, (__nv_bfloat16 *)out);
  else
    scale_rows_by_table_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float *)x, table, idx, m, chunks, bias,
                                                                            (float *)out);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_add_table_rows(const void *x, const void *table, const int32_t *idx, int64_t m, int c, int elem_size,
                                   void *out, void *stream) {
  if (c <= 0 || c % 8 || (elem_size != 2 && elem_size != 4)) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  const int chunks = c / 8;
  const unsigned g = (unsigned)cdiv(m * chunks, 256);
  if (elem_size == 2)
    add_table_rows_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)x, (const __nv_bfloat16 *)table, idx, m, chunks, (__nv_bfloat16 *)out);
  else
    add_table_rows_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float *)x, (const float *)table, idx, m,
                                                                       chunks, (float *)out);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_layernorm_residual(const void *x, const void *resid, const float *w, const float *b, int64_t m, int c,
                                       float eps, int elem_size, void *out, void *stream) {
  if (c <= 0 || c % 8) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  int rc;
  if (elem_size == 4)
    rc = launch_ln<float>((const float *)x, (const float *)resid, w, b, m, c, eps, (float *)out, (cudaStream_t)stream);
  else if (elem_size == 2)
    rc = launch_ln<__nv_bfloat16>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)resid, w, b, m, c, eps,
                                  (__nv_bfloat16 *)out, (cudaStream_t)stream);
  else
    return OS3D_ERR_BAD_ARG;
  if (rc) return rc;
  OS3D_LAUNCH_CHECK();
  return 0;
}
