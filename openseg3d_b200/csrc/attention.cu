// attention.cu -- stage 4, floating-point part: variable-length cosine window attention.
// replaces flat2window -> CosineMultiheadAttention core (_scaled_cosine_attention,
// seg3d/models/layers/cosine_msa.py:115-177) -> window2flat (seg3d/utils/swformer_utils.py:34-85) as driven by
// WindowAttention.forward (seg3d/models/layers/point_transformer_layer.py:233-258).
//
// The reference pads every window to its batching level's max_tokens, materialises [R*h, T, T] scores plus a
// head-averaged copy it throws away, and masks the padding with -inf.  Projections are per token, padding rows are
// discarded, so the same result is obtained on the FLAT token list: this kernel walks the window segments of
// os3d_window_partition directly -- no padding, no masks, no score tensor in HBM, 3.8x fewer QK^T/PV FLOPs
// (SURVEY.md §7.3 item 7).
//
// v1 (this file): FP32 SIMT.  A warp owns (window, head, 32-query chunk); lane = query; keys/values of the window
// are streamed with warp-broadcast vector loads (all lanes read the same K/V row -> one transaction), online softmax
// in registers.  Work items are enumerated level by level from the device-side window counts, heaviest level
// first, so no host round trip is needed.
#include "common.cuh"

namespace os3d {

template <typename T> struct Ld;
template <> struct Ld<float> {
  static __device__ __forceinline__ float get(const float *p) { return __ldg(p); }
};
template <> struct Ld<__nv_bfloat16> {
  static __device__ __forceinline__ float get(const __nv_bfloat16 *p) {
    return __bfloat162float(__ldg(p));
  }
};

// Load D contiguous elements (one head slice) as floats with the widest vector the slice alignment allows
// (slice offset and row pitch are multiples of D*sizeof(T); base pointers are 16-byte aligned).
template <typename T, int D>
__device__ __forceinline__ void load_slice(const T *__restrict__ p, float *dst) {
  constexpr int kBytes = D * (int)sizeof(T);
  constexpr int kVec = (kBytes % 16 == 0) ? 16 : (kBytes % 8 == 0) ? 8 : (kBytes % 4 == 0) ? 4 : (int)sizeof(T);
  constexpr int kPer = kVec / (int)sizeof(T);
  if constexpr (kVec == (int)sizeof(T)) {
#pragma unroll
    for (int i = 0; i < D; ++i) dst[i] = Ld<T>::get(p + i);
  } else {
#pragma unroll
    for (int i = 0; i < D / kPer; ++i) {
      uint32_t w[4];
      if constexpr (kVec == 16) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
      } else if constexpr (kVec == 8) {
        const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p) + i);
        w[0] = u.x; w[1] = u.y;
      } else {
        w[0] = __ldg(reinterpret_cast<const uint32_t *>(p) + i);
      }
#pragma unroll
      for (int j = 0; j < kVec / 4; ++j) {
        if constexpr (sizeof(T) == 4) {
          dst[i * kPer + j] = __uint_as_float(w[j]);
        } else {  // two bf16 per word: bf16 -> f32 is a 16-bit shift
          dst[i * kPer + 2 * j] = __uint_as_float(w[j] << 16);
          dst[i * kPer + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
      }
    }
  }
}

// In-place L2 normalisation of every head slice of q and k (F.normalize, eps 1e-12; cosine_msa.py:152-153).
template <typename T>
__global__ void qk_normalize_kernel(T *__restrict__ q, T *__restrict__ k, int64_t ld, int64_t m, int heads, int d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * heads * 2) return;
  const int which = (int)(t & 1);
  const int64_t r = (t >> 1) / heads;
  const int h = (int)((t >> 1) - r * heads);
  T *p = (which ? k : q) + r * ld + h * d;
  float ss = 0.0f;
  for (int i = 0; i < d; ++i) {
    const float v = Ld<T>::get(p + i);
    ss = fmaf(v, v, ss);
  }
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int i = 0; i < d; ++i) {
    const float v = Ld<T>::get(p + i) * inv;
    if constexpr (sizeof(T) == 4) p[i] = v; else p[i] = __float2bfloat16(v);
  }
}

struct AttnLevels {
  int chunks[OS3D_MAX_LEVELS];  // ceil(max_tokens / 32) per level
};

template <typename T, int D>
__global__ void __launch_bounds__(128) window_attention_kernel(const T *__restrict__ q, const T *__restrict__ k,
                                                                const T *__restrict__ v, int64_t ld, int64_t ldv, int heads,
                                                                const int32_t *__restrict__ order,
                                                                const int32_t *__restrict__ seg_start,
                                                                const int32_t *__restrict__ seg_len,
                                                                const int32_t *__restrict__ level_info,
                                                                AttnLevels lv, const float *__restrict__ tau,
                                                                float tau_min, T *__restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // softmax in base 2: scores * log2(e) / max(tau, tau_min)
  const float scale = 1.4426950408889634f / fmaxf(__ldg(tau), tau_min);

  int64_t items_before = 0;
  for (int lvl = OS3D_MAX_LEVELS - 1; lvl >= 0; --lvl) {  // heaviest windows first
    const int n_windows = __ldg(level_info + lvl);
    const int first = __ldg(level_info + 4 + lvl);
    const int chunks = lv.chunks[lvl];
    const int64_t items = (int64_t)n_windows * heads * chunks;
    // first item of this level owned by this warp
    int64_t it = warp - (items_before % n_warps);
    if (it < 0) it += n_warps;
    for (; it < items; it += n_warps) {
      const int chunk = (int)(it % chunks);
      const int h = (int)((it / chunks) % heads);
      const int win = first + (int)(it / ((int64_t)chunks * heads));
      const int n = __ldg(seg_len + win);
      if (chunk * 32 >= n) continue;
      const int32_t *seg = order + __ldg(seg_start + win);
      const int qi = chunk * 32 + lane;
      const bool active = qi < n;
      const int32_t qrow = __ldg(seg + (active ? qi : 0));
      float qr[D], acc[D];
      {
        load_slice<T, D>(q + (int64_t)qrow * ld + h * D, qr);
#pragma unroll
        for (int i = 0; i < D; ++i) { qr[i] *= scale; acc[i] = 0.0f; }
      }
      float mx = -INFINITY, sum = 0.0f;
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int32_t my_key = j0 + lane < n ? __ldg(seg + j0 + lane) : 0;
        const int jn = min(32, n - j0);
        for (int jj = 0; jj < jn; ++jj) {
          const int32_t krow = __shfl_sync(0xffffffffu, my_key, jj);
          float kr[D], vr[D];
          load_slice<T, D>(k + (int64_t)krow * ld + h * D, kr);
          load_slice<T, D>(v + (int64_t)krow * ldv + h * D, vr);
          float s = 0.0f;
#pragma unroll
          for (int i = 0; i < D; ++i) s = fmaf(qr[i], kr[i], s);
          if (s > mx) {  // rescale lazily: only when the running max moves
            const float corr = exp2f(mx - s);
            sum *= corr;
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] *= corr;
            mx = s;
          }
          const float p = exp2f(s - mx);
          sum += p;
#pragma unroll
          for (int i = 0; i < D; ++i) acc[i] = fmaf(p, vr[i], acc[i]);
        }
      }
      if (active) {
        const float inv = 1.0f / sum;
        T *op = out + (int64_t)qrow * ldo + h * D;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          if constexpr (sizeof(T) == 4) op[i] = acc[i] * inv; else op[i] = __float2bfloat16(acc[i] * inv);
        }
      }
    }
    items_before += items;
  }
}

template <typename T>
static int launch_attention(const T *q, const T *k, const T *v, int64_t ld, int64_t ldv, int64_t m, int c, int heads,
                            const int32_t *order, const int32_t *seg_start, const int32_t *seg_len,
                            const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min, T *out,
                            cudaStream_t st) {
  const int d = c / heads;
  AttnLevels lv;
  for (int l = 0; l < OS3D_MAX_LEVELS; ++l) lv.chunks[l] = (lvl_tokens[l] + 31) / 32;
  // enough warps to cover the machine several times over; every warp strides over the device-side item list
  const int64_t want_warps = (m / 4 + 1) * heads;
  const int64_t max_blocks = 148 * 16;
  const unsigned blocks = (unsigned)max((int64_t)1, min(max_blocks, cdiv(want_warps, 4)));
#define OS3D_ATTN_CASE(DD)                                                                                          \
  case DD:                                                                                                          \
    window_attention_kernel<T, DD><<<blocks, 128, 0, st>>>(q, k, v, ld, ldv, heads, order, seg_start, seg_len, level_info, \
                                                           lv, tau, tau_min, out, (int64_t)c);                      \
    break;
  switch (d) {
    OS3D_ATTN_CASE(3)
    OS3D_ATTN_CASE(6)
    OS3D_ATTN_CASE(12)
    OS3D_ATTN_CASE(16)
    OS3D_ATTN_CASE(24)
    OS3D_ATTN_CASE(32)
    OS3D_ATTN_CASE(48)
    OS3D_ATTN_CASE(64)
    default:
      return OS3D_ERR_BAD_ARG;
  }
#undef OS3D_ATTN_CASE
  return 0;
}

}  // namespace os3d

using namespace os3d;

extern "C" int os3d_qk_normalize(void *q, void *k, int64_t ld, int64_t m, int c, int heads, int elem_size, void *stream) {
  if (m == 0) return 0;
  if (heads <= 0 || c % heads) return OS3D_ERR_BAD_ARG;
  const unsigned g = (unsigned)cdiv(m * heads * 2, 256);
  if (elem_size == 4)
    qk_normalize_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((float *)q, (float *)k, ld, m, heads, c / heads);
  else if (elem_size == 2)
    qk_normalize_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16 *)q, (__nv_bfloat16 *)k, ld, m,
                                                                          heads, c / heads);
  else
    return OS3D_ERR_BAD_ARG;
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_window_attention(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m, int c,
                                     int heads, const int32_t *order, const int32_t *seg_start, const int32_t *seg_len,
                                     const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min,
                                     int elem_size, void *out, void *stream) {
  if (m == 0) return 0;
  if (heads <= 0 || c % heads) return OS3D_ERR_BAD_ARG;
  int rc;
  if (elem_size == 4)
    rc = launch_attention<float>((const float *)q, (const float *)k, (const float *)v, ld, ldv, m, c, heads, order, seg_start,
                                 seg_len, level_info, lvl_tokens, tau, tau_min, (float *)out, (cudaStream_t)stream);
  else if (elem_size == 2)
    rc = launch_attention<__nv_bfloat16>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v, ld,
                                         ldv, m, c, heads, order, seg_start, seg_len, level_info, lvl_tokens, tau, tau_min,
                                         (__nv_bfloat16 *)out, (cudaStream_t)stream);
  else
    return OS3D_ERR_BAD_ARG;
  if (rc) return rc;
  OS3D_LAUNCH_CHECK();
  return 0;
}
