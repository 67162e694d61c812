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

// The same with the head width a compile-time constant: one thread per (row, q | k, head) slice, the slice moved with the
// widest vectors its alignment allows (16 bytes for the bf16 head-padded layouts: 1 GB read + written in ~0.2 ms where the
// element-wise kernel above needed 1.2 ms).
template <typename T, int D>
__global__ void __launch_bounds__(256) qk_normalize_vec_kernel(T *__restrict__ q, T *__restrict__ k, int64_t ld, int64_t m,
                                                               int heads) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * heads * 2) return;
  // consecutive threads take consecutive head slices of one row (coalesced), q rows then k rows per row
  const int64_t r = t / (2 * heads);
  const int rem = (int)(t - r * 2 * heads);
  const int which = rem >= heads, h = which ? rem - heads : rem;
  T *p = (which ? k : q) + r * ld + h * D;
  float f[D];
  load_slice<T, D>(p, f);
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < D; ++i) ss = fmaf(f[i], f[i], ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  constexpr int kBytes = D * (int)sizeof(T);
  constexpr int kVec = (kBytes % 16 == 0) ? 16 : (kBytes % 8 == 0) ? 8 : (kBytes % 4 == 0) ? 4 : (int)sizeof(T);
  constexpr int kPer = kVec / (int)sizeof(T);
  if constexpr (kVec == (int)sizeof(T)) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if constexpr (sizeof(T) == 4) p[i] = f[i] * inv; else p[i] = __float2bfloat16(f[i] * inv);
    }
  } else {
#pragma unroll
    for (int i = 0; i < D / kPer; ++i) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < kVec / 4; ++j) {
        if constexpr (sizeof(T) == 4) {
          w[j] = __float_as_uint(f[i * kPer + j] * inv);
        } else {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(f[i * kPer + 2 * j] * inv, f[i * kPer + 2 * j + 1] * inv);
          w[j] = *reinterpret_cast<const uint32_t *>(&hh);
        }
      }
      if constexpr (kVec == 16) reinterpret_cast<uint4 *>(p)[i] = make_uint4(w[0], w[1], w[2], w[3]);
      else if constexpr (kVec == 8) reinterpret_cast<uint2 *>(p)[i] = make_uint2(w[0], w[1]);
      else reinterpret_cast<uint32_t *>(p)[i] = w[0];
    }
  }
}

// Attention dropout (training; cosine_msa.py:173-174): keep-mask of weight (head, query row, key row) from a
// counter-based hash of (seed, head, rows), so forward and backward regenerate the same mask without storing it.  Not
// torch's Philox stream: a run is reproducible from its seed, not bit-identical to the reference's dropout.
__device__ __forceinline__ float drop_keep(uint64_t seed, int h, int32_t qrow, int32_t krow, float drop_p, float keep_scale) {
  if (drop_p <= 0.0f) return 1.0f;
  const uint64_t x = mix64(seed ^ ((uint64_t)(uint32_t)qrow << 32 | (uint32_t)krow) ^ ((uint64_t)h * 0x9e3779b97f4a7c15ULL));
  const float u = (float)(x >> 40) * (1.0f / 16777216.0f);          // 24 random bits -> [0, 1)
  return u >= drop_p ? keep_scale : 0.0f;
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
                                                                float tau_min, float drop_p, uint64_t seed,
                                                                T *__restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // softmax in base 2: scores * log2(e) / max(tau, tau_min)
  const float scale = 1.4426950408889634f / fmaxf(__ldg(tau), tau_min);
  const float keep_scale = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;

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
          sum += p;                                                     // the softmax denominator ignores dropout
          const float pd = p * drop_keep(seed, h, qrow, krow, drop_p, keep_scale);
#pragma unroll
          for (int i = 0; i < D; ++i) acc[i] = fmaf(pd, vr[i], acc[i]);
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
                            const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min,
                            float drop_p, uint64_t seed, T *out, cudaStream_t st) {
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
                                                           lv, tau, tau_min, drop_p, seed, out, (int64_t)c);        \
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

// ---- backward ----------------------------------------------------------------------------------------------------
// With z_ij = (q_i . k_j) / tau', P = softmax_j(z), P' = dropout(P), O = P' V:
//   D_i   = dO_i . O_i  (= sum_j P_ij dP_ij),   dP_ij = keep_ij (dO_i . v_j),   dz_ij = P_ij (dP_ij - D_i)
//   dq_i  = sum_j dz_ij k_j / tau'      dk_j = sum_i dz_ij q_i / tau'      dv_j = sum_i P'_ij dO_i
//   d(1/tau') = sum_ij dz_ij (q_i . k_j)
// Pass A (lane = query): row statistics (max, sum, D) into `stats`, dq, the 1/tau' gradient.
// Pass B (lane = key)  : dk, dv accumulate in the key lane's registers while the window's queries are broadcast --
//                        no cross-lane reductions, no atomics on dk / dv.
template <typename T, int D>
__device__ __forceinline__ void store_slice(T *p, const float *v) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if constexpr (sizeof(T) == 4) p[i] = v[i]; else p[i] = __float2bfloat16(v[i]);
  }
}

template <typename T, int D, int PASS>
__global__ void __launch_bounds__(128) window_attention_bwd_kernel(
    const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v, const T *__restrict__ o,
    const T *__restrict__ go, int64_t ld, int heads, const int32_t *__restrict__ order,
    const int32_t *__restrict__ seg_start, const int32_t *__restrict__ seg_len, const int32_t *__restrict__ level_info,
    AttnLevels lv, const float *__restrict__ tau, float tau_min, float drop_p, uint64_t seed, float *__restrict__ stats,
    T *__restrict__ gq, T *__restrict__ gk, T *__restrict__ gv, float *__restrict__ g_inv_tau) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_tau = 1.0f / fmaxf(__ldg(tau), tau_min);
  const float scale2 = 1.4426950408889634f * inv_tau;
  const float keep_scale = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  float g_scale_acc = 0.0f;

  int64_t items_before = 0;
  for (int lvl = OS3D_MAX_LEVELS - 1; lvl >= 0; --lvl) {
    const int n_windows = __ldg(level_info + lvl);
    const int first = __ldg(level_info + 4 + lvl);
    const int chunks = lv.chunks[lvl];
    const int64_t items = (int64_t)n_windows * heads * chunks;
    int64_t it = warp - (items_before % n_warps);
    if (it < 0) it += n_warps;
    for (; it < items; it += n_warps) {
      const int chunk = (int)(it % chunks);
      const int h = (int)((it / chunks) % heads);
      const int win = first + (int)(it / ((int64_t)chunks * heads));
      const int n = __ldg(seg_len + win);
      if (chunk * 32 >= n) continue;
      const int32_t *seg = order + __ldg(seg_start + win);
      const int mi = chunk * 32 + lane;                 // this lane's query (pass A) / key (pass B) inside the window
      const bool active = mi < n;
      const int32_t myrow = __ldg(seg + (active ? mi : 0));
      if constexpr (PASS == 0) {
        float qr[D], gor[D], acc[D];
        load_slice<T, D>(q + (int64_t)myrow * ld + h * D, qr);
        load_slice<T, D>(go + (int64_t)myrow * ld + h * D, gor);
        float dsum = 0.0f;
        {
          float orow[D];
          load_slice<T, D>(o + (int64_t)myrow * ld + h * D, orow);
#pragma unroll
          for (int i = 0; i < D; ++i) { dsum = fmaf(gor[i], orow[i], dsum); acc[i] = 0.0f; }
        }
        float mx = -INFINITY, sum = 0.0f;
        for (int j0 = 0; j0 < n; j0 += 32) {            // row statistics
          const int32_t my_key = j0 + lane < n ? __ldg(seg + j0 + lane) : 0;
          const int jn = min(32, n - j0);
          for (int jj = 0; jj < jn; ++jj) {
            const int32_t krow = __shfl_sync(0xffffffffu, my_key, jj);
            float kr[D];
            load_slice<T, D>(k + (int64_t)krow * ld + h * D, kr);
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < D; ++i) s = fmaf(qr[i], kr[i], s);
            s *= scale2;
            const float mn = fmaxf(mx, s);
            sum = sum * exp2f(mx - mn) + exp2f(s - mn);
            mx = mn;
          }
        }
        const float inv_l = 1.0f / sum;
        for (int j0 = 0; j0 < n; j0 += 32) {            // dq, d(1/tau')
          const int32_t my_key = j0 + lane < n ? __ldg(seg + j0 + lane) : 0;
          const int jn = min(32, n - j0);
          for (int jj = 0; jj < jn; ++jj) {
            const int32_t krow = __shfl_sync(0xffffffffu, my_key, jj);
            float kr[D], vr[D];
            load_slice<T, D>(k + (int64_t)krow * ld + h * D, kr);
            load_slice<T, D>(v + (int64_t)krow * ld + h * D, vr);
            float s = 0.0f, dp = 0.0f;
#pragma unroll
            for (int i = 0; i < D; ++i) { s = fmaf(qr[i], kr[i], s); dp = fmaf(gor[i], vr[i], dp); }
            const float p = exp2f(s * scale2 - mx) * inv_l;
            const float dz = p * (dp * drop_keep(seed, h, myrow, krow, drop_p, keep_scale) - dsum);
            if (active) g_scale_acc = fmaf(dz, s, g_scale_acc);
            const float w = dz * inv_tau;
#pragma unroll
            for (int i = 0; i < D; ++i) acc[i] = fmaf(w, kr[i], acc[i]);
          }
        }
        if (active) {
          store_slice<T, D>(gq + (int64_t)myrow * ld + h * D, acc);
          float *st = stats + ((int64_t)myrow * heads + h) * 3;
          st[0] = mx; st[1] = inv_l; st[2] = dsum;
        }
      } else {
        float kr[D], vr[D], dk[D], dv[D];
        load_slice<T, D>(k + (int64_t)myrow * ld + h * D, kr);
        load_slice<T, D>(v + (int64_t)myrow * ld + h * D, vr);
#pragma unroll
        for (int i = 0; i < D; ++i) { dk[i] = 0.0f; dv[i] = 0.0f; }
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int32_t my_q = i0 + lane < n ? __ldg(seg + i0 + lane) : 0;
          const int in = min(32, n - i0);
          for (int ii = 0; ii < in; ++ii) {
            const int32_t qrow = __shfl_sync(0xffffffffu, my_q, ii);
            float qr[D], gor[D];
            load_slice<T, D>(q + (int64_t)qrow * ld + h * D, qr);
            load_slice<T, D>(go + (int64_t)qrow * ld + h * D, gor);
            const float *st = stats + ((int64_t)qrow * heads + h) * 3;
            const float mx = __ldg(st), inv_l = __ldg(st + 1), dsum = __ldg(st + 2);
            float s = 0.0f, dp = 0.0f;
#pragma unroll
            for (int i = 0; i < D; ++i) { s = fmaf(qr[i], kr[i], s); dp = fmaf(gor[i], vr[i], dp); }
            const float p = exp2f(s * scale2 - mx) * inv_l;
            const float keep = drop_keep(seed, h, qrow, myrow, drop_p, keep_scale);
            const float w = p * (dp * keep - dsum) * inv_tau;
            const float pk = p * keep;
#pragma unroll
            for (int i = 0; i < D; ++i) { dk[i] = fmaf(w, qr[i], dk[i]); dv[i] = fmaf(pk, gor[i], dv[i]); }
          }
        }
        if (active) {
          store_slice<T, D>(gk + (int64_t)myrow * ld + h * D, dk);
          store_slice<T, D>(gv + (int64_t)myrow * ld + h * D, dv);
        }
      }
    }
    items_before += items;
  }
  if constexpr (PASS == 0) {
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) g_scale_acc += __shfl_xor_sync(0xffffffffu, g_scale_acc, o2);
    if (lane == 0 && g_scale_acc != 0.0f) atomicAdd(g_inv_tau, g_scale_acc);
  }
}

template <typename T>
static int launch_attention_bwd(const T *q, const T *k, const T *v, const T *o, const T *go, int64_t m, int c, int heads,
                                const int32_t *order, const int32_t *seg_start, const int32_t *seg_len,
                                const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min,
                                float drop_p, uint64_t seed, float *stats, T *gq, T *gk, T *gv, float *g_inv_tau,
                                cudaStream_t st) {
  const int d = c / heads;
  AttnLevels lv;
  for (int l = 0; l < OS3D_MAX_LEVELS; ++l) lv.chunks[l] = (lvl_tokens[l] + 31) / 32;
  const int64_t want_warps = (m / 4 + 1) * heads;
  const unsigned blocks = (unsigned)max((int64_t)1, min((int64_t)148 * 16, cdiv(want_warps, 4)));
#define OS3D_ATTN_BWD_CASE(DD)                                                                                          \
  case DD:                                                                                                              \
    window_attention_bwd_kernel<T, DD, 0><<<blocks, 128, 0, st>>>(q, k, v, o, go, (int64_t)c, heads, order, seg_start,  \
        seg_len, level_info, lv, tau, tau_min, drop_p, seed, stats, gq, gk, gv, g_inv_tau);                             \
    window_attention_bwd_kernel<T, DD, 1><<<blocks, 128, 0, st>>>(q, k, v, o, go, (int64_t)c, heads, order, seg_start,  \
        seg_len, level_info, lv, tau, tau_min, drop_p, seed, stats, gq, gk, gv, g_inv_tau);                             \
    break;
  switch (d) {
    OS3D_ATTN_BWD_CASE(3)
    OS3D_ATTN_BWD_CASE(6)
    OS3D_ATTN_BWD_CASE(12)
    OS3D_ATTN_BWD_CASE(16)
    OS3D_ATTN_BWD_CASE(24)
    OS3D_ATTN_BWD_CASE(32)
    OS3D_ATTN_BWD_CASE(48)
    OS3D_ATTN_BWD_CASE(64)
    default:
      return OS3D_ERR_BAD_ARG;
  }
#undef OS3D_ATTN_BWD_CASE
  return 0;
}

}  // namespace os3d

using namespace os3d;

template <typename T>
static int launch_qk_normalize(T *q, T *k, int64_t ld, int64_t m, int heads, int d, cudaStream_t st) {
  const unsigned g = (unsigned)cdiv(m * heads * 2, 256);
  // vector path: slices (and therefore rows) aligned to the slice's vector width
  const bool aligned = ((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((ld * (int64_t)sizeof(T)) % 16 == 0);
#define OS3D_QKN_CASE(D)                                                          \
  case D:                                                                          \
    qk_normalize_vec_kernel<T, D><<<g, 256, 0, st>>>(q, k, ld, m, heads);          \
    return 0;
  if (aligned) {
    switch (d) {
      OS3D_QKN_CASE(6)
      OS3D_QKN_CASE(12)
      OS3D_QKN_CASE(16)
      OS3D_QKN_CASE(24)
      OS3D_QKN_CASE(32)
      OS3D_QKN_CASE(48)
      OS3D_QKN_CASE(64)
      default: break;
    }
  }
#undef OS3D_QKN_CASE
  qk_normalize_kernel<T><<<g, 256, 0, st>>>(q, k, ld, m, heads, d);
  return 0;
}

extern "C" int os3d_qk_normalize(void *q, void *k, int64_t ld, int64_t m, int c, int heads, int elem_size, void *stream) {
  if (m == 0) return 0;
  if (heads <= 0 || c % heads) return OS3D_ERR_BAD_ARG;
  if (elem_size == 4)
    launch_qk_normalize<float>((float *)q, (float *)k, ld, m, heads, c / heads, (cudaStream_t)stream);
  else if (elem_size == 2)
    launch_qk_normalize<__nv_bfloat16>((__nv_bfloat16 *)q, (__nv_bfloat16 *)k, ld, m, heads, c / heads, (cudaStream_t)stream);
  else
    return OS3D_ERR_BAD_ARG;
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_window_attention(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m, int c,
                                     int heads, const int32_t *order, const int32_t *seg_start, const int32_t *seg_len,
                                     const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min,
                                     float drop_p, uint64_t seed, int elem_size, void *out, void *stream) {
  if (m == 0) return 0;
  if (heads <= 0 || c % heads || drop_p < 0.0f || drop_p >= 1.0f) return OS3D_ERR_BAD_ARG;
  int rc;
  if (elem_size == 4)
    rc = launch_attention<float>((const float *)q, (const float *)k, (const float *)v, ld, ldv, m, c, heads, order, seg_start,
                                 seg_len, level_info, lvl_tokens, tau, tau_min, drop_p, seed, (float *)out, (cudaStream_t)stream);
  else if (elem_size == 2)
    rc = launch_attention<__nv_bfloat16>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v, ld,
                                         ldv, m, c, heads, order, seg_start, seg_len, level_info, lvl_tokens, tau, tau_min,
                                         drop_p, seed, (__nv_bfloat16 *)out, (cudaStream_t)stream);
  else
    return OS3D_ERR_BAD_ARG;
  if (rc) return rc;
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_window_attention_bwd(const void *q, const void *k, const void *v, const void *o, const void *go,
                                         int64_t m, int c, int heads, const int32_t *order, const int32_t *seg_start,
                                         const int32_t *seg_len, const int32_t *level_info, const int *lvl_tokens,
                                         const float *tau, float tau_min, float drop_p, uint64_t seed, int elem_size,
                                         float *stats, void *gq, void *gk, void *gv, float *g_inv_tau, void *stream) {
  if (m == 0) return 0;
  if (heads <= 0 || c % heads || drop_p < 0.0f || drop_p >= 1.0f) return OS3D_ERR_BAD_ARG;
  int rc;
  if (elem_size == 4)
    rc = launch_attention_bwd<float>((const float *)q, (const float *)k, (const float *)v, (const float *)o,
                                     (const float *)go, m, c, heads, order, seg_start, seg_len, level_info, lvl_tokens, tau,
                                     tau_min, drop_p, seed, stats, (float *)gq, (float *)gk, (float *)gv, g_inv_tau,
                                     (cudaStream_t)stream);
  else if (elem_size == 2)
    rc = launch_attention_bwd<__nv_bfloat16>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v,
                                             (const __nv_bfloat16 *)o, (const __nv_bfloat16 *)go, m, c, heads, order,
                                             seg_start, seg_len, level_info, lvl_tokens, tau, tau_min, drop_p, seed, stats,
                                             (__nv_bfloat16 *)gq, (__nv_bfloat16 *)gk, (__nv_bfloat16 *)gv, g_inv_tau,
                                             (cudaStream_t)stream);
  else
    return OS3D_ERR_BAD_ARG;
  if (rc) return rc;
  OS3D_LAUNCH_CHECK();
  return 0;
}
