// attention_v2.cu -- stage 4 on the tensor cores, second design: variable-length cosine window attention with ALL heads
// of a head group in one CTA, warp-specialised (bf16 operands, tcgen05 / TMEM).
//
// replaces flat2window -> CosineMultiheadAttention core (_scaled_cosine_attention, seg3d/models/layers/cosine_msa.py:115-177)
// -> window2flat (seg3d/utils/swformer_utils.py:34-85) as driven by WindowAttention.forward
// (seg3d/models/layers/point_transformer_layer.py:233-258); q and k arrive ALREADY L2-normalised per head (the projection
// kernel's epilogue does F.normalize, cosine_msa.py:152-153).
//
// Why a second design (ncu of attention_tc.cu, profiles/r01c_attention_l2_full.txt): one CTA per (128-query tile, ONE head)
// made eight CTAs re-gather the same rows in 32-byte slices, re-normalise every key once per query tile that sees it,
// and stall 24 % of the time on two block-wide barriers in front of a single-thread MMA issue.  Here:
//
//   work item : (128 consecutive positions of the window-grouped order, one GROUP of HG heads with HG * DP = 128 or 96
//               columns).  Keys = the contiguous range of `order` from the start of the first touched window to the end
//               of the last, walked in blocks of 64 keys.  A "unit" = (key block, head).
//   warp 5    : loader.  cp.async (LDGSTS, 16 B) of WHOLE row slices (HG * DP * 2 = 192-256 B per row: all heads of the
//               group at once) of K and V for the next key block into a 2-stage ring, Q once; `order` is read once per
//               key, not once per head.  Completion through cp.async.mbarrier.arrive.noinc.
//   warp 4    : MMA issuer.  S = Q_h K_h^T into one of two TMEM score buffers, O_h += P V_h; MMA 1 of unit u+1 is issued
//               before MMA 2 of unit u so the tensor pipe computes the next scores while the softmax warps work.
//   warps 0-3 : softmax, thread = query row (tcgen05.ld 32x32b: lane = row).  Fixed-maximum softmax: q, k are unit
//               vectors, so every score is <= 1 and p = 2^(scale * (s - 1)) needs neither a running maximum nor a
//               rescale of O; the host selects this kernel only when scale = log2(e) / max(tau, tau_min) <= 60 (no
//               underflow of a whole row), else attention_tc.cu's online-maximum kernel runs.  Keys of other windows
//               are masked with the row's window range as a bit mask; a warp none of whose rows can see a key block
//               skips the TMEM load and the exponentials.  P goes to shared memory as bf16 in the K-major core-matrix
//               layout (lane = row: every STS.128 of a warp is one contiguous 512-byte run).
//   hand-offs : mbarriers only (s_full / s_empty, p_full / p_empty, kv_full / kv_empty); no __syncthreads in the loop.
//
// Shared memory: Q 32 KB + 2 x (K 16 KB + V 16 KB) + P 16 KB = 112 KB -> 2 CTAs per SM; TMEM 256 columns per CTA
// (2 x 64 scores + HG * DP output accumulators).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace attn_v2 {
using namespace ptx;

constexpr int kTileQ = 128;
constexpr int kKB = 64;                   // keys per block
constexpr int kThreads = 192;             // 4 softmax warps + issuer + loader

struct Params {
  const __nv_bfloat16 *q, *k, *v;         // rows: q, k pitch ld (elements), v pitch ldv; head h at column h * dp; q, k normalised
  int64_t ld, ldv, ldo;
  const int32_t *order;                   // [n_tokens] voxel row of each grouped position
  const int2 *pos_seg;                    // [n_tokens] (window start position, window length) of each grouped position
  const int32_t *level_info;              // [14] = number of grouped positions
  const float *tau;
  float tau_min;
  __nv_bfloat16 *out;
  int groups;                             // head groups per tile (heads / HG)
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Shared-memory operand layout (no swizzle, core matrices of 8 rows x 16 bytes): [16-byte chunk c][row r][16 B], i.e. a
// chunk of 8 consecutive columns of all rows is contiguous.  As a K-major operand (Q, K, P): LBO (between chunks along K)
// = rows * 16, SBO (between 8-row groups) = 128.  As the MN-major B operand (V: N = head dims, K = keys): LBO (between
// 8-key groups) = 128, SBO (between 8-dim groups) = rows * 16.
template <int HG, int DP>
__global__ void __launch_bounds__(kThreads, 2) window_attention_v2_kernel(const Params p) {
  constexpr int kW = HG * DP;                          // columns of the group's row slice
  constexpr int kChunks = kW / 8;                      // 16-byte chunks per row slice
  constexpr int kQBytes = kChunks * kTileQ * 16;
  constexpr int kKVBytes = kChunks * kKB * 16;
  constexpr int kPBytes = (kKB / 8) * kTileQ * 16;
  constexpr int kTmemCols = 256;
  constexpr int kOCol = 2 * kKB;                       // O accumulators start after the two score buffers
  static_assert(kW <= 128 && DP % 16 == 0, "head group too wide");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *q_s = smem;
  uint8_t *k_s = q_s + kQBytes;                        // [2][kKVBytes]
  uint8_t *v_s = k_s + 2 * kKVBytes;                   // [2][kKVBytes]
  uint8_t *p_s = v_s + 2 * kKVBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(p_s + kPBytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // groups of one query tile are adjacent in launch order: they run at the same time and share `order` / pos_seg lines
  const int g = blockIdx.x % p.groups;
  const int n_tok = __ldg(p.level_info + 14);
  const int p0 = (blockIdx.x / p.groups) * kTileQ;
  if (p0 >= n_tok) return;
  const int p_last = min(p0 + kTileQ, n_tok) - 1;

  const uint32_t q_full = smem_u32(&bars[0]), o_done = smem_u32(&bars[1]);
  const uint32_t kv_full = smem_u32(&bars[2]), kv_empty = smem_u32(&bars[4]);      // [2] each, 8 bytes apart
  const uint32_t s_full = smem_u32(&bars[6]), s_empty = smem_u32(&bars[8]);        // [2] each
  const uint32_t p_full = smem_u32(&bars[10]), p_empty = smem_u32(&bars[11]);
  if (tid == 0) {
    mbar_init(q_full, 32);
    mbar_init(o_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kv_full + 8 * i, 32);
      mbar_init(kv_empty + 8 * i, 1);
      mbar_init(s_full + 8 * i, 1);
      mbar_init(s_empty + 8 * i, 4);
    }
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), kTmemCols);

  // key range of the tile: start of the first touched window .. end of the last touched window
  const int2 seg_first = __ldg(&p.pos_seg[p0]);
  const int2 seg_last = __ldg(&p.pos_seg[p_last]);
  const int ks = seg_first.x, ke = seg_last.x + seg_last.y;
  const int n_blocks = (ke - ks + kKB - 1) / kKB;
  const int n_units = n_blocks * HG;

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    // ------------------------------------------------------------------ loader ----
    const int64_t col0 = (int64_t)g * kW;
    {  // Q tile: rows p0 .. p0+127 (zero-filled past the end); lane handles rows lane, lane+32, ...
#pragma unroll
      for (int i = 0; i < kTileQ / 32; ++i) {
        const int r = lane + 32 * i;
        const int qp = p0 + r;
        const bool ok = qp <= p_last;
        const int32_t row = ok ? __ldg(p.order + qp) : 0;
        const __nv_bfloat16 *src = p.q + (int64_t)row * p.ld + col0;
        const uint32_t dst = smem_u32(q_s) + r * 16;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) cp_async_16(dst + c * (kTileQ * 16), src + c * 8, ok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(q_full);
    }
    for (int b = 0; b < n_blocks; ++b) {
      const int st = b & 1;
      if (b >= 2) mbar_wait(kv_empty + 8 * st, (uint32_t)((b >> 1) - 1) & 1u);
      const uint32_t kd = smem_u32(k_s) + st * kKVBytes, vd = smem_u32(v_s) + st * kKVBytes;
#pragma unroll
      for (int i = 0; i < kKB / 32; ++i) {
        const int r = lane + 32 * i;
        const int kp = ks + b * kKB + r;
        const bool ok = kp < ke;
        const int32_t row = ok ? __ldg(p.order + kp) : 0;
        const __nv_bfloat16 *ksrc = p.k + (int64_t)row * p.ld + col0;
        const __nv_bfloat16 *vsrc = p.v + (int64_t)row * p.ldv + col0;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          cp_async_16(kd + c * (kKB * 16) + r * 16, ksrc + c * 8, ok ? 16u : 0u);
          cp_async_16(vd + c * (kKB * 16) + r * 16, vsrc + c * 8, ok ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive_noinc(kv_full + 8 * st);
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer ----
    const uint32_t idesc1 = make_idesc_bf16(kTileQ, kKB);
    const uint32_t idesc2 = make_idesc_bf16(kTileQ, DP) | (1u << 16);          // bit 16: B (V) is MN-major
    const uint64_t qd = make_kmajor_nosw_desc(smem_u32(q_s), kTileQ * 16, 128);
    const uint64_t pd = make_kmajor_nosw_desc(smem_u32(p_s), kTileQ * 16, 128);
    constexpr uint64_t kStepQ = (uint64_t)(2 * kTileQ * 16) >> 4;              // one K = 16 step (2 chunks) of Q / P
    constexpr uint64_t kStepK = (uint64_t)(2 * kKB * 16) >> 4;                 // ... of K
    constexpr uint64_t kStepV = (uint64_t)(2 * 128) >> 4;                      // 16 keys of V (MN-major: 8-key groups 128 B apart)
    constexpr uint64_t kHeadQ = (uint64_t)((DP / 8) * kTileQ * 16) >> 4;       // next head's slice of Q
    constexpr uint64_t kHeadKV = (uint64_t)((DP / 8) * kKB * 16) >> 4;         // ... of K / V
    mbar_wait(q_full, 0);
    fence_proxy_async();                 // LDGSTS (generic proxy) writes observed through the barrier -> visible to the MMA's async proxy
    tc_fence_after();
    auto mma2 = [&](int u) {             // O_h += P V_h for unit u
      const int b = u / HG, h = u - b * HG, st = b & 1;
      mbar_wait(p_full, (uint32_t)u & 1u);
      tc_fence_after();
      const uint64_t vd = make_kmajor_nosw_desc(smem_u32(v_s) + st * kKVBytes, 128, kKB * 16) + h * kHeadKV;
      if (elect_one()) {
#pragma unroll
        for (int s2 = 0; s2 < kKB / 16; ++s2)
          umma_bf16(tmem_base + kOCol + h * DP, pd + s2 * kStepQ, vd + s2 * kStepV, idesc2, (b > 0 || s2 > 0) ? 1u : 0u);
        umma_commit(p_empty);
        if (h == HG - 1) umma_commit(kv_empty + 8 * st);
        if (u == n_units - 1) umma_commit(o_done);
      }
      __syncwarp();
    };
    for (int u = 0; u < n_units; ++u) {
      const int b = u / HG, h = u - b * HG, st = b & 1, sb = u & 1;
      if (h == 0) {
        mbar_wait(kv_full + 8 * st, (uint32_t)(b >> 1) & 1u);
        fence_proxy_async();
      }
      if (u >= 2) mbar_wait(s_empty + 8 * sb, (uint32_t)((u >> 1) - 1) & 1u);
      tc_fence_after();
      const uint64_t kd = make_kmajor_nosw_desc(smem_u32(k_s) + st * kKVBytes, kKB * 16, 128) + h * kHeadKV;
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < DP / 16; ++s)
          umma_bf16(tmem_base + sb * kKB, qd + h * kHeadQ + s * kStepQ, kd + s * kStepK, idesc1, s > 0 ? 1u : 0u);
        umma_commit(s_full + 8 * sb);
      }
      __syncwarp();
      if (u > 0) mma2(u - 1);
    }
    mma2(n_units - 1);
  } else {
    // ------------------------------------------------------------------ softmax warps ----
    const int qp = p0 + tid;
    const bool q_ok = qp <= p_last;
    int qws = 0, qlen = 0;                 // this row's window = grouped positions [qws, qws + qlen)
    int32_t qrow = 0;
    if (q_ok) {
      const int2 seg = __ldg(&p.pos_seg[qp]);
      qws = seg.x;
      qlen = seg.y;
      qrow = __ldg(p.order + qp);
    }
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float scale = 1.4426950408889634f / fmaxf(__ldg(p.tau), p.tau_min);
    const float neg_scale = -scale;
    float l_run[HG];
#pragma unroll
    for (int h = 0; h < HG; ++h) l_run[h] = 0.0f;

    uint32_t v_lo = 0, v_hi = 0;
    bool warp_has_keys = false, all_valid = false;
    int u = 0;
    for (int b = 0; b < n_blocks; ++b) {
      {  // valid keys of this block for this row: the index range [lo, hi) as a bit mask
        const int kb0 = ks + b * kKB;
        const int lo = max(qws - kb0, 0), hi = min(qws + qlen - kb0, kKB);
        const uint64_t valid = hi > lo ? ((~0ull >> (64 - (hi - lo))) << lo) : 0ull;
        v_lo = (uint32_t)valid;
        v_hi = (uint32_t)(valid >> 32);
        warp_has_keys = __any_sync(0xffffffffu, valid != 0ull);
        all_valid = __all_sync(0xffffffffu, valid == ~0ull);
      }
#pragma unroll
      for (int h = 0; h < HG; ++h, ++u) {
        const int sb = u & 1;
        uint32_t pk[kKB / 2];
        if (warp_has_keys) {
          mbar_wait(s_full + 8 * sb, (uint32_t)(u >> 1) & 1u);
          tc_fence_after();
          float s[kKB];
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(tmem_row + sb * kKB, r0);
            tmem_ld32(tmem_row + sb * kKB + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = __uint_as_float(r0[i]);
              s[32 + i] = __uint_as_float(r1[i]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty + 8 * sb);          // the score buffer may be overwritten by MMA 1 of unit u + 2
          if (!all_valid) {
#pragma unroll
            for (int j = 0; j < kKB; ++j) {
              const bool ok = ((j < 32 ? v_lo : v_hi) >> (j & 31)) & 1u;
              s[j] = ok ? s[j] : -INFINITY;
            }
          }
          float l_blk = 0.0f;
#pragma unroll
          for (int j = 0; j < kKB; j += 2) {
            const float a = ex2_ftz(fmaf(s[j], scale, neg_scale)), b2 = ex2_ftz(fmaf(s[j + 1], scale, neg_scale));
            l_blk += a + b2;
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b2);
            pk[j >> 1] = *reinterpret_cast<const uint32_t *>(&hh);
          }
          l_run[h] += l_blk;
        } else {
          // no row of this warp sees this key block: P rows are zero, the score buffer is released unread
          if (lane == 0) mbar_arrive(s_empty + 8 * sb);
#pragma unroll
          for (int i = 0; i < kKB / 2; ++i) pk[i] = 0u;
        }
        if (u > 0) mbar_wait(p_empty, (uint32_t)(u - 1) & 1u);   // MMA 2 of the previous unit has read P
#pragma unroll
        for (int c = 0; c < kKB / 8; ++c)
          *reinterpret_cast<uint4 *>(p_s + c * (kTileQ * 16) + tid * 16) =
              make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
    }

    // ---- epilogue: O_h / l_h -> bf16 -> out[row, (g * HG + h) * DP ...] in the original voxel order ----
    mbar_wait(o_done, 0);
    tc_fence_after();
    __nv_bfloat16 *dst = p.out + (int64_t)qrow * p.ldo + (int64_t)g * kW;
#pragma unroll
    for (int h = 0; h < HG; ++h) {
      const float inv_l = l_run[h] > 0.0f ? 1.0f / l_run[h] : 0.0f;
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 16) {
        uint32_t o[16];
        tmem_ld16(tmem_row + kOCol + h * DP + c0, o);
        tmem_ld_wait();
        if (q_ok) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
            w[i] = *reinterpret_cast<const uint32_t *>(&hh);
          }
          reinterpret_cast<uint4 *>(dst + h * DP + c0)[0] = make_uint4(w[0], w[1], w[2], w[3]);
          reinterpret_cast<uint4 *>(dst + h * DP + c0)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int HG, int DP>
int launch(const Params &p, int64_t m, int heads, cudaStream_t st) {
  constexpr int kW = HG * DP;
  constexpr int kChunks = kW / 8;
  constexpr size_t smem = (size_t)kChunks * kTileQ * 16 + 4 * (size_t)kChunks * kKB * 16 + (kKB / 8) * kTileQ * 16 + 16 * 8 + 16;
  static int configured_dev[64] = {0};
  int dev = 0;
  OS3D_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured_dev[dev]) {
    OS3D_CUDA(cudaFuncSetAttribute(window_attention_v2_kernel<HG, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured_dev[dev] = 1;
  }
  Params q = p;
  q.groups = heads / HG;
  dim3 grid((unsigned)(cdiv(m, kTileQ) * q.groups));
  window_attention_v2_kernel<HG, DP><<<grid, kThreads, smem, st>>>(q);
  OS3D_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn_v2
}  // namespace os3d

using namespace os3d;

extern "C" int os3d_window_attention_bf16_v2(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m,
                                             int heads, int dp, const int32_t *order, const int32_t *pos_seg,
                                             const int32_t *level_info, const float *tau, float tau_min, void *out,
                                             int64_t ldo, void *stream) {
  if (m == 0) return 0;
  if (heads <= 0 || ld % 8 || ldv % 8 || ldo % 8) return OS3D_ERR_BAD_ARG;
  attn_v2::Params p;
  p.q = (const __nv_bfloat16 *)q;
  p.k = (const __nv_bfloat16 *)k;
  p.v = (const __nv_bfloat16 *)v;
  p.ld = ld; p.ldv = ldv; p.ldo = ldo;
  p.order = order;
  p.pos_seg = (const int2 *)pos_seg;
  p.level_info = level_info;
  p.tau = tau;
  p.tau_min = tau_min;
  p.out = (__nv_bfloat16 *)out;
  p.groups = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (dp == 16 && heads % 8 == 0) return attn_v2::launch<8, 16>(p, m, heads, st);
  if (dp == 32 && heads % 4 == 0) return attn_v2::launch<4, 32>(p, m, heads, st);
  if (dp == 48 && heads % 2 == 0) return attn_v2::launch<2, 48>(p, m, heads, st);
  return OS3D_ERR_BAD_ARG;
}
