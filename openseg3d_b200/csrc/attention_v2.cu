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
//   warp 9    : loader.  cp.async (LDGSTS, 16 B) of WHOLE row slices (HG * DP * 2 = 192-256 B per row: all heads of the
//               group at once) of K and V for the next key block into a 2-stage ring, Q once; `order` is read once per
//               key, not once per head.  Completion through cp.async.mbarrier.arrive.noinc.
//   warp 8    : MMA issuer.  S = Q_h K_h^T into one of four TMEM score buffers, O_h += P V_h; MMA 1 of unit u+4 is issued
//               right after MMA 2 of unit u, so scores are ready two units before a warpgroup needs them.
//   warps 0-7 : softmax, two warpgroups taking alternate units (even / odd heads), thread = query row (tcgen05.ld 32x32b: lane = row).  Fixed-maximum softmax: q, k are unit
//               vectors, so every score is <= 1 and p = 2^(scale * (s - 1)) needs neither a running maximum nor a
//               rescale of O; the host selects this kernel only when scale = log2(e) / max(tau, tau_min) <= 60 (no
//               underflow of a whole row), else attention_tc.cu's online-maximum kernel runs.  Keys of other windows
//               are masked with the row's window range as a bit mask; a warp none of whose rows can see a key block
//               skips the TMEM load and the exponentials.  P goes to shared memory as bf16 in the K-major core-matrix
//               layout (lane = row: every STS.128 of a warp is one contiguous 512-byte run).
//   hand-offs : mbarriers only (s_full / s_empty, p_full / p_empty, kv_full / kv_empty); no __syncthreads in the loop.
//
// Shared memory: Q 32 KB + 2 x (K 16 KB + V 16 KB) + 4 x P 16 KB = 160 KB -> 1 CTA per SM with 8 softmax warps; TMEM
// 4 x 64 score columns + HG * DP output accumulators (<= 384 of 512).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace attn_v2 {
using namespace ptx;

constexpr int kTileQ = 128;
constexpr int kKB = 64;                   // keys per block
constexpr int kThreads = 192;             // 4 softmax warps + issuer + loader
constexpr int kIssuer = 4, kLoader = 5;   // warp indices
constexpr int kNB = 2;                    // score buffers (TMEM): unit u uses buffer u & 1

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
  int vcs;                                // byte stride between 8-dim chunks of V in shared memory (= kKB * 16), passed at RUN TIME:
                                          // with the V descriptor's SBO a compile-time constant nvcc 12.9 emitted MMA 2 with the P
                                          // descriptor's high word for both operands (dims >= 8 of every head read from the wrong chunk;
                                          // reproduced on B200, tools/debug_attn_v2.py) -- a value the compiler cannot fold avoids it
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
//
// Hand-off protocol (two mbarrier families per unit, ONE tcgen05.commit per unit):
//   issuer  : MMA 1 of units 0, 1;  then per unit u:  wait p_full(u) -> O_h += P V_h -> S[u & 1] = Q K^T of unit u + 2 ->
//             commit s_full[u & 1].  The tensor pipe executes one thread's MMAs in issue order, so that commit also says
//             "MMA 2 of unit u has read P" and "S[u & 1] holds unit u + 2".  (A commit is issued even when unit u + 2
//             does not exist: the last units need it as the P-free signal.)
//   softmax : wait s_full(u) -> tcgen05.ld -> exponentials -> wait s_full(u + 1) (= MMA 2 of unit u - 1 has read the
//             single P buffer; the scores of unit u + 1 arrive with the same signal) -> P -> fence -> arrive p_full(u).
// Measured alternatives (profiles/r02b_*, r02c_*): separate s_empty / p_empty barriers (two more commits per unit: the
// issuer became the pacer); one resident CTA with two softmax warpgroups and two issuers (prologue, Q load and epilogue
// of the only CTA on the SM are exposed: 25 % of all stall samples).  Two co-resident CTAs overlap each other's
// start-up, epilogue and hand-off latencies.
template <int HG, int DP>
__global__ void __launch_bounds__(kThreads, 2) window_attention_v2_kernel(const Params p) {
  constexpr int kW = HG * DP;                          // columns of the group's row slice
  constexpr int kChunks = kW / 8;                      // 16-byte chunks per row slice
  constexpr int kNS = 2;                               // K / V ring depth
  constexpr int kQBytes = kChunks * kTileQ * 16;
  constexpr int kKVBytes = kChunks * kKB * 16;
  constexpr int kPBytes = (kKB / 8) * kTileQ * 16;
  constexpr int kTmemCols = 256;
  constexpr int kOCol = kNB * kKB;                     // O accumulators start after the score buffers
  static_assert(kW <= 128 && DP % 16 == 0, "head group");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *q_s = smem;
  uint8_t *k_s = q_s + kQBytes;                        // [kNS][kKVBytes]
  uint8_t *v_s = k_s + kNS * kKVBytes;                 // [kNS][kKVBytes]
  uint8_t *p_s = v_s + kNS * kKVBytes;                 // [kPBytes]
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
  const uint32_t kv_full = smem_u32(&bars[2]), kv_empty = smem_u32(&bars[4]);      // [kNS] each, 8 bytes apart
  const uint32_t s_full = smem_u32(&bars[6]), p_full = smem_u32(&bars[8]);         // s_full[kNB = 2], p_full
  if (tid == 0) {
    mbar_init(q_full, 32);
    mbar_init(o_done, 1);
    for (int i = 0; i < kNS; ++i) {
      mbar_init(kv_full + 8 * i, 32);
      mbar_init(kv_empty + 8 * i, 1);
    }
    for (int i = 0; i < kNB; ++i) mbar_init(s_full + 8 * i, 1);
    mbar_init(p_full, 4);
    fence_barrier_init();
  }
  if (warp == kIssuer) tmem_alloc(smem_u32(tmem_slot), kTmemCols);

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

  if (warp == kLoader) {
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
          cp_async_16(vd + c * p.vcs + r * 16, vsrc + c * 8, ok ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive_noinc(kv_full + 8 * st);
    }
  } else if (warp == kIssuer) {
    // ------------------------------------------------------------------ MMA issuer ----
    const uint32_t idesc1 = make_idesc_bf16(kTileQ, kKB);
    const uint32_t idesc2 = make_idesc_bf16(kTileQ, DP) | (1u << 16);          // bit 16: B (V) is MN-major
    // descriptor words (tc_ptx.cuh: umma_bf16_words).  Start addresses advance in units of 16 bytes in the low word.
    const uint32_t hi_k = nosw_desc_hi(128);                                   // K-major operands (Q, K, P): SBO = 128
    const uint32_t hi_v = nosw_desc_hi((uint32_t)p.vcs);                       // V (MN-major): SBO = chunk stride
    const uint32_t q_lo = nosw_desc_lo(smem_u32(q_s), kTileQ * 16);
    const uint32_t p_lo = nosw_desc_lo(smem_u32(p_s), kTileQ * 16);
    constexpr uint32_t kStepQ = (2 * kTileQ * 16) >> 4;                        // one K = 16 step (2 chunks) of Q / P
    constexpr uint32_t kStepK = (2 * kKB * 16) >> 4;                           // ... of K
    constexpr uint32_t kStepV = (2 * 128) >> 4;                                // 16 keys of V (MN-major: 8-key groups 128 B apart)
    constexpr uint32_t kHeadQ = ((DP / 8) * kTileQ * 16) >> 4;                 // next head's slice of Q
    constexpr uint32_t kHeadK = ((DP / 8) * kKB * 16) >> 4;                    // ... of K
    mbar_wait(q_full, 0);
    fence_proxy_async();                 // LDGSTS (generic proxy) writes observed through the barrier -> visible to the MMA's async proxy
    tc_fence_after();
    int kv_seen = -1;                    // last block whose kv_full barrier has been waited for
    auto mma1 = [&](int u) {             // S[u & 1] = Q_h K_h^T (if unit u exists), then commit -> s_full[u & 1]
      const int sb = u & (kNB - 1);
      if (u < n_units) {
        const int b = u / HG, h = u - b * HG, st = b & 1;
        if (b > kv_seen) {
          mbar_wait(kv_full + 8 * st, (uint32_t)(b >> 1) & 1u);
          fence_proxy_async();
          tc_fence_after();
          kv_seen = b;
        }
        const uint32_t k_lo = nosw_desc_lo(smem_u32(k_s) + st * kKVBytes, kKB * 16) + h * kHeadK;
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < DP / 16; ++s)
            umma_bf16_words(tmem_base + sb * kKB, q_lo + h * kHeadQ + s * kStepQ, hi_k, k_lo + s * kStepK, hi_k, idesc1,
                            s > 0 ? 1u : 0u);
          umma_commit(s_full + 8 * sb);
        }
      } else if (elect_one()) {
        umma_commit(s_full + 8 * sb);    // no such unit: the commit only reports "MMA 2 of unit u - 2 is done"
      }
      __syncwarp();
    };
    mma1(0);
    mma1(1);
    for (int u = 0; u < n_units; ++u) {
      const int b = u / HG, h = u - b * HG, st = b & 1;
      mbar_wait(p_full, (uint32_t)u & 1u);
      tc_fence_after();
      const uint32_t v_lo = nosw_desc_lo(smem_u32(v_s) + st * kKVBytes + h * (DP / 8) * p.vcs, 128);
      if (elect_one()) {
#pragma unroll
        for (int s2 = 0; s2 < kKB / 16; ++s2)
          umma_bf16_words(tmem_base + kOCol + h * DP, p_lo + s2 * kStepQ, hi_k, v_lo + s2 * kStepV, hi_v, idesc2,
                          (b > 0 || s2 > 0) ? 1u : 0u);
        if (h == HG - 1) umma_commit(kv_empty + 8 * st);
        if (u == n_units - 1) umma_commit(o_done);
      }
      __syncwarp();
      if (u + 1 < n_units) mma1(u + 2);          // (the last unit needs no successor signal)
    }
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

    int u = 0;
    for (int b = 0; b < n_blocks; ++b) {
      // valid keys of this block for this row: the index range [lo, hi) as a bit mask; per 16-key group: does ANY row of
      // the warp see a key of the group (else its 16 scores are neither loaded nor exponentiated: windows are short, most
      // of a 128 x 64 score tile lies off the block diagonal), do ALL rows see all of them (no masking needed)
      uint64_t valid;
      uint32_t grp_any = 0, grp_all = 0;
      {
        const int kb0 = ks + b * kKB;
        const int lo = max(qws - kb0, 0), hi = min(qws + qlen - kb0, kKB);
        valid = hi > lo ? ((~0ull >> (64 - (hi - lo))) << lo) : 0ull;
#pragma unroll
        for (int gq = 0; gq < kKB / 16; ++gq) {
          const uint32_t bits = (uint32_t)(valid >> (16 * gq)) & 0xffffu;
          grp_any |= (__any_sync(0xffffffffu, bits != 0u) ? 1u : 0u) << gq;
          grp_all |= (__all_sync(0xffffffffu, bits == 0xffffu) ? 1u : 0u) << gq;
        }
      }
#pragma unroll
      for (int h = 0; h < HG; ++h, ++u) {
        const int sb = u & (kNB - 1);
        uint32_t pk[kKB / 2];
        // EVERY warp waits for the unit's scores, also one that will not read them: a warp that skipped the wait could reach
        // the buffer's NEXT use before this phase completes, and mbarrier parity waits cannot tell "two phases ahead" from
        // "done" (seen on hardware: rows that skipped block 0 read block 1's scores before MMA 1 had been issued)
        mbar_wait(s_full + 8 * sb, (uint32_t)(u >> 1) & 1u);
        float l_blk = 0.0f;
        if (grp_any) {
          tc_fence_after();
          uint32_t r[kKB / 16][16];
#pragma unroll
          for (int gq = 0; gq < kKB / 16; ++gq)
            if ((grp_any >> gq) & 1u) tmem_ld16(tmem_row + sb * kKB + 16 * gq, r[gq]);
          tmem_ld_wait();
#pragma unroll
          for (int gq = 0; gq < kKB / 16; ++gq) {
            if ((grp_any >> gq) & 1u) {
              const uint32_t bits = (uint32_t)(valid >> (16 * gq)) & 0xffffu;
              const bool full = (grp_all >> gq) & 1u;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float s0 = __uint_as_float(r[gq][j]), s1 = __uint_as_float(r[gq][j + 1]);
                if (!full) {
                  s0 = ((bits >> j) & 1u) ? s0 : -INFINITY;
                  s1 = ((bits >> (j + 1)) & 1u) ? s1 : -INFINITY;
                }
                const float a = ex2_ftz(fmaf(s0, scale, neg_scale)), b2 = ex2_ftz(fmaf(s1, scale, neg_scale));
                l_blk += a + b2;
                const __nv_bfloat162 hv = __floats2bfloat162_rn(a, b2);
                pk[8 * gq + (j >> 1)] = *reinterpret_cast<const uint32_t *>(&hv);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[8 * gq + i] = 0u;
            }
          }
          l_run[h] += l_blk;
        } else {
          // no row of this warp sees this key block: P rows are zero, the scores stay unread
#pragma unroll
          for (int i = 0; i < kKB / 2; ++i) pk[i] = 0u;
        }
        // the single P buffer is free once MMA 2 of unit u - 1 has read it: that is what s_full(u + 1) reports (it also
        // carries the next unit's scores, so the wait at the top of the next iteration returns at once)
        if (u >= 1) mbar_wait(s_full + 8 * (sb ^ 1), (uint32_t)((u + 1) >> 1) & 1u);
#pragma unroll
        for (int c = 0; c < kKB / 8; ++c)
          *reinterpret_cast<uint4 *>(p_s + c * (kTileQ * 16) + tid * 16) =
              make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async();
        tc_fence_before();                                       // orders this warp's tcgen05.ld of S(u) before MMA 1 of unit u + 2
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
            const __nv_bfloat162 hv = __floats2bfloat162_rn(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
            w[i] = *reinterpret_cast<const uint32_t *>(&hv);
          }
          reinterpret_cast<uint4 *>(dst + h * DP + c0)[0] = make_uint4(w[0], w[1], w[2], w[3]);
          reinterpret_cast<uint4 *>(dst + h * DP + c0)[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIssuer) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int HG, int DP>
int launch(const Params &p, int64_t m, int heads, cudaStream_t st) {
  constexpr int kW = HG * DP;
  constexpr int kChunks = kW / 8;
  constexpr size_t smem = (size_t)kChunks * kTileQ * 16 + 4 * (size_t)kChunks * kKB * 16 + (size_t)(kKB / 8) * kTileQ * 16 + 16 * 8 + 16;
  static int configured_dev[64] = {0};
  int dev = 0;
  OS3D_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured_dev[dev]) {
    OS3D_CUDA(cudaFuncSetAttribute(window_attention_v2_kernel<HG, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured_dev[dev] = 1;
  }
  Params q = p;
  q.groups = heads / HG;
  q.vcs = kKB * 16;
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
