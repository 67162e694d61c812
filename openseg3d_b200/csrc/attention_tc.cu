// attention_tc.cu -- stage 4 on the tensor cores: variable-length cosine window attention with tcgen05 / TMEM (bf16).
//
// Work item = (128 consecutive tokens of the window-grouped order, one head).  Because `order` groups voxels by
// window, the keys such a query tile can see are one contiguous range of `order` (from the start of the first touched
// window to the end of the last), walked in blocks of 64 keys:
//
//   gather  : q / k / v head slices (head-padded layout: dp = d rounded up to 16, so a slice is whole UMMA K-steps)
//             -> shared memory in the K-major no-swizzle core-matrix layout.  q and k are L2-normalised on the way
//             (F.normalize, cosine_msa.py:152-153) -- the separate normalisation pass disappears; V is transposed.
//   MMA 1   : S[128 x 64] = Q K^T                  tcgen05.mma kind::f16, accumulator in TMEM columns [0, 64)
//   softmax : thread t owns query row t (tcgen05.ld 32x32b: lane = row, no shuffles): scale by log2(e)/max(tau,tau_min),
//             mask keys of other windows (block-diagonal structure: compare window starts), online max / sum,
//             P = exp2(s - m) as bf16 -> shared memory (A operand of MMA 2); if a row maximum moved, the O accumulator
//             is rescaled in place in TMEM (tcgen05.ld -> mul -> tcgen05.st).
//   MMA 2   : O[128 x dp] += P V                   accumulator in TMEM columns [64, 64 + dp)
//   epilogue: O / l -> bf16 -> out[row, head*dp ...] in the original voxel order.
//
// No padding to max_tokens, no [R*h, T, T] score tensor, no key-padding masks (the reference builds all three:
// swformer_utils.py:34-64, cosine_msa.py:154-176, point_transformer_layer.py:210-220).  128 TMEM columns and ~45 KB
// of shared memory per CTA -> 4 CTAs per SM overlap each other's gather / MMA / softmax phases.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace attn_tc {
using namespace ptx;

#ifndef OS3D_ATTN_CTAS32
#define OS3D_ATTN_CTAS32 4       /* 5 (96 registers, S / O as two TMEM allocations): spills, 0.64 against 0.63 ms at level 3 */
#endif
#ifndef OS3D_ATTN_CTAS
#define OS3D_ATTN_CTAS 7
#endif
constexpr int kTileQ = 128;
constexpr int kThreads = 128;
constexpr int kLbo = 128;                 // bytes between core matrices along K

struct Params {
  const __nv_bfloat16 *q, *k, *v;         // rows: q, k pitch ld (elements), v pitch ldv; head h at column h * dp
  int64_t ld, ldv, ldo;
  const int32_t *order;                   // [n_tokens] voxel row of each grouped position
  const int2 *pos_seg;                    // [n_tokens] (window start position, window length) of each grouped position
  const int32_t *level_info;              // [14] = number of grouped positions
  const float *tau;
  float tau_min;
  __nv_bfloat16 *out;
  int heads;
  float drop_p;                           // DROP = 1: attention dropout (training), keep mask = hash(seed, head, query row, key row)
  uint64_t seed;
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed pairs of floats for the two-wide FP32 instructions of sm_100 (FFMA2 / FADD2)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// smem offset of element chunk (row r, 16-byte K-chunk c) in the K-major no-swizzle layout with 8-row group stride sbo
__device__ __forceinline__ uint32_t core_off(int r, int c, int sbo) { return (r >> 3) * sbo + c * kLbo + (r & 7) * 16; }

// DP: padded head dim.  KB: keys per block.  TMEM: S in columns [0, KB), O in [KB, KB + DP) of a power-of-two allocation
// -- with KB = 32 and DP = 16 that is 64 columns, so up to 8 CTAs share an SM (and short windows waste half as many
// masked score columns); KB = 64 otherwise (128 columns, 4 CTAs).
//
// KB = 64: P never touches shared memory (kPT): each thread packs its row of probabilities to bf16 pairs and writes them with
// tcgen05.st over the first KB / 2 columns of its own S row (which it holds in registers by then); MMA 2 takes its A
// operand from tensor memory.  Warps whose rows see no key of a block write nothing at all: MMA 2 is issued with their
// 32 lanes in the disable-output-lane mask (O is zeroed once up front, so every MMA 2 accumulates).  Before this the P
// tile went through 16-byte shared-memory stores -- 60 % of the kernel's shared-memory wavefronts on an LSU pipe that ncu
// showed 71-73 % busy (profiles/r02h_attention_l{1,3}_full.txt) -- including the all-zero rows of skipped warps.
// Measured per layer: level 3 0.666 -> 0.625 ms, level 4 0.387 -> 0.370 ms.  With KB = 32 (dp = 16) the P row is only four
// 16-byte stores and the tensor-memory round trip was slower (level 1 0.670 -> 0.704 ms), so that variant keeps P in
// shared memory as the A operand of a descriptor MMA.
//
// PRENORM = 1: q and k arrive already L2-normalised per head (the in-projection kernel's epilogue, qkv_tc.cu): the gather
// is a plain copy -- the per-(tile, head) re-normalisation of every key was ~20 % of the kernel's instructions.
// (A variant with a fifth, MMA-issuing warp and mbarrier hand-offs was built and measured in round 1 and superseded by
// attention_v2.cu in round 2; both lose to this block-synchronous kernel at 4-7 CTAs per SM: DESIGN.md 3.2.)
// DROP = 1 (training forward): the probabilities that reach MMA 2 are multiplied by the keep mask of attention.cu's
// drop_keep() -- the same hash of (seed, head, query row, key row), so os3d_window_attention_bwd regenerates it -- while the
// softmax denominator stays undropped (cosine_msa.py:173-174: dropout acts on the normalised weights).
// HP = 2 (dp = 16, 32-key blocks): two adjacent heads per CTA.  One gather of 64-byte row slices feeds both heads, the key
// rows / window masks / barriers of a block are shared, and both heads' MMAs go out under one commit: the per-(tile, head)
// fixed work -- two thirds of the instructions of the one-head kernel at levels 1-2 -- is paid once per pair.
template <int DP, int KB, int PRENORM = 0, int DROP = 0, int HP = 1>
__global__ void __launch_bounds__(kThreads, HP == 2 ? 5 : (KB == 32 ? OS3D_ATTN_CTAS : (DP == 32 ? OS3D_ATTN_CTAS32 : (DP < 32 ? 4 : 3))))
    window_attention_tc_kernel(const Params p) {
  constexpr int kBlockKeys = KB;
  static_assert(HP == 1 || (HP == 2 && PRENORM && KB == 32), "two heads per CTA: pre-normalised q / k, 32-key blocks");
  constexpr int kTmemCols = HP * (KB + DP) <= 64 ? 64 : 128;
  // HP = 2: the scores (64 columns) and the outputs (32) are two tensor-memory allocations -- 96 columns per CTA, five
  // CTAs per SM -- where one power-of-two block would take 128 (four CTAs)
  constexpr bool kSplitAlloc = HP == 2 && HP * KB == 64 && HP * DP == 32;
  constexpr int kOCol = HP * KB;                          // S of head hh in columns [hh KB, hh KB + KB), O in [kOCol + hh DP, ...)
  constexpr int kChunks = DP / 8;                         // 16-byte chunks per head slice
  constexpr int kSboQ = kChunks * kLbo + 16;              // +16: stagger 8-row groups across banks
  constexpr bool kPT = KB == 64;                          // P in tensor memory
  constexpr int kSboP = (kBlockKeys / 8) * kLbo + 16;
  constexpr int kSboV = (kBlockKeys / 8) * kLbo + 16;     // V as the MN-major B operand: stride between 8-dim groups; 8-key groups are kLbo apart
  constexpr int kQBytes = (kTileQ / 8) * kSboQ;
  constexpr int kKBytes = (kBlockKeys / 8) * kSboQ;
  constexpr int kVBytes = (DP / 8) * kSboV;
  constexpr int kPBytes = kPT ? 16 : (kTileQ / 8) * kSboP;

  __shared__ __align__(128) uint8_t q_s[HP][kQBytes];
  // PRENORM: K / V arrive ready to use, so the gather is cp.async (LDGSTS) straight into the operand layout of the NEXT
  // block's buffer -- no register staging (16 registers and their spills gone), no st.shared -- hence two buffers.
  constexpr int kBufs = PRENORM ? 2 : 1;
  __shared__ __align__(128) uint8_t k_s[kBufs][HP][kKBytes];
  __shared__ __align__(128) uint8_t v_s[kBufs][HP][kVBytes];
  __shared__ __align__(128) uint8_t p_s[HP][kPBytes];         // !kPT only
  __shared__ __align__(8) uint64_t bars[2];               // MMA 1 done, MMA 2 done
  __shared__ uint32_t tmem_slot, tmem_slot_o;
  __shared__ uint32_t lanes_off[4];
  __shared__ int32_t krow_s[DROP ? KB : 1];               // DROP: voxel row of each key of the block                       // per warp: 0xffffffff when its rows sit out MMA 2 of this block

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // heads of one query tile are adjacent in launch order: they run at the same time and share the q / k / v rows (each
  // head reads a 32-96 byte slice of the same lines) while those are still in L2.  With heads on the slow grid axis
  // every line came from HBM once per head (ncu: 3.1 GB of DRAM reads per launch for 0.7 GB of q / k / v).
  const int h = (blockIdx.x % (p.heads / HP)) * HP;       // first head of this CTA
  const int n_tok = __ldg(p.level_info + 14);
  const int p0 = (blockIdx.x / (p.heads / HP)) * kTileQ;
  if (p0 >= n_tok) return;
  const int p_last = min(p0 + kTileQ, n_tok) - 1;

  const uint32_t bar1 = smem_u32(&bars[0]), bar2 = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if constexpr (kSplitAlloc) tmem_alloc2(smem_u32(&tmem_slot), HP * KB, smem_u32(&tmem_slot_o), HP * DP);
    else tmem_alloc(smem_u32(&tmem_slot), kTmemCols);
  }

  // ---- this thread's query row ----
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
  if (PRENORM) {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.q + (int64_t)qrow * p.ld + h * DP);     // HP heads: one contiguous slice
#pragma unroll
    for (int hh = 0; hh < HP; ++hh)
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        *reinterpret_cast<uint4 *>(q_s[hh] + core_off(tid, c, kSboQ)) = q_ok ? __ldg(src + hh * kChunks + c) : make_uint4(0, 0, 0, 0);
  } else {
    float f[DP];
    if (q_ok) {
      const uint4 *src = reinterpret_cast<const uint4 *>(p.q + (int64_t)qrow * p.ld + h * DP);
      float ss = 0.0f;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const uint4 u = __ldg(src + c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f[c * 8 + 2 * i] = __uint_as_float(w[i] << 16);
          f[c * 8 + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
          ss = fmaf(f[c * 8 + 2 * i], f[c * 8 + 2 * i], ss);
          ss = fmaf(f[c * 8 + 2 * i + 1], f[c * 8 + 2 * i + 1], ss);
        }
      }
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < DP; ++i) f[i] *= inv;
    } else {
#pragma unroll
      for (int i = 0; i < DP; ++i) f[i] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(f[c * 8 + 2 * i], f[c * 8 + 2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t *>(&hh);
      }
      *reinterpret_cast<uint4 *>(q_s[0] + core_off(tid, c, kSboQ)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }

  // key range of the tile: start of the first touched window .. end of the last touched window
  const int2 seg_first = __ldg(&p.pos_seg[p0]);
  const int2 seg_last = __ldg(&p.pos_seg[p_last]);
  const int ks = seg_first.x, ke = seg_last.x + seg_last.y;
  const int n_blocks = (ke - ks + kBlockKeys - 1) / kBlockKeys;

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_o = kSplitAlloc ? tmem_slot_o : tmem_base + kOCol;                   // O accumulators
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16), tmem_row_o = tmem_o + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc1 = make_idesc_bf16(kTileQ, kBlockKeys), idesc2 = make_idesc_bf16(kTileQ, DP) | (1u << 16);   // bit 16: B is MN-major
  const float scale = 1.4426950408889634f / fmaxf(__ldg(p.tau), p.tau_min);
  const bool fixed_max = scale <= 60.0f;
  if constexpr (kPT) {
    // O starts at zero: MMA 2 always accumulates (rows of warps that sit a block out are masked, not multiplied by 0)
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
    for (int c0 = 0; c0 < HP * DP; c0 += 16) tmem_st16(tmem_row_o + c0, z);
    tmem_st_wait();
  }

  float m_run[HP], l_run[HP];
#pragma unroll
  for (int hh = 0; hh < HP; ++hh) { m_run[hh] = -INFINITY; l_run[hh] = 0.0f; }
  uint32_t ph1 = 0, ph2 = 0;
  const float keep_scale = DROP ? 1.0f / (1.0f - p.drop_p) : 1.0f;
  const uint64_t drop_row = p.seed ^ ((uint64_t)(uint32_t)qrow << 32);

  // Two threads per key; both read the whole K slice (the norm needs it), each stores the chunks c with (c & 1) == half
  // and transposes the same chunks of V.  The raw slices of block b+1 are fetched into registers while block b is in
  // its MMA / softmax phases, so the two dependent global loads (order -> row) are off the critical path.
  // (one head, KB = 32: threads 64..127 have no key and k_ok_next stays false; two heads: thread pairs 2 key + hh)
  const int key = tid / (2 * HP), half = tid & 1, hk = (tid >> 1) % HP;       // hk: which of the CTA's heads this thread gathers
  constexpr int kVChunks = (kChunks + 1) / 2;
  uint4 k_raw[PRENORM ? 1 : kChunks], v_raw[PRENORM ? 1 : kVChunks];          // !PRENORM: register staging (the norm needs the slice)
  bool k_ok_next = false;
  // the voxel row of this thread's key is requested TWO blocks ahead, the K / V slices one block ahead: the dependent pair
  // (order -> row -> slice) was one L2 round trip too long for a one-block lookahead (ncu: long-scoreboard stalls on the
  // slice address, and block-wide barrier stalls behind the gathering warps)
  int32_t krow_ahead = -1, krow_cur = -1;
  auto fetch_row = [&](int blk) {
    const int kp = ks + blk * kBlockKeys + key;
    krow_ahead = (blk < n_blocks && kp < ke && key < kBlockKeys) ? __ldg(p.order + kp) : -1;
  };
  auto prefetch = [&](int blk) {
    const int32_t krow = krow_ahead;                       // fetched by fetch_row(blk) one block earlier
    krow_cur = krow;
    k_ok_next = krow >= 0;
    fetch_row(blk + 1);
    if constexpr (PRENORM) {
      if (key < kBlockKeys && blk < n_blocks) {
        const int buf = blk & 1;
        const int32_t row = max(krow, 0);
        const uint32_t nb = k_ok_next ? 16u : 0u;            // no key here: zero-fill (the address stays valid)
        const uint4 *ksrc = reinterpret_cast<const uint4 *>(p.k + (int64_t)row * p.ld + (h + hk) * DP);
        const uint4 *vsrc = reinterpret_cast<const uint4 *>(p.v + (int64_t)row * p.ldv + (h + hk) * DP);
        const uint32_t kd = smem_u32(k_s[buf][hk]), vd = smem_u32(v_s[buf][hk]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          if ((c & 1) != half) continue;
          cp_async_16(kd + core_off(key, c, kSboQ), ksrc + c, nb);
          // V is the B operand of O += P V (N = head dims, K = keys), MN-major no-swizzle core matrices:
          //   element (dim n, key) -> (n / 8) * sbo + (key / 8) * lbo + (key % 8) * 16 + (n % 8) * 2
          cp_async_16(vd + c * kSboV + (key >> 3) * kLbo + (key & 7) * 16, vsrc + c, nb);
        }
      }
      cp_async_commit();
    } else if (k_ok_next) {
      const uint4 *ksrc = reinterpret_cast<const uint4 *>(p.k + (int64_t)krow * p.ld + (h + hk) * DP);
      const uint4 *vsrc = reinterpret_cast<const uint4 *>(p.v + (int64_t)krow * p.ldv + (h + hk) * DP);
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        if (!PRENORM || (c & 1) == half) k_raw[c] = __ldg(ksrc + c);      // PRENORM: only the chunks this thread stores
#pragma unroll
      for (int c = 0; c < kVChunks; ++c)
        if (2 * c + half < kChunks) v_raw[c] = __ldg(vsrc + 2 * c + half);
    }
  };
  fetch_row(0);
  prefetch(0);

  for (int blk = 0; blk < n_blocks; ++blk) {
    if (blk > 0) {            // MMA 2 of the previous block still reads V (shared memory) and P (tensor memory, under S)
      mbar_wait(bar2, ph2);
      ph2 ^= 1;
      tc_fence_after();
    }
    const int buf = PRENORM ? (blk & 1) : 0;
    // ---- PRENORM: this block's cp.async copies have landed; else registers -> shared memory: K normalised (K-major) ----
    if constexpr (PRENORM) {
      cp_async_wait<0>();
    } else if (key < kBlockKeys) {
      float f[DP];
      float ss = 0.0f;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const uint32_t w[4] = {k_raw[c].x, k_raw[c].y, k_raw[c].z, k_raw[c].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = k_ok_next ? __uint_as_float(w[i] << 16) : 0.0f;
          const float b2 = k_ok_next ? __uint_as_float(w[i] & 0xffff0000u) : 0.0f;
          f[c * 8 + 2 * i] = a;
          f[c * 8 + 2 * i + 1] = b2;
          ss = fmaf(a, a, ss);
          ss = fmaf(b2, b2, ss);
        }
      }
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if ((c & 1) != half) continue;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(f[c * 8 + 2 * i] * inv, f[c * 8 + 2 * i + 1] * inv);
          w[i] = *reinterpret_cast<const uint32_t *>(&hh);
        }
        *reinterpret_cast<uint4 *>(k_s[0][hk] + core_off(key, c, kSboQ)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    if (!PRENORM && key < kBlockKeys) {
      // V is the B operand of O += P V with N = head dims, K = keys.  Its rows (keys) are stored as they come from
      // global memory -- 16-byte chunks of 8 dims -- in the MN-major no-swizzle core-matrix layout:
      //   element (dim n, key) -> (n / 8) * sbo + (key / 8) * lbo + (key % 8) * 16 + (n % 8) * 2
      // so the gather is plain vector stores (the K-major layout needed V transposed: 2-byte scattered stores, 65 % of
      // the LSU data-pipe wavefronts of the kernel).
#pragma unroll
      for (int cv = 0; cv < kVChunks; ++cv) {
        const int c = 2 * cv + half;
        if (c >= kChunks) continue;
        const uint4 u = k_ok_next ? v_raw[cv] : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(v_s[0][hk] + c * kSboV + (key >> 3) * kLbo + (key & 7) * 16) = u;
      }
    }
    if (DROP && key < kBlockKeys && half == 0 && hk == 0) krow_s[key] = krow_cur;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < HP; ++hh)
#pragma unroll
        for (int s = 0; s < DP / 16; ++s)
          umma_bf16(tmem_base + hh * kBlockKeys, make_kmajor_nosw_desc(smem_u32(q_s[hh]) + s * 2 * kLbo, kLbo, kSboQ),
                    make_kmajor_nosw_desc(smem_u32(k_s[buf][hh]) + s * 2 * kLbo, kLbo, kSboQ), idesc1, s > 0 ? 1u : 0u);
      umma_commit(bar1);
    }
    prefetch(blk + 1);        // global loads for the next key block fly during MMA 1, the softmax and MMA 2
    mbar_wait(bar1, ph1);
    ph1 ^= 1;
    tc_fence_after();

    // ---- softmax on this thread's row ----
    // A key at grouped position kp is in this row's window iff qws <= kp < qws + qlen, i.e. the valid keys of the block
    // are the index range [lo, hi): no mask array in memory.
    const int kb0 = ks + blk * kBlockKeys;
    const int lo = max(qws - kb0, 0), hi = min(qws + qlen - kb0, kBlockKeys);
    // valid keys as a bit mask: one LOP3 + FSEL per element instead of two compares; the scale and the running maximum
    // are folded into one FFMA feeding ex2.approx.ftz (masked scores are -inf -> probability exactly 0)
    const uint64_t valid = hi > lo ? ((~0ull >> (64 - (hi - lo))) << lo) : 0ull;
    const uint32_t v_lo = (uint32_t)valid, v_hi = (uint32_t)(valid >> 32);
    // Windows are short, so most (warp, key block) pairs are entirely off the block diagonal: one vote skips the TMEM
    // load, the whole softmax and the P store for them, and takes their rows out of MMA 2.
    const bool warp_has_keys = __any_sync(0xffffffffu, valid != 0ull);
    if (kPT && lane == 0) lanes_off[warp] = warp_has_keys ? 0u : 0xffffffffu;
    // q and k are unit vectors, so a raw score never exceeds 1 (+ bf16 rounding): with a moderate temperature the
    // softmax can use the FIXED reference maximum 1 -- no running maximum, no FMNMX per score, no rescaling of O --
    // and cannot underflow (2 * scale <= 120 binades).  Small temperatures keep the online maximum.
    const uint64_t full_mask = kBlockKeys == 64 ? ~0ull : ((1ull << (kBlockKeys & 63)) - 1ull);
    const bool all_valid = __all_sync(0xffffffffu, valid == full_mask);
#pragma unroll
    for (int hh = 0; hh < HP; ++hh) {                       // the heads of this CTA share the block's masks and key rows
    float alpha = 1.0f;
    uint32_t pk[kBlockKeys / 2];                            // the row of probabilities, packed bf16 pairs
    const uint32_t s_col = tmem_row + hh * kBlockKeys, o_col = tmem_row_o + hh * DP;
    if (warp_has_keys) {
    float s[kBlockKeys];
    {
      uint32_t r[32];
      tmem_ld32(s_col, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
      if constexpr (kBlockKeys == 64) {
        tmem_ld32(s_col + 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[32 + i] = __uint_as_float(r[i]);
      }
    }
    float m_new = m_run[hh], neg_ms;
    if (fixed_max) {
      if (!all_valid) {
#pragma unroll
        for (int j = 0; j < kBlockKeys; ++j) {
          const bool ok = ((j < 32 ? v_lo : v_hi) >> (j & 31)) & 1u;
          s[j] = ok ? s[j] : -INFINITY;
        }
      }
      neg_ms = -scale;
    } else {
#pragma unroll
      for (int j = 0; j < kBlockKeys; ++j) {                       // maxima are kept in raw (unscaled) score units
        const bool ok = ((j < 32 ? v_lo : v_hi) >> (j & 31)) & 1u;
        s[j] = ok ? s[j] : -INFINITY;
        m_new = fmaxf(m_new, s[j]);
      }
      const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;    // nothing valid so far: ex2(-inf) = 0 everywhere
      alpha = ex2_ftz((m_run[hh] - m_use) * scale);        // m_run = -inf -> 0 (l_run, O are still 0 then)
      neg_ms = -m_use * scale;
    }
    // two scores per instruction: FFMA2 / FADD2 (fma.rn.f32x2, add.rn.f32x2 -- the kernel is bound by its instruction
    // count, not by latency: 65 % of the issue slots were busy).  Each group of 32 keys goes back to tensor memory as 16
    // packed bf16 pairs (one tcgen05.st) as soon as it is done.
    const uint64_t sc2 = pack2(scale, scale), ng2 = pack2(neg_ms, neg_ms);
    uint64_t l2 = pack2(0.0f, 0.0f);
    const uint64_t drop_base = drop_row ^ ((uint64_t)(h + hh) * 0x9e3779b97f4a7c15ULL);
    auto keep_of = [&](int32_t krow) -> float {               // drop_keep() of attention.cu, bit for bit
      const uint64_t x = mix64(drop_base ^ (uint64_t)(uint32_t)krow);
      return (float)(x >> 40) * (1.0f / 16777216.0f) >= p.drop_p ? keep_scale : 0.0f;
    };
#pragma unroll
    for (int g = 0; g < kBlockKeys; g += 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float x0, x1;
        unpack2(ffma2(pack2(s[g + j], s[g + j + 1]), sc2, ng2), x0, x1);
        float a = ex2_ftz(x0), b = ex2_ftz(x1);
        l2 = fadd2(l2, pack2(a, b));                                  // the denominator ignores dropout
        if constexpr (DROP) {
          a *= keep_of(krow_s[g + j]);
          b *= keep_of(krow_s[g + j + 1]);
        }
        const __nv_bfloat162 ab = __floats2bfloat162_rn(a, b);
        pk[(g + j) >> 1] = *reinterpret_cast<const uint32_t *>(&ab);
      }
      if constexpr (kPT) tmem_st16(s_col + g / 2, reinterpret_cast<const uint32_t (&)[16]>(pk[g / 2]));
    }
    float l_lo, l_hi;
    unpack2(l2, l_lo, l_hi);
    const float l_blk = l_lo + l_hi;
    l_run[hh] = l_run[hh] * alpha + l_blk;
    m_run[hh] = m_new;
    } else if constexpr (!kPT) {
#pragma unroll
      for (int i = 0; i < kBlockKeys / 2; ++i) pk[i] = 0u;
    }
    if constexpr (!kPT) {
#pragma unroll
      for (int c = 0; c < kBlockKeys / 8; ++c)
        *reinterpret_cast<uint4 *>(p_s[hh] + core_off(tid, c, kSboP)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    }
    // rescale O if any row of this warp moved its maximum (tcgen05.ld / st are warp-collective)
    if (blk > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 16) {
        uint32_t o[16];
        tmem_ld16(o_col + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(o_col + c0, o);
      }
      if constexpr (!kPT) tmem_st_wait();
    }
    }   // heads of this CTA
    if constexpr (kPT) {
      if (warp_has_keys) tmem_st_wait();
    }
    if constexpr (!kPT) fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (!kPT && tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < HP; ++hh)
#pragma unroll
        for (int s2 = 0; s2 < kBlockKeys / 16; ++s2)
          umma_bf16(tmem_o + hh * DP, make_kmajor_nosw_desc(smem_u32(p_s[hh]) + s2 * 2 * kLbo, kLbo, kSboP),
                    make_kmajor_nosw_desc(smem_u32(v_s[buf][hh]) + s2 * 2 * kLbo, kLbo, kSboV), idesc2, (blk > 0 || s2 > 0) ? 1u : 0u);
      umma_commit(bar2);
    }
    if (kPT && tid == 0) {
      tc_fence_after();
      const uint32_t off0 = lanes_off[0], off1 = lanes_off[1], off2 = lanes_off[2], off3 = lanes_off[3];
      if ((off0 & off1 & off2 & off3) == 0u) {
        const uint32_t v_hi_w = nosw_desc_hi(kSboV);
#pragma unroll
        for (int s2 = 0; s2 < kBlockKeys / 16; ++s2)
          umma_bf16_ts_acc(tmem_o, tmem_base + s2 * 8, nosw_desc_lo(smem_u32(v_s[buf][0]) + s2 * 2 * kLbo, kLbo), v_hi_w,
                           idesc2, off0, off1, off2, off3);
      }
      umma_commit(bar2);
    }
  }

  // ---- epilogue ----
  mbar_wait(bar2, ph2);
  tc_fence_after();
#pragma unroll
  for (int hh = 0; hh < HP; ++hh) {
    const float inv_l = l_run[hh] > 0.0f ? 1.0f / l_run[hh] : 0.0f;
    __nv_bfloat16 *dst = p.out + (int64_t)qrow * p.ldo + (h + hh) * DP;
#pragma unroll
    for (int c0 = 0; c0 < DP; c0 += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_row_o + hh * DP + c0, o);
      tmem_ld_wait();
      if (q_ok) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 ab = __floats2bfloat162_rn(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
          w[i] = *reinterpret_cast<const uint32_t *>(&ab);
        }
        reinterpret_cast<uint4 *>(dst + c0)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4 *>(dst + c0)[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (kSplitAlloc) {
      tmem_dealloc(tmem_base, HP * KB);
      tmem_dealloc(tmem_o, HP * DP);
    } else {
      tmem_dealloc(tmem_base, kTmemCols);
    }
  }
}

}  // namespace attn_tc
}  // namespace os3d

using namespace os3d;

static int launch_attn_tc(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m, int heads, int dp,
                          const int32_t *order, const int32_t *pos_seg, const int32_t *level_info, const float *tau,
                          float tau_min, void *out, int64_t ldo, void *stream, int prenorm, float drop_p = 0.0f,
                          uint64_t seed = 0) {
  if (m == 0) return 0;
  if (heads <= 0 || (dp != 16 && dp != 32 && dp != 48) || ld % 8 || ldv % 8 || ldo % 8) return OS3D_ERR_BAD_ARG;
  attn_tc::Params p;
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
  p.heads = heads;
  p.drop_p = drop_p;
  p.seed = seed;
  dim3 grid((unsigned)(cdiv(m, attn_tc::kTileQ) * heads));
  cudaStream_t st = (cudaStream_t)stream;
  const char *e = getenv("OS3D_ATTN_KB");                       // tuning override: keys per block for dp = 16
  const bool kb32 = e ? atoi(e) != 64 : true;
  const char *e2 = getenv("OS3D_ATTN_HP");                      // tuning override: heads per CTA for dp = 16 (1 | 2)
  const bool pair = dp == 16 && kb32 && prenorm && heads % 2 == 0 && (e2 ? atoi(e2) == 2 : true);
  dim3 grid2((unsigned)(cdiv(m, attn_tc::kTileQ) * (heads / 2)));
  if (drop_p > 0.0f) {
    if (!prenorm || drop_p >= 1.0f) return OS3D_ERR_BAD_ARG;
    if (pair) attn_tc::window_attention_tc_kernel<16, 32, 1, 1, 2><<<grid2, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 16 && kb32) attn_tc::window_attention_tc_kernel<16, 32, 1, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 16) attn_tc::window_attention_tc_kernel<16, 64, 1, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 32) attn_tc::window_attention_tc_kernel<32, 64, 1, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else attn_tc::window_attention_tc_kernel<48, 64, 1, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
  } else if (prenorm) {
    // (32 keys per block was also measured for dp = 32: 0.95 ms against 0.69 ms per level-3 layer with 64)
    if (pair) attn_tc::window_attention_tc_kernel<16, 32, 1, 0, 2><<<grid2, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 16 && kb32) attn_tc::window_attention_tc_kernel<16, 32, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 16) attn_tc::window_attention_tc_kernel<16, 64, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 32) attn_tc::window_attention_tc_kernel<32, 64, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else attn_tc::window_attention_tc_kernel<48, 64, 1><<<grid, attn_tc::kThreads, 0, st>>>(p);
  } else {
    if (dp == 16 && kb32) attn_tc::window_attention_tc_kernel<16, 32><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 16) attn_tc::window_attention_tc_kernel<16, 64><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else if (dp == 32) attn_tc::window_attention_tc_kernel<32, 64><<<grid, attn_tc::kThreads, 0, st>>>(p);
    else attn_tc::window_attention_tc_kernel<48, 64><<<grid, attn_tc::kThreads, 0, st>>>(p);
  }
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_window_attention_bf16_tc(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv,
                                             int64_t m, int heads, int dp, const int32_t *order, const int32_t *pos_seg,
                                             const int32_t *level_info, const float *tau, float tau_min, void *out,
                                             int64_t ldo, void *stream) {
  return launch_attn_tc(q, k, v, ld, ldv, m, heads, dp, order, pos_seg, level_info, tau, tau_min, out, ldo, stream, 0);
}

extern "C" int os3d_window_attention_bf16_tc_prenorm(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv,
                                                     int64_t m, int heads, int dp, const int32_t *order,
                                                     const int32_t *pos_seg, const int32_t *level_info, const float *tau,
                                                     float tau_min, void *out, int64_t ldo, void *stream) {
  return launch_attn_tc(q, k, v, ld, ldv, m, heads, dp, order, pos_seg, level_info, tau, tau_min, out, ldo, stream, 1);
}

// training forward: q / k pre-normalised, attention dropout with the keep mask os3d_window_attention_bwd regenerates
extern "C" int os3d_window_attention_bf16_tc_drop(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv,
                                                  int64_t m, int heads, int dp, const int32_t *order,
                                                  const int32_t *pos_seg, const int32_t *level_info, const float *tau,
                                                  float tau_min, float drop_p, uint64_t seed, void *out, int64_t ldo,
                                                  void *stream) {
  if (drop_p < 0.0f || drop_p >= 1.0f) return OS3D_ERR_BAD_ARG;
  return launch_attn_tc(q, k, v, ld, ldv, m, heads, dp, order, pos_seg, level_info, tau, tau_min, out, ldo, stream, 1, drop_p,
                        seed);
}
