// mlp2_tc.cu -- the SWFormer MLP as ONE persistent tcgen05 kernel:
//
//   out = x + LayerNorm( fc2( GELU( fc1(x) ) ) )          (EncoderLayer, point_transformer_layer.py:260-298)
//
// The [M, 2C] hidden tensor never exists in HBM: per 128-row tile it is produced and consumed in chunks of HC hidden
// columns --  acc1 = x . W1[chunk]^T  ->  GELU -> bf16 -> shared memory  ->  acc2 += h_chunk . W2[:, chunk]^T  -- and the
// weights are STREAMED chunk by chunk from L2 (cp.async.bulk, separate 2-4 stage rings for W1 and W2), so the kernel also takes C = 192 whose two
// weight matrices (2 x 147 KB) do not fit in shared memory (the resident-weight chain kernel mlp_tc.cu stops at C = 96).
//
// What paces such a chain is not the tensor pipe but the hand-offs (MMA commit -> epilogue warps -> st.shared -> MMA: about
// a microsecond each when serialised, measured on mlp_tc.cu).  Here nothing on the epilogue warps' path waits for one:
//   * MMA 1 runs two chunks ahead of MMA 2 (two acc1 buffers, each handed back right after the epilogue warps' tcgen05.ld
//     of it), so its result is ready when the epilogue warps finish the chunk before;
//   * h has two buffers; MMA 2 (c) only has to finish before the epilogue of chunk c+2 writes its buffer again;
//   * acc2 has two buffers and the LayerNorm epilogue of tile t runs after the second hidden chunk of tile t+1.
//
//   * W1 and W2 chunks travel through separate rings: a W1 stage is free as soon as MMA 1 has read it (a chunk before
//     MMA 2 of the same chunk), so the next W1 copy -- the one on the critical path -- starts a whole chunk earlier than
//     with one combined stage (first version: 17.5 us per C = 192 tile, the L2 -> smem copy latency twice per chunk).
//
//   warps 0-15 : epilogue (4 per TMEM lane quarter)   warp 16 : producer (x tiles by TMA, W1 chunks)
//   warp 17 : MMA 1 issuer   warp 18 : MMA 2 issuer (+ the W2 chunk copies).  Two issuing warps because ONE was the
//   bottleneck: a clock64 trace (tools/exp_mlp2_trace.py) showed ~100 cycles per UTCHMMA issued and 150-500 per mbarrier wait,
//   3550 of the 3900 cycles of a chunk period on that single warp, with the epilogue warps waiting 1500 cycles for MMA 2.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace mlp2 {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kBlkBytes = kTileM * 128;
constexpr int kEpiWarps = 16;                   // 4 per TMEM lane quarter: the epilogue math (GELU, LayerNorm) is what bounds the kernel
constexpr int kThreads = (kEpiWarps + 3) * 32;
constexpr int kMaxStages = 4;
constexpr int kStageRow = 144;                  // staging row: 128 bytes of a 64-column block + 16 (bank stagger)
constexpr int kStageQBytes = 32 * kStageRow;    // per TMEM lane quarter
constexpr int kStageBytes = 4 * kStageQBytes;

struct alignas(64) Params {
  CUtensorMap tmap_x;               // x [m, c] bf16, box {64, 128}, SWIZZLE_128B
  const __nv_bfloat16 *w1_img;      // os3d_pack_linear_bf16 image of fc1.weight [h, c]: [ncb_c][h][64]
  const __nv_bfloat16 *w2_img;      // image of fc2.weight [c, h]: [ncb_h][c][64]
  const float *b1, *b2, *gamma, *beta;
  float ln_eps;
  const __nv_bfloat16 *x;           // residual
  __nv_bfloat16 *out;
  int64_t m;
  int c, h, hc, nj, ncb_c, ncb_hc, xb, n_tiles;
  int a1_stride, acc2_base;         // TMEM columns
  int s1, s2;                       // stages of the W1 / W2 rings
  uint32_t idesc1, idesc2, w1_bytes, w2_bytes;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

#ifdef OS3D_EXP_TRACE
__device__ long long *g_trace;
#define TRACE(role, c, k) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (c) < 64) g_trace[((role) * 64 + (c)) * 8 + (k)] = clock64(); } while (0)
#else
#define TRACE(role, c, k) do { } while (0)
#endif

__device__ __forceinline__ uint32_t sw128_off(int r, int col) {
  return (uint32_t)(col >> 6) * kBlkBytes + (uint32_t)r * 128u + ((((uint32_t)(col & 63) >> 3) ^ ((uint32_t)r & 7u)) << 4);
}

__global__ void __launch_bounds__(kThreads, 1) swformer_mlp_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const uint32_t x_bytes = (uint32_t)(p.ncb_c * kBlkBytes), h_bytes = (uint32_t)(p.ncb_hc * kBlkBytes);
  const uint32_t x_base = base;
  const uint32_t w1_base = x_base + p.xb * x_bytes;
  const uint32_t w2_base = w1_base + p.s1 * p.w1_bytes;
  const uint32_t h_base = w2_base + p.s2 * p.w2_bytes;
  uint8_t *h_ptr = smem + (h_base - base);
  uint8_t *stage_ptr = smem + (h_base - base) + 2u * h_bytes;        // epilogue staging (kStageBytes)
  uint8_t *tail = stage_ptr + kStageBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
  // x_full[2] x_empty[2] acc1_full[2] h_ready[2] h_free[2] acc2_full[2] acc2_free[2] w1_full[4] w1_free[4] w2_full[4] w2_free[4]
  const uint32_t x_full = smem_u32(bars), x_empty = smem_u32(bars + 2), acc1_full = smem_u32(bars + 4);
  const uint32_t h_ready = smem_u32(bars + 6), h_free = smem_u32(bars + 8);
  const uint32_t acc2_full = smem_u32(bars + 10), acc2_free = smem_u32(bars + 12);
  const uint32_t w1_full = smem_u32(bars + 14), w1_free = smem_u32(bars + 18), w2_full = smem_u32(bars + 22), w2_free = smem_u32(bars + 26);
  const uint32_t acc1_free = smem_u32(bars + 30);
  uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 32);
  float2 *part = reinterpret_cast<float2 *>(misc + 4);                // LayerNorm partial sums [8 warps][32 lanes]
  float *prm_s = reinterpret_cast<float *>(part + kEpiWarps * 32);    // b1 [h] | b2 [c] | gamma [c] | beta [c]
  const int o_b2 = p.h, o_g = p.h + p.c, o_be = p.h + 2 * p.c;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(x_full + 8 * i, 1);
      mbar_init(x_empty + 8 * i, 1);
      mbar_init(acc1_full + 8 * i, 1);
      mbar_init(h_ready + 8 * i, kEpiWarps);
      mbar_init(h_free + 8 * i, 1);
      mbar_init(acc1_free + 8 * i, kEpiWarps);
      mbar_init(acc2_full + 8 * i, 1);
      mbar_init(acc2_free + 8 * i, kEpiWarps);
    }
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(w1_full + 8 * i, 1);
      mbar_init(w1_free + 8 * i, 1);
      mbar_init(w2_full + 8 * i, 1);
      mbar_init(w2_free + 8 * i, 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < p.h; i += kThreads) prm_s[i] = p.b1 ? __ldg(p.b1 + i) : 0.0f;
  for (int i = tid; i < p.c; i += kThreads) {
    prm_s[o_b2 + i] = p.b2 ? __ldg(p.b2 + i) : 0.0f;
    prm_s[o_g + i] = __ldg(p.gamma + i);
    prm_s[o_be + i] = __ldg(p.beta + i);
  }
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(&misc[0]), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_chunks = n_my * p.nj;

  if (warp == kEpiWarps) {
    // ================================ producer ================================
    if (lane == 0) {
      auto load_x = [&](int t) {
        const uint32_t b = (uint32_t)(t % p.xb), use = (uint32_t)(t / p.xb);
        mbar_wait(x_empty + 8 * b, (use & 1u) ^ 1u);
        mbar_arrive_expect_tx(x_full + 8 * b, x_bytes);
        const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * kTileM;
        for (int cb = 0; cb < p.ncb_c; ++cb)
          tma_load_2d(x_base + b * x_bytes + cb * kBlkBytes, &p.tmap_x, cb * kBlockK, row0, x_full + 8 * b);
      };
      // W1 chunk c goes into stage c % s1 once MMA 1 of chunk c - s1 has read it.  Copies are issued in the order their
      // stages become free, and an x tile (free only when the LAST MMA 1 of a tile is done) never holds up W1 copies
      // that could already go.
      int w1_next = 0;
      auto issue_w1_upto = [&](int limit) {
        limit = limit < n_chunks ? limit : n_chunks;
        for (; w1_next < limit; ++w1_next) {
          const int c = w1_next, t = c / p.nj, j = c - t * p.nj;
          const uint32_t s = (uint32_t)(c % p.s1), use = (uint32_t)(c / p.s1);
          mbar_wait(w1_free + 8 * s, (use & 1u) ^ 1u);
#ifdef OS3D_EXP_NOCOPY
          mbar_arrive_expect_tx(w1_full + 8 * s, 16u * p.ncb_c);
          for (int cb = 0; cb < p.ncb_c; ++cb)
            bulk_g2s(w1_base + s * p.w1_bytes + cb * p.hc * 128, p.w1_img + ((int64_t)cb * p.h + (int64_t)j * p.hc) * kBlockK, 16u, w1_full + 8 * s);
#else
          mbar_arrive_expect_tx(w1_full + 8 * s, p.w1_bytes);
          for (int cb = 0; cb < p.ncb_c; ++cb)                             // W1 rows [j hc, j hc + hc) of K block cb
            bulk_g2s(w1_base + s * p.w1_bytes + cb * p.hc * 128, p.w1_img + ((int64_t)cb * p.h + (int64_t)j * p.hc) * kBlockK,
                     (uint32_t)(p.hc * 128), w1_full + 8 * s);
#endif
        }
      };
      for (int t = 0; t < p.xb && t < n_my; ++t) load_x(t);
      for (int t = 0; t < n_my; ++t) {
        issue_w1_upto((t + 1) * p.nj - 1 + p.s1);
        if (t + p.xb < n_my) load_x(t + p.xb);
      }
      issue_w1_upto(n_chunks);
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ================================ MMA 1 issuer ================================
    // acc1[c & 1] = x . W1[chunk c]^T.  Runs ahead of the epilogue as far as the two accumulators allow: each is handed
    // back right after the epilogue warps' tcgen05.ld of it (acc1_free), long before the GELU of that chunk is done.
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t x_lo0 = (uint32_t)make_kmajor_sw128_desc(x_base), w1_lo0 = (uint32_t)make_kmajor_sw128_desc(w1_base);
    const int c_steps = p.c >> 4;
    const uint32_t w1_blk = (uint32_t)(p.hc * 128) >> 4;
    for (int c = 0, t = 0, j = 0; c < n_chunks; ++c) {
      const uint32_t s = (uint32_t)c & 1u, xb = (uint32_t)(t % p.xb), ws = (uint32_t)(c % p.s1);
      TRACE(1, c, 0);
      if (c >= 2) mbar_wait(acc1_free + 8 * s, (((uint32_t)c >> 1) - 1u) & 1u);
      TRACE(1, c, 1);
      if (j == 0) mbar_wait(x_full + 8 * xb, (uint32_t)(t / p.xb) & 1u);
      mbar_wait(w1_full + 8 * ws, (uint32_t)(c / p.s1) & 1u);
      TRACE(1, c, 6);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + s * (uint32_t)p.a1_stride;
        uint32_t a = x_lo0 + xb * (x_bytes >> 4), b = w1_lo0 + ws * (p.w1_bytes >> 4);
        for (int ks = 0; ks < c_steps; ++ks) {
          umma_bf16_lo(d, a, b, desc_hi, p.idesc1, ks > 0 ? 1u : 0u);
          const bool wrap = (ks & 3) == 3;                    // next 64-column K block
          a += wrap ? (kBlkBytes >> 4) - 6 : 2;
          b += wrap ? w1_blk - 6 : 2;
        }
        umma_commit(w1_free + 8 * ws);
        umma_commit(acc1_full + 8 * s);
        if (j == p.nj - 1) umma_commit(x_empty + 8 * xb);
      }
      __syncwarp();
      TRACE(1, c, 2);
      j = j + 1 == p.nj ? 0 : j + 1;
      t += (j == 0);
    }
  } else if (warp == kEpiWarps + 2) {
    // ================================ MMA 2 issuer (+ W2 chunk copies) ================================
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t w2_lo0 = (uint32_t)make_kmajor_sw128_desc(w2_base), h_lo0 = (uint32_t)make_kmajor_sw128_desc(h_base);
    const int hc_steps = p.hc >> 4;
    const uint32_t w2_blk = (uint32_t)(p.c * 128) >> 4;
    // W2 chunk c goes into stage c % s2 once MMA 2 of chunk c - s2 has read it; issued one iteration after that MMA 2 was
    // committed, when the wait is (almost always) already satisfied
    auto load_w2 = [&](int c) {
      if (c >= n_chunks) return;
      if (lane == 0) {
        const int j = c % p.nj;
        const uint32_t s = (uint32_t)(c % p.s2), use = (uint32_t)(c / p.s2);
        mbar_wait(w2_free + 8 * s, (use & 1u) ^ 1u);
#ifdef OS3D_EXP_NOCOPY
        mbar_arrive_expect_tx(w2_full + 8 * s, 16u * p.ncb_hc);
        for (int b = 0; b < p.ncb_hc; ++b)
          bulk_g2s(w2_base + s * p.w2_bytes + b * p.c * 128, p.w2_img + ((int64_t)(j * p.ncb_hc + b) * p.c) * kBlockK, 16u, w2_full + 8 * s);
#else
        mbar_arrive_expect_tx(w2_full + 8 * s, p.w2_bytes);
        for (int b = 0; b < p.ncb_hc; ++b)                                 // W2 K blocks of the hidden chunk
          bulk_g2s(w2_base + s * p.w2_bytes + b * p.c * 128, p.w2_img + ((int64_t)(j * p.ncb_hc + b) * p.c) * kBlockK,
                   (uint32_t)(p.c * 128), w2_full + 8 * s);
#endif
      }
      __syncwarp();
    };
    for (int c = 0; c < p.s2; ++c) load_w2(c);
    for (int c = 0, t = 0, j = 0; c < n_chunks; ++c) {
      const uint32_t s = (uint32_t)c & 1u, ab = (uint32_t)t & 1u, ws = (uint32_t)(c % p.s2);
      TRACE(1, c, 3);
      if (c >= 1) load_w2(c - 1 + p.s2);
      TRACE(1, c, 4);
      if (j == 0 && t >= 2) mbar_wait(acc2_free + 8 * ab, (((uint32_t)t >> 1) - 1u) & 1u);
      mbar_wait(w2_full + 8 * ws, (uint32_t)(c / p.s2) & 1u);
      mbar_wait(h_ready + 8 * s, ((uint32_t)c >> 1) & 1u);
      TRACE(1, c, 7);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)p.acc2_base + ab * (uint32_t)p.c;
        uint32_t a = h_lo0 + s * (h_bytes >> 4), b = w2_lo0 + ws * (p.w2_bytes >> 4);
        for (int ks = 0; ks < hc_steps; ++ks) {
          umma_bf16_lo(d, a, b, desc_hi, p.idesc2, (j > 0 || ks > 0) ? 1u : 0u);
          const bool wrap = (ks & 3) == 3;
          a += wrap ? (kBlkBytes >> 4) - 6 : 2;
          b += wrap ? w2_blk - 6 : 2;
        }
        umma_commit(w2_free + 8 * ws);
        umma_commit(h_free + 8 * s);
        if (j == p.nj - 1) umma_commit(acc2_full + 8 * ab);
      }
      __syncwarp();
      TRACE(1, c, 5);
      j = j + 1 == p.nj ? 0 : j + 1;
      t += (j == 0);
    }
  } else {
    // ================================ epilogue warps ================================
    const int quarter = warp & 3, cpart = warp >> 2;      // lane quarter of TMEM; which 16 of every 64 columns
    const int r = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    // LayerNorm + residual epilogue of tile t.  A warp owns at most three 16-column chunks (c <= 192): they stay in
    // registers between the statistics and the normalisation, the residual rows are requested before the accumulator
    // is waited for, and the accumulator is handed back as soon as it has been read.
    // Global traffic of the LayerNorm epilogue goes through a small staging buffer per lane quarter (32 rows x one
    // 64-column block, rows padded to 144 bytes): a thread owns a ROW of the accumulator, so direct loads / stores touch 32
    // different 128-byte lines per instruction -- the clock64 trace put 6300 of the 8350 cycles of this epilogue there
    // (l1tex tag stage, not bytes).  Staged, 8 lanes cover one row's 128 bytes: 4 lines per instruction.
    uint8_t *stq = stage_ptr + quarter * kStageQBytes;
    const int g = cpart * 32 + lane;                     // thread index within the quarter's four warps
    auto final_epilogue = [&](int t) {
      const uint32_t ab = (uint32_t)t & 1u;
      const int64_t row0q = (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * kTileM + quarter * 32;
      const uint32_t t_row = tmem_base + (uint32_t)p.acc2_base + ab * (uint32_t)p.c + lane_sel;
      // this thread's two 16-byte pieces of every 32-row x 64-column block: (row, piece) = (g / 8, g % 8), (g / 8 + 16, g % 8)
      const int prow = g >> 3, piece = g & 7;
      uint4 rz[3][2];
      if (warp == 0) TRACE(2, t, 0);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        rz[i][0] = rz[i][1] = make_uint4(0, 0, 0, 0);
        const int col = 64 * i + piece * 8;
        if (col < p.c) {
          if (row0q + prow < p.m) rz[i][0] = __ldg(reinterpret_cast<const uint4 *>(p.x + (row0q + prow) * p.c + col));
          if (row0q + prow + 16 < p.m) rz[i][1] = __ldg(reinterpret_cast<const uint4 *>(p.x + (row0q + prow + 16) * p.c + col));
        }
      }
      mbar_wait(acc2_full + 8 * ab, ((uint32_t)t >> 1) & 1u);
      if (warp == 0) TRACE(2, t, 1);
      tc_fence_after();
      uint32_t v[3][16];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (cpart * 16 + 64 * i < p.c) tmem_ld16(t_row + (uint32_t)(cpart * 16 + 64 * i), v[i]);
      tmem_ld_wait();
      if (warp == 0) TRACE(2, t, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_free + 8 * ab);
      float sum = 0.0f, sq = 0.0f;                  // (values stay in v[][] as float bits: one register array, not two)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int col = cpart * 16 + 64 * i;
        if (col < p.c) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b = *reinterpret_cast<const float4 *>(prm_s + o_b2 + col + 4 * q4);
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float yy = __uint_as_float(v[i][4 * q4 + e]) + bb[e];
              v[i][4 * q4 + e] = __float_as_uint(yy);
              sum += yy;
              sq = fmaf(yy, yy, sq);
            }
          }
        }
      }
      // the four warps of a lane quarter exchange their partial sums through shared memory (128-thread named barrier)
      if (warp == 0) TRACE(2, t, 3);
      part[warp * 32 + lane] = make_float2(sum, sq);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      if (warp == 0) TRACE(2, t, 4);
      sum = 0.0f;
      sq = 0.0f;
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        const float2 o2 = part[(w4 * 4 + quarter) * 32 + lane];
        sum += o2.x;
        sq += o2.y;
      }
      if (warp == 0) TRACE(2, t, 5);
      const float mean = sum / (float)p.c;
      const float rstd = rsqrtf(fmaxf(sq / (float)p.c - mean * mean, 0.0f) + p.ln_eps);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (64 * i >= p.c) break;
        const int col = cpart * 16 + 64 * i;
        // residual block -> staging (also orders the previous block's reads / the statistics exchange before the writes)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
        *reinterpret_cast<uint4 *>(stq + prow * kStageRow + piece * 16) = rz[i][0];
        *reinterpret_cast<uint4 *>(stq + (prow + 16) * kStageRow + piece * 16) = rz[i][1];
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
        if (col < p.c) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 g4 = *reinterpret_cast<const float4 *>(prm_s + o_g + col + 4 * q4);
            const float4 be = *reinterpret_cast<const float4 *>(prm_s + o_be + col + 4 * q4);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, eb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              v[i][4 * q4 + e] = __float_as_uint(fmaf((__uint_as_float(v[i][4 * q4 + e]) - mean) * rstd, gg[e], eb[e]));
          }
          uint4 *mine = reinterpret_cast<uint4 *>(stq + lane * kStageRow + cpart * 32);
          const uint4 r0 = mine[0], r1 = mine[1];
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          uint32_t o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(v[i][2 * k]) + __uint_as_float(rw[k] << 16),
                                                            __uint_as_float(v[i][2 * k + 1]) + __uint_as_float(rw[k] & 0xffff0000u));
            o[k] = *reinterpret_cast<const uint32_t *>(&hh);
          }
          mine[0] = make_uint4(o[0], o[1], o[2], o[3]);
          mine[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
        const int ocol = 64 * i + piece * 8;
        if (ocol < p.c) {
          if (row0q + prow < p.m)
            *reinterpret_cast<uint4 *>(p.out + (row0q + prow) * p.c + ocol) = *reinterpret_cast<const uint4 *>(stq + prow * kStageRow + piece * 16);
          if (row0q + prow + 16 < p.m)
            *reinterpret_cast<uint4 *>(p.out + (row0q + prow + 16) * p.c + ocol) =
                *reinterpret_cast<const uint4 *>(stq + (prow + 16) * kStageRow + piece * 16);
        }
        if (warp == 0) TRACE(2, t, 6 + (i > 0));
      }
    };

    const int jf = p.nj > 1 ? 1 : 0;              // the chunk of tile t after which tile t - 1 gets its LayerNorm epilogue
    for (int c = 0, t = 0, j = 0; c < n_chunks; ++c, j = (j + 1 == p.nj ? 0 : j + 1), t += (j == 0)) {
      const uint32_t s = (uint32_t)c & 1u;
      if (warp == 0) TRACE(0, c, 0);
      mbar_wait(acc1_full + 8 * s, ((uint32_t)c >> 1) & 1u);
      if (warp == 0) TRACE(0, c, 1);
      if (c >= 2) mbar_wait(h_free + 8 * s, (((uint32_t)c >> 1) - 1u) & 1u);       // MMA 2 of chunk c - 2 has read h[s]
      tc_fence_after();
      if (warp == 0) TRACE(0, c, 2);
      const uint32_t t_row = tmem_base + s * (uint32_t)p.a1_stride + lane_sel;
      uint8_t *hb = h_ptr + s * h_bytes;
      const float *b1 = prm_s + j * p.hc;
      for (int col = cpart * 16; col < p.hc; col += 64) {
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)col, v);
        tmem_ld_wait();
        if (warp == 0) TRACE(0, c, 3);
        if (col + 64 >= p.hc) {                    // last read of this accumulator by this warp: hand it back to MMA 1
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc1_free + 8 * s);
        }
        uint32_t o[8];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 b = *reinterpret_cast<const float4 *>(b1 + col + 4 * q4);
#ifdef OS3D_EXP_NOGELU
          const float y0 = __uint_as_float(v[4 * q4 + 0]) + b.x, y1 = __uint_as_float(v[4 * q4 + 1]) + b.y;
          const float y2 = __uint_as_float(v[4 * q4 + 2]) + b.z, y3 = __uint_as_float(v[4 * q4 + 3]) + b.w;
#else
          float y0, y1, y2, y3;
          gelu_erf_fast2(__uint_as_float(v[4 * q4 + 0]), __uint_as_float(v[4 * q4 + 1]), b.x, b.y, y0, y1);
          gelu_erf_fast2(__uint_as_float(v[4 * q4 + 2]), __uint_as_float(v[4 * q4 + 3]), b.z, b.w, y2, y3);
#endif
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(y0, y1), h1 = __floats2bfloat162_rn(y2, y3);
          o[2 * q4] = *reinterpret_cast<const uint32_t *>(&h0);
          o[2 * q4 + 1] = *reinterpret_cast<const uint32_t *>(&h1);
        }
        *reinterpret_cast<uint4 *>(hb + sw128_off(r, col)) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(hb + sw128_off(r, col + 8)) = make_uint4(o[4], o[5], o[6], o[7]);
      }
      if (warp == 0) TRACE(0, c, 4);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready + 8 * s);
      if (warp == 0) TRACE(0, c, 5);
      if (j == jf && t >= 1) final_epilogue(t - 1);     // deferred: MMA 2 of the previous tile's last chunk has long finished
      if (warp == 0) TRACE(0, c, 6);
    }
    if (n_my > 0) final_epilogue(n_my - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)sym;
  }
  return fn;
}

struct Plan {
  int hc, nj, ncb_c, ncb_hc, xb, a1_stride, acc2_base, smem, tail, s1, s2;
  uint32_t w1_bytes, w2_bytes;
  bool ok;
};

static Plan plan(int c, int h) {
  Plan pl = {};
  if (c < 16 || c % 16 || c > 192 || h < 16 || h % 16 || h > 1024) return pl;
  int hc = 0;
  if (h % 64 == 0) hc = 64;
  else if (h <= 128) hc = h;               // one chunk: the whole hidden layer
  else return pl;
  pl.hc = hc;
  pl.nj = h / hc;
  pl.ncb_c = (int)cdiv(c, kBlockK);
  pl.ncb_hc = (int)cdiv(hc, kBlockK);
  pl.a1_stride = hc <= 64 ? 64 : 128;
  pl.acc2_base = 2 * pl.a1_stride;
  if (pl.acc2_base + 2 * c > 512) return pl;
  pl.w1_bytes = (uint32_t)(pl.ncb_c * hc * 128);
  pl.w2_bytes = (uint32_t)(pl.ncb_hc * c * 128);
  if (pl.w1_bytes % 1024 || pl.w2_bytes % 1024) return pl;      // UMMA operand bases: 1024-byte aligned (hc, c % 8 == 0)
  pl.tail = 32 * 8 + 16 + kEpiWarps * 32 * 8 + (h + 3 * c) * 4 + 64;
  const int x_bytes = pl.ncb_c * kBlkBytes, limit = 227 * 1024;
  auto total = [&](int xb, int s1, int s2) {
    return 1024 + xb * x_bytes + s1 * (int)pl.w1_bytes + s2 * (int)pl.w2_bytes + 2 * pl.ncb_hc * kBlkBytes + kStageBytes + pl.tail;
  };
  if (total(1, 2, 2) > limit) return pl;
  pl.xb = 1; pl.s1 = 2; pl.s2 = 2;
  // spend what is left on the W1 ring first (its copy latency is on the critical path), then x, then W2
  if (total(pl.xb, 3, pl.s2) <= limit) pl.s1 = 3;
  if (total(2, pl.s1, pl.s2) <= limit) pl.xb = 2;
  if (total(pl.xb, pl.s1, 3) <= limit) pl.s2 = 3;
  if (pl.s1 == 3 && total(pl.xb, 4, pl.s2) <= limit) pl.s1 = 4;
  if (pl.s2 == 3 && total(pl.xb, pl.s1, 4) <= limit) pl.s2 = 4;
  pl.smem = total(pl.xb, pl.s1, pl.s2);
  pl.ok = true;
  return pl;
}

}  // namespace mlp2
}  // namespace os3d

using namespace os3d;

#ifdef OS3D_EXP_TRACE
extern "C" int os3d_exp_set_trace(long long *ptr) { return (int)cudaMemcpyToSymbol(mlp2::g_trace, &ptr, sizeof(ptr)); }
#endif

extern "C" int os3d_swformer_mlp_fits(int c, int h) { return mlp2::plan(c, h).ok ? 1 : 0; }

extern "C" int os3d_swformer_mlp_bf16(const void *x, int64_t m, int c, int h, const void *w1, const float *b1,
                                      const void *w2, const float *b2, const float *ln_gamma, const float *ln_beta,
                                      float ln_eps, void *out, void *stream) {
  const mlp2::Plan pl = mlp2::plan(c, h);
  if (!pl.ok || m < 0 || !w1 || !w2 || !ln_gamma || !ln_beta || ((uintptr_t)x & 15) || ((uintptr_t)out & 15))
    return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  mlp2::encode_tiled_fn enc = mlp2::encode_tiled();
  if (!enc) return OS3D_ERR_BAD_ARG;
  mlp2::Params p;
  memset(&p, 0, sizeof(p));
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)c, (cuuint64_t)m};
    const cuuint64_t gstr[1] = {(cuuint64_t)c * 2};
    const cuuint32_t box[2] = {(cuuint32_t)mlp2::kBlockK, (cuuint32_t)mlp2::kTileM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&p.tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OS3D_ERR_BAD_ARG;
  }
  p.w1_img = (const __nv_bfloat16 *)w1;
  p.w2_img = (const __nv_bfloat16 *)w2;
  p.b1 = b1; p.b2 = b2; p.gamma = ln_gamma; p.beta = ln_beta; p.ln_eps = ln_eps;
  p.x = (const __nv_bfloat16 *)x;
  p.out = (__nv_bfloat16 *)out;
  p.m = m;
  p.c = c; p.h = h; p.hc = pl.hc; p.nj = pl.nj; p.ncb_c = pl.ncb_c; p.ncb_hc = pl.ncb_hc; p.xb = pl.xb;
  p.n_tiles = (int)cdiv(m, mlp2::kTileM);
  p.a1_stride = pl.a1_stride;
  p.acc2_base = pl.acc2_base;
  p.idesc1 = ptx::make_idesc_bf16(mlp2::kTileM, pl.hc);
  p.idesc2 = ptx::make_idesc_bf16(mlp2::kTileM, c);
  p.s1 = pl.s1; p.s2 = pl.s2;
  p.w1_bytes = pl.w1_bytes;
  p.w2_bytes = pl.w2_bytes;
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: once per device, not once per process
  static bool configured[64] = {false};
  int cfg_dev = 0;
  OS3D_CUDA(cudaGetDevice(&cfg_dev));
  if (cfg_dev < 0 || cfg_dev >= 64 || !configured[cfg_dev]) {
    OS3D_CUDA(cudaFuncSetAttribute(mlp2::swformer_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (cfg_dev >= 0 && cfg_dev < 64) configured[cfg_dev] = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  mlp2::swformer_mlp_tc_kernel<<<grid, mlp2::kThreads, pl.smem, (cudaStream_t)stream>>>(p);
  OS3D_LAUNCH_CHECK();
  return 0;
}
