// mlp_tc.cu -- a chain of Linear layers in ONE persistent tcgen05 kernel, activations kept on chip:
//
//   y = L_{n-1}( ... act_1( L_1( act_0( L_0(x) ) ) ) )       L_l(a) = a . W_l^T + b_l,  act = none | ReLU | exact GELU
//
// with an optional fp32 SIMT layer in front (raw point coordinates do not survive a bf16 cast) and an optional
// residual + LayerNorm on the last layer.  Two users:
//   * the point-wise MLPs of Segformer with eval-mode BatchNorm folded into the weights (point_encoder
//     D -> 64 -> 128 -> 256 -> 64, fusion_encoder 96 -> 256 -> 128 -> 64, classifier 64 -> 64 -> classes:
//     seg3d/models/segmentors/segformer.py:21-32,58-76): the [N, 256] intermediates never reach HBM;
//   * the SWFormer MLP  x + LayerNorm(fc2(GELU(fc1(x))))  (point_transformer_layer.py:260-298) for C <= 96, where both
//     weight matrices fit in shared memory: the [M, 2C] hidden tensor (written, read by the GELU kernel, written, read by
//     fc2) disappears.
//
// One CTA per SM walks the 128-row tiles.  All weight images stay resident in shared memory (loaded once, cp.async.bulk).
// Per tile, layer l:  MMA warp: A = activation buffer (or the TMA-loaded x slots for layer 0), B = W_l, accumulator l & 1
// of two in tensor memory -> commit;  8 epilogue warps: tcgen05.ld, bias / activation, bf16, st.shared into the
// activation buffer in the SWIZZLE_128B K-major layout the next layer's UMMA descriptor reads (fence.proxy.async, then an
// mbarrier arrival hands it to the MMA warp).  The last layer's epilogue writes global memory (optionally + residual,
// LayerNorm).  x tiles of the next row tile are prefetched by TMA while the chain runs (when shared memory allows; the
// widest chain loads x straight into the activation buffer instead).
//
//   warp 8 : TMA producer      warp 9 : MMA issuer (warp-uniform, tcgen05 under elect.sync)      warps 0-7 : epilogue
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace mlp {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kBlkBytes = kTileM * 128;        // one 64-column K block of a 128-row bf16 tile
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kMaxSlots = 8;
constexpr int kMaxLayers = 4;
constexpr int kN0 = 64;                         // width of the fp32 front layer

struct Layer {
  const __nv_bfloat16 *w_img;   // [ncb][n][64] swizzled (os3d_pack_linear_bf16)
  const float *bias;            // [n] or null
  int k, n, act, ncb;
  uint32_t w_off, prm_off, idesc;   // byte offset of the image in smem, float offset of the bias in prm_s
};

struct alignas(64) Params {
  CUtensorMap tmap_x;           // bf16 input [m, k0] (pitch ldx), box {64, 128}, SWIZZLE_128B
  Layer layer[kMaxLayers];
  int n_layers;
  // fp32 front layer (x32 != null): a0 = relu(x32[:, :k32] . w32 + b32), 64 wide
  const float *x32, *w32, *b32;   // w32: [k32][64]
  int64_t ld32;
  int k32, act32;
  const __nv_bfloat16 *residual;  // [m, n_last] pitch ldr
  int64_t ldr;
  const float *ln_gamma, *ln_beta;
  float ln_eps;
  void *out;                      // bf16 (or fp32 when out_f32) [m, n_out] pitch ldo
  int64_t m, ldo;
  int n_out, out_f32;
  int n_tiles, slots, act_blocks, tmem_cols, acc_stride;
  uint32_t w_bytes_total;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// byte offset of the 16-byte chunk holding columns [col, col + 8) of row r in a K-major SWIZZLE_128B tile of 64-column blocks
__device__ __forceinline__ uint32_t sw128_off(int r, int col) {
  return (uint32_t)(col >> 6) * kBlkBytes + (uint32_t)r * 128u + ((((uint32_t)(col & 63) >> 3) ^ ((uint32_t)r & 7u)) << 4);
}

__global__ void __launch_bounds__(kThreads, 1) mlp_chain_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const uint32_t w_base = base;
  const uint32_t act_base = base + p.w_bytes_total;
  uint8_t *act_ptr = smem + p.w_bytes_total;
  const uint32_t a_base = act_base + p.act_blocks * kBlkBytes;           // x slots (ring mode)
  uint8_t *tail = smem + p.w_bytes_total + (p.act_blocks + p.slots) * kBlkBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);   // a_full[8] a_empty[8] w_full act_ready acc_full[2] act_free
  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + kMaxSlots), w_full = smem_u32(bars + 2 * kMaxSlots);
  const uint32_t act_ready = smem_u32(bars + 2 * kMaxSlots + 1), acc_full = smem_u32(bars + 2 * kMaxSlots + 2);
  const uint32_t act_free = smem_u32(bars + 2 * kMaxSlots + 4);
  uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * kMaxSlots + 6);       // [0] tmem base
  float2 *part = reinterpret_cast<float2 *>(misc + 4);      // LayerNorm partial sums [8 warps][32 lanes] (16-byte aligned from here on)
  float *prm_s = reinterpret_cast<float *>(part + kEpiWarps * 32);                // biases | gamma | beta | w32 | b32

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const Layer &last = p.layer[L - 1];
  const bool front = p.x32 != nullptr;
  const bool direct = !front && p.slots == 0;      // x tiles land in the activation buffer (no ring)
  int prm_gamma = 0;
  for (int l = 0; l < L; ++l) prm_gamma += p.layer[l].n;
  const int prm_beta = prm_gamma + last.n, prm_w32 = prm_beta + last.n, prm_b32 = prm_w32 + 16 * kN0;

  if (tid == 0) {
    for (int s = 0; s < kMaxSlots; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
    mbar_init(w_full, 1);
    mbar_init(act_ready, kEpiWarps);
    mbar_init(acc_full, 1);
    mbar_init(acc_full + 8, 1);
    mbar_init(act_free, 1);
    fence_barrier_init();
  }
  for (int l = 0; l < L; ++l)
    for (int i = tid; i < p.layer[l].n; i += kThreads)
      prm_s[p.layer[l].prm_off + i] = p.layer[l].bias ? __ldg(p.layer[l].bias + i) : 0.0f;
  if (p.ln_gamma)
    for (int i = tid; i < last.n; i += kThreads) {
      prm_s[prm_gamma + i] = __ldg(p.ln_gamma + i);
      prm_s[prm_beta + i] = __ldg(p.ln_beta + i);
    }
  if (front) {
    for (int i = tid; i < p.k32 * kN0; i += kThreads) prm_s[prm_w32 + i] = __ldg(p.w32 + i);
    for (int i = tid; i < kN0; i += kThreads) prm_s[prm_b32 + i] = p.b32 ? __ldg(p.b32 + i) : 0.0f;
  }
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(&misc[0]), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // tiles of this CTA
  const int ncb0 = p.layer[0].ncb;

  if (warp == kEpiWarps) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, p.w_bytes_total);
      for (int l = 0; l < L; ++l)
        for (int cb = 0; cb < p.layer[l].ncb; ++cb)
          bulk_g2s(w_base + p.layer[l].w_off + cb * p.layer[l].n * 128, p.layer[l].w_img + (int64_t)cb * p.layer[l].n * kBlockK,
                   (uint32_t)(p.layer[l].n * 128), w_full);
      if (!front) {
        uint32_t q = 0;
        for (int it = 0; it < n_my; ++it) {
          const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * kTileM;
          if (direct) {
            // the activation buffer is free once the last layer's MMAs of the previous tile have completed (its own
            // barrier: a parity wait is only meaningful one phase ahead, and acc_full runs several phases per tile)
            if (it > 0) mbar_wait(act_free, (uint32_t)(it - 1) & 1u);
            mbar_arrive_expect_tx(a_full, (uint32_t)(ncb0 * kBlkBytes));
            for (int cb = 0; cb < ncb0; ++cb) tma_load_2d(act_base + cb * kBlkBytes, &p.tmap_x, cb * kBlockK, row0, a_full);
          } else {
            for (int cb = 0; cb < ncb0; ++cb, ++q) {
              const uint32_t slot = q % (uint32_t)p.slots, ph = ((q / (uint32_t)p.slots) & 1u) ^ 1u;
              mbar_wait(a_empty + 8 * slot, ph);
              mbar_arrive_expect_tx(a_full + 8 * slot, (uint32_t)kBlkBytes);
              tma_load_2d(a_base + slot * kBlkBytes, &p.tmap_x, cb * kBlockK, row0, a_full + 8 * slot);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ================================ MMA issuer ================================
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t act_lo0 = (uint32_t)make_kmajor_sw128_desc(act_base), ring_lo0 = (uint32_t)make_kmajor_sw128_desc(a_base);
    const uint32_t w_lo0 = (uint32_t)make_kmajor_sw128_desc(w_base);
    mbar_wait(w_full, 0);
    uint32_t q = 0, n_act = 0, gl = 0;
    for (int it = 0; it < n_my; ++it) {
      for (int l = 0; l < L; ++l, ++gl) {
        const Layer &ly = p.layer[l];
        const uint32_t d = tmem_base + (gl & 1u) * (uint32_t)p.acc_stride;
        const uint32_t w_lo = w_lo0 + (ly.w_off >> 4), w_step = (uint32_t)(ly.n * 128) >> 4;
        const int steps_total = ly.k >> 4;
        if (l == 0 && !front && !direct) {
          for (int cb = 0; cb < ncb0; ++cb, ++q) {
            const uint32_t slot = q % (uint32_t)p.slots, ph = (q / (uint32_t)p.slots) & 1u;
            const int steps = min(kBlockK / 16, steps_total - cb * (kBlockK / 16));
            mbar_wait(a_full + 8 * slot, ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = ring_lo0 + slot * (kBlkBytes >> 4), b_lo = w_lo + (uint32_t)cb * w_step;
#pragma unroll
              for (int ks = 0; ks < kBlockK / 16; ++ks)
                if (ks < steps) umma_bf16_lo(d, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, ly.idesc, (cb > 0 || ks > 0) ? 1u : 0u);
              umma_commit(a_empty + 8 * slot);
              if (cb + 1 == ncb0) umma_commit(acc_full + 8 * (gl & 1u));
            }
            __syncwarp();
          }
        } else {
          if (l == 0 && direct) {
            mbar_wait(a_full, (uint32_t)it & 1u);
          } else {
            mbar_wait(act_ready, n_act & 1u);         // the epilogue warps have written this layer's input
            ++n_act;
          }
          tc_fence_after();
          if (elect_one()) {
            for (int cb = 0; cb < ly.ncb; ++cb) {
              const int steps = min(kBlockK / 16, steps_total - cb * (kBlockK / 16));
              const uint32_t a_lo = act_lo0 + (uint32_t)cb * (kBlkBytes >> 4), b_lo = w_lo + (uint32_t)cb * w_step;
#pragma unroll
              for (int ks = 0; ks < kBlockK / 16; ++ks)
                if (ks < steps) umma_bf16_lo(d, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, ly.idesc, (cb > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(acc_full + 8 * (gl & 1u));
            if (direct && l + 1 == L) umma_commit(act_free);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int quarter = warp & 3, half = warp >> 2;      // TMEM lanes [32 q, 32 q + 32): warps q and q + 4 take alternate chunks
    const int r = quarter * 32 + lane;                   // row of the tile
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    uint32_t gl = 0;
    for (int it = 0; it < n_my; ++it) {
      const int64_t row = (int64_t)((int)blockIdx.x + it * (int)gridDim.x) * kTileM + r;
      const bool row_ok = row < p.m;
      if (front) {
        // ---- fp32 front layer on the FP32 pipes: this thread computes 32 of the 64 outputs of its row ----
        // (the previous tile's last MMAs have completed: this warp waited for their accumulator below)
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = prm_s[prm_b32 + half * 32 + j];
        if (row_ok) {
          const float *xr = p.x32 + row * p.ld32;
          for (int dd = 0; dd < p.k32; ++dd) {
            const float xv = __ldg(xr + dd);
            const float4 *wr = reinterpret_cast<const float4 *>(prm_s + prm_w32 + dd * kN0 + half * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w = wr[j4];
              acc[4 * j4 + 0] = fmaf(xv, w.x, acc[4 * j4 + 0]);
              acc[4 * j4 + 1] = fmaf(xv, w.y, acc[4 * j4 + 1]);
              acc[4 * j4 + 2] = fmaf(xv, w.z, acc[4 * j4 + 2]);
              acc[4 * j4 + 3] = fmaf(xv, w.w, acc[4 * j4 + 3]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = acc[8 * c + 2 * i], b = acc[8 * c + 2 * i + 1];
            if (p.act32 == 1) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
            if (!row_ok) { a = 0.0f; b = 0.0f; }
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
            o[i] = *reinterpret_cast<const uint32_t *>(&h);
          }
          *reinterpret_cast<uint4 *>(act_ptr + sw128_off(r, half * 32 + 8 * c)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(act_ready);
      }
      for (int l = 0; l < L; ++l, ++gl) {
        const Layer &ly = p.layer[l];
        const uint32_t t_row = tmem_base + (gl & 1u) * (uint32_t)p.acc_stride + ((uint32_t)(quarter * 32) << 16);
        const float *bias_s = prm_s + ly.prm_off;
        // acc + bias for columns [col, col + 16) of this thread's row
        auto add_bias = [&](const uint32_t (&v)[16], int col, float (&y)[16]) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b = *reinterpret_cast<const float4 *>(bias_s + col + 4 * q4);
            y[4 * q4 + 0] = __uint_as_float(v[4 * q4 + 0]) + b.x;
            y[4 * q4 + 1] = __uint_as_float(v[4 * q4 + 1]) + b.y;
            y[4 * q4 + 2] = __uint_as_float(v[4 * q4 + 2]) + b.z;
            y[4 * q4 + 3] = __uint_as_float(v[4 * q4 + 3]) + b.w;
          }
        };
        auto load16 = [&](int col, float (&y)[16]) {
          uint32_t v[16];
          tmem_ld16(t_row + (uint32_t)col, v);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b = *reinterpret_cast<const float4 *>(bias_s + col + 4 * q4);
            y[4 * q4 + 0] = __uint_as_float(v[4 * q4 + 0]) + b.x;
            y[4 * q4 + 1] = __uint_as_float(v[4 * q4 + 1]) + b.y;
            y[4 * q4 + 2] = __uint_as_float(v[4 * q4 + 2]) + b.z;
            y[4 * q4 + 3] = __uint_as_float(v[4 * q4 + 3]) + b.w;
          }
        };
        if (l + 1 < L) {
          // ---- hidden layer: activation -> bf16 -> the next layer's A operand in shared memory ----
          mbar_wait(acc_full + 8 * (gl & 1u), (gl >> 1) & 1u);
          tc_fence_after();
          // tcgen05.ld of the next chunk is in flight while this one is processed (two warps per scheduler do not hide
          // the TMEM read latency by themselves)
          auto hidden16 = [&](const uint32_t (&v)[16], int col) {
            float y[16];
            add_bias(v, col, y);
            if (ly.act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.0f);
            } else if (ly.act == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = gelu_erf_fast(y[i]);
            }
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
              o[i] = *reinterpret_cast<const uint32_t *>(&h);
            }
            *reinterpret_cast<uint4 *>(act_ptr + sw128_off(r, col)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4 *>(act_ptr + sw128_off(r, col + 8)) = make_uint4(o[4], o[5], o[6], o[7]);
          };
          uint32_t va[16], vb[16];
          int col = half * 16;
          if (col < ly.n) tmem_ld16(t_row + (uint32_t)col, va);
          for (; col < ly.n; col += 64) {
            const bool has_b = col + 32 < ly.n;
            tmem_ld_wait();
            if (has_b) tmem_ld16(t_row + (uint32_t)(col + 32), vb);
            hidden16(va, col);
            if (has_b) {
              tmem_ld_wait();
              if (col + 64 < ly.n) tmem_ld16(t_row + (uint32_t)(col + 64), va);
              hidden16(vb, col + 32);
            }
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(act_ready);
        } else {
          // ---- last layer: (+ residual, LayerNorm) -> global memory ----
          const bool do_ln = p.ln_gamma != nullptr;
          const __nv_bfloat16 *rrow = (p.residual && row_ok) ? p.residual + row * p.ldr : nullptr;
          uint4 nx[2] = {zero4, zero4};
          auto fetch = [&](int col) {
            if (rrow) {
              nx[0] = __ldg(reinterpret_cast<const uint4 *>(rrow + col));
              nx[1] = __ldg(reinterpret_cast<const uint4 *>(rrow + col) + 1);
            }
          };
          if (half * 16 < ly.n) fetch(half * 16);
          mbar_wait(acc_full + 8 * (gl & 1u), (gl >> 1) & 1u);
          tc_fence_after();
          float mean = 0.0f, rstd = 1.0f;
          if (do_ln) {
            float sum = 0.0f, sq = 0.0f;
            for (int col = half * 16; col < ly.n; col += 32) {
              float y[16];
              load16(col, y);
#pragma unroll
              for (int i = 0; i < 16; ++i) { sum += y[i]; sq = fmaf(y[i], y[i], sq); }
            }
            part[warp * 32 + lane] = make_float2(sum, sq);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            const float2 other = part[(warp ^ 4) * 32 + lane];
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            sum += other.x;
            sq += other.y;
            mean = sum / (float)ly.n;
            rstd = rsqrtf(fmaxf(sq / (float)ly.n - mean * mean, 0.0f) + p.ln_eps);
          }
          const bool vec_ok = !p.out_f32 && (p.ldo % 8 == 0) && (p.n_out % 16 == 0);
          for (int col = half * 16; col < ly.n; col += 32) {
            const uint4 c0 = nx[0], c1 = nx[1];
            if (col + 32 < ly.n) fetch(col + 32);
            float y[16];
            load16(col, y);
            if (!row_ok || col >= p.n_out) continue;
            if (do_ln) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float4 g = *reinterpret_cast<const float4 *>(prm_s + prm_gamma + col + 4 * q4);
                const float4 be = *reinterpret_cast<const float4 *>(prm_s + prm_beta + col + 4 * q4);
                y[4 * q4 + 0] = fmaf((y[4 * q4 + 0] - mean) * rstd, g.x, be.x);
                y[4 * q4 + 1] = fmaf((y[4 * q4 + 1] - mean) * rstd, g.y, be.y);
                y[4 * q4 + 2] = fmaf((y[4 * q4 + 2] - mean) * rstd, g.z, be.z);
                y[4 * q4 + 3] = fmaf((y[4 * q4 + 3] - mean) * rstd, g.w, be.w);
              }
            }
            if (rrow) {
              const uint32_t rw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                y[2 * i] += __uint_as_float(rw[i] << 16);
                y[2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
              }
            }
            if (ly.act == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.0f);
            } else if (ly.act == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = gelu_erf_fast(y[i]);
            }
            if (vec_ok) {
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
                o[i] = *reinterpret_cast<const uint32_t *>(&h);
              }
              __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.out) + row * p.ldo + col;
              reinterpret_cast<uint4 *>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
              reinterpret_cast<uint4 *>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else if (p.out_f32) {
              float *dst = reinterpret_cast<float *>(p.out) + row * p.ldo + col;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col + i < p.n_out) dst[i] = y[i];
            } else {
              __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(p.out) + row * p.ldo + col;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col + i < p.n_out) dst[i] = __float2bfloat16_rn(y[i]);
            }
          }
          tc_fence_before();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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
  int act_blocks, slots, tail, smem, tmem_cols, acc_stride;
  uint32_t w_bytes;
  bool ok;
};

static Plan plan(const os3d_mlp_layer *layers, int n_layers, bool front) {
  Plan pl = {};
  if (n_layers < 2 || n_layers > kMaxLayers) return pl;
  int n_max = 0, k_act = front ? kN0 : 0, n_sum = 0;
  uint32_t w_bytes = 0;
  for (int l = 0; l < n_layers; ++l) {
    const int k = layers[l].k, n = layers[l].n;
    if (k <= 0 || k % 16 || k > 256 || n < 16 || n % 16 || n > 256) return pl;
    if (l > 0 && k != layers[l - 1].n) return pl;
    if (layers[l].act < 0 || layers[l].act > 2 || !layers[l].w) return pl;
    w_bytes += (uint32_t)(cdiv(k, kBlockK) * n * 128);
    n_max = n > n_max ? n : n_max;
    n_sum += n;
    if (l > 0 && k > k_act) k_act = k;
  }
  if (front && layers[0].k != kN0) return pl;
  int stride = 32;
  while (stride < n_max) stride *= 2;
  pl.acc_stride = stride;
  pl.tmem_cols = 2 * stride;
  pl.w_bytes = w_bytes;
  const int ncb0 = (int)cdiv(layers[0].k, kBlockK);
  pl.tail = (2 * kMaxSlots + 6) * 8 + 16 + kEpiWarps * 32 * 8 + (n_sum + 2 * layers[n_layers - 1].n + 17 * kN0) * 4 + 64;
  const int budget = 227 * 1024 - 1024 - pl.tail - (int)w_bytes;
  int act_blocks = (int)cdiv(k_act, kBlockK);
  int slots = 0;
  if (!front) {
    slots = (budget - act_blocks * kBlkBytes) / kBlkBytes;
    slots = slots > kMaxSlots ? kMaxSlots : slots;
    if (slots < ncb0) {                  // no room for a ring: x goes straight into the activation buffer
      slots = 0;
      if (act_blocks < ncb0) act_blocks = ncb0;
    }
  }
  if (budget < act_blocks * kBlkBytes) return pl;
  pl.act_blocks = act_blocks;
  pl.slots = slots;
  pl.smem = 1024 + (int)w_bytes + (act_blocks + slots) * kBlkBytes + pl.tail;
  pl.ok = true;
  return pl;
}

}  // namespace mlp
}  // namespace os3d

using namespace os3d;

// 1 when os3d_mlp_chain_bf16 can run the chain (2..4 layers, widths multiples of 16 up to 256, all weight images plus one
// activation tile fit in shared memory).
extern "C" int os3d_mlp_chain_fits(const os3d_mlp_layer *layers, int n_layers, int has_front) {
  if (!layers) return 0;
  return mlp::plan(layers, n_layers, has_front != 0).ok ? 1 : 0;
}

extern "C" int os3d_mlp_chain_bf16(const void *x, int64_t m, int64_t ldx, const float *x32, int64_t ld32, int k32,
                                   const float *w32, const float *b32, int act32, const os3d_mlp_layer *layers,
                                   int n_layers, const void *residual, int64_t ldr, const float *ln_gamma,
                                   const float *ln_beta, float ln_eps, void *out, int64_t ldo, int n_out, int out_f32,
                                   void *stream) {
  const bool front = x32 != nullptr;
  if (!layers || m < 0 || (front == (x != nullptr))) return OS3D_ERR_BAD_ARG;
  const mlp::Plan pl = mlp::plan(layers, n_layers, front);
  if (!pl.ok) return OS3D_ERR_BAD_ARG;
  const os3d_mlp_layer &last = layers[n_layers - 1];
  if (front && (k32 <= 0 || k32 > 16 || !w32 || ld32 < k32 || act32 < 0 || act32 > 1)) return OS3D_ERR_BAD_ARG;
  if (!front && (((uintptr_t)x & 15) || ldx % 8 || ldx < layers[0].k)) return OS3D_ERR_BAD_ARG;
  if (n_out <= 0 || n_out > last.n || ldo < n_out || !out) return OS3D_ERR_BAD_ARG;
  if ((ln_gamma == nullptr) != (ln_beta == nullptr)) return OS3D_ERR_BAD_ARG;
  if (residual && (ldr % 8 || ldr < last.n || ((uintptr_t)residual & 15))) return OS3D_ERR_BAD_ARG;
  if (!out_f32 && ldo % 8 == 0 && n_out % 16 == 0 && ((uintptr_t)out & 15)) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  mlp::Params p;
  memset(&p, 0, sizeof(p));
  if (!front) {
    mlp::encode_tiled_fn enc = mlp::encode_tiled();
    if (!enc) return OS3D_ERR_BAD_ARG;
    const cuuint64_t gdim[2] = {(cuuint64_t)layers[0].k, (cuuint64_t)m};
    const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)mlp::kBlockK, (cuuint32_t)mlp::kTileM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&p.tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OS3D_ERR_BAD_ARG;
  }
  uint32_t w_off = 0, prm_off = 0;
  for (int l = 0; l < n_layers; ++l) {
    mlp::Layer &ly = p.layer[l];
    ly.w_img = (const __nv_bfloat16 *)layers[l].w;
    ly.bias = layers[l].bias;
    ly.k = layers[l].k;
    ly.n = layers[l].n;
    ly.act = layers[l].act;
    ly.ncb = (int)cdiv(ly.k, mlp::kBlockK);
    ly.w_off = w_off;
    ly.prm_off = prm_off;
    ly.idesc = ptx::make_idesc_bf16(mlp::kTileM, ly.n);
    w_off += (uint32_t)(ly.ncb * ly.n * 128);
    prm_off += (uint32_t)ly.n;
  }
  p.n_layers = n_layers;
  p.x32 = x32; p.w32 = w32; p.b32 = b32; p.ld32 = ld32; p.k32 = k32; p.act32 = act32;
  p.residual = (const __nv_bfloat16 *)residual;
  p.ldr = ldr;
  p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.ln_eps = ln_eps;
  p.out = out; p.m = m; p.ldo = ldo; p.n_out = n_out; p.out_f32 = out_f32;
  p.n_tiles = (int)cdiv(m, mlp::kTileM);
  p.slots = pl.slots;
  p.act_blocks = pl.act_blocks;
  p.tmem_cols = pl.tmem_cols;
  p.acc_stride = pl.acc_stride;
  p.w_bytes_total = pl.w_bytes;
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: once per device, not once per process
  static bool configured[64] = {false};
  int cfg_dev = 0;
  OS3D_CUDA(cudaGetDevice(&cfg_dev));
  if (cfg_dev < 0 || cfg_dev >= 64 || !configured[cfg_dev]) {
    OS3D_CUDA(cudaFuncSetAttribute(mlp::mlp_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (cfg_dev >= 0 && cfg_dev < 64) configured[cfg_dev] = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  mlp::mlp_chain_tc_kernel<<<grid, mlp::kThreads, pl.smem, (cudaStream_t)stream>>>(p);
  OS3D_LAUNCH_CHECK();
  return 0;
}
