// linear_tc.cu -- persistent dense Linear on tcgen05 for the SWFormer encoder layer:
//
//   out[r, :n] = epi( x[r, :k] . W^T + bias ),   epi = ReLU | exact GELU | residual + LayerNorm | table add
//
// os3d_linear_bf16 first ran on the sparse-conv kernel's dense mode (spconv_tc.cu): one short CTA per 128-row tile,
// weights re-fetched from L2 for every tile, epilogue after the mainloop on the same warps.  ncu showed that design
// latency-bound for Linear layers (21 us per CTA of which the MMAs are 1 us; as much weight traffic as activation
// traffic).  This kernel is the Linear-specific shape of the same pipeline:
//
//   * one persistent CTA per SM walks the row tiles (static round robin);
//   * the weight image W [ncb][n][128 B, swizzled] is loaded ONCE per CTA (cp.async.bulk) and stays in shared memory;
//   * activations come by TMA 2-D tile loads (cp.async.bulk.tensor.2d, box 64 x 128, SWIZZLE_128B = the UMMA K-major
//     layout; rows past m and columns past k are zero-filled by the TMA unit) through a ring of 16 KB slots;
//   * two accumulators in tensor memory (2 x n <= 512 columns): the MMA warp fills one while the 8 epilogue warps drain
//     the other, so loads, MMAs and the (global-latency-bound) epilogue of different tiles overlap inside the CTA.
//
//   warp 8 : TMA producer      warp 9 : MMA issuer (warp-uniform, tcgen05 under elect.sync)      warps 0-7 : epilogue
//
// replaces: nn.Linear + GELU / LayerNorm / residual of EncoderLayer.forward and the attention output projection
//           (seg3d/models/layers/point_transformer_layer.py:260-298, seg3d/models/layers/cosine_msa.py:403).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace lin {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kSlotBytes = kTileM * 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kMaxSlots = 8;

struct alignas(64) Params {
  CUtensorMap tmap_x;         // x [m, k] bf16, box {64, 128}, SWIZZLE_128B, zero OOB fill
  const __nv_bfloat16 *w_img; // [ncb][n][64] swizzled weight image (os3d_pack_linear_bf16)
  const float *bias;
  const __nv_bfloat16 *residual;
  const float *ln_gamma, *ln_beta;
  float ln_eps;
  const __nv_bfloat16 *table;
  const int32_t *tab_idx;
  int tab_cols;
  __nv_bfloat16 *out;
  int64_t m, ldo;
  int n_tiles, k16, n, ncb, flags, slots, n_parts, n_per_part;
  uint32_t idesc;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) linear_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const int w_bytes = p.ncb * p.n * 128;
  const uint32_t w_base = base;
  const uint32_t a_base = base + w_bytes;                          // w_bytes is a multiple of 1024 (n % 8 == 0)
  uint8_t *tail = smem + w_bytes + p.slots * kSlotBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);            // a_full[8] a_empty[8] w_full acc_full[2] acc_empty[2]
  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + kMaxSlots), w_full = smem_u32(bars + 2 * kMaxSlots);
  const uint32_t acc_full = smem_u32(bars + 2 * kMaxSlots + 1), acc_empty = smem_u32(bars + 2 * kMaxSlots + 3);
  uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * kMaxSlots + 5);       // [0] tmem base
  float *prm_s = reinterpret_cast<float *>(misc + 2);              // [bias | gamma | beta], n floats each (16-byte aligned)
  float2 *part = reinterpret_cast<float2 *>(prm_s + 3 * p.n);      // LayerNorm partial sums [8 warps][32 lanes]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full + 8 * b, 1); mbar_init(acc_empty + 8 * b, kEpiWarps); }
    fence_barrier_init();
  }
  for (int i = tid; i < p.n; i += kThreads) {
    prm_s[i] = p.bias ? __ldg(p.bias + i) : 0.0f;
    if (p.flags & 8) {
      prm_s[p.n + i] = __ldg(p.ln_gamma + i);
      prm_s[2 * p.n + i] = __ldg(p.ln_beta + i);
    }
  }
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(&misc[0]), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // tiles of this CTA

  if (warp == kEpiWarps) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (uint32_t)w_bytes);
      for (int cb = 0; cb < p.ncb; ++cb)
        bulk_g2s(w_base + cb * p.n * 128, p.w_img + (int64_t)cb * p.n * kBlockK, (uint32_t)(p.n * 128), w_full);
      uint32_t q = 0;
      for (int it = 0; it < n_my; ++it) {
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * kTileM;
        for (int cb = 0; cb < p.ncb; ++cb, ++q) {
          const uint32_t slot = q % (uint32_t)p.slots, ph = ((q / (uint32_t)p.slots) & 1u) ^ 1u;
          mbar_wait(a_empty + 8 * slot, ph);
          mbar_arrive_expect_tx(a_full + 8 * slot, (uint32_t)kSlotBytes);
          tma_load_2d(a_base + slot * kSlotBytes, &p.tmap_x, cb * kBlockK, row0, a_full + 8 * slot);
        }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ================================ MMA issuer ================================
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_kmajor_sw128_desc(a_base), w_lo0 = (uint32_t)make_kmajor_sw128_desc(w_base);
    const uint32_t w_step = (uint32_t)(p.n * 128) >> 4, part_lo = (uint32_t)(p.n_per_part * 128) >> 4;
    const bool two_parts = p.n_parts == 2;
    const int last_steps = (p.k16 - (p.ncb - 1) * kBlockK) >> 4;
    mbar_wait(w_full, 0);
    uint32_t q = 0;
    for (int it = 0; it < n_my; ++it) {
      const uint32_t buf = (uint32_t)it & 1u;
      mbar_wait(acc_empty + 8 * buf, (((uint32_t)it >> 1) & 1u) ^ 1u);          // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d = tmem_base + buf * 256u;
      for (int cb = 0; cb < p.ncb; ++cb, ++q) {
        const uint32_t slot = q % (uint32_t)p.slots, ph = (q / (uint32_t)p.slots) & 1u;
        const int steps = cb + 1 < p.ncb ? kBlockK / 16 : last_steps;
        mbar_wait(a_full + 8 * slot, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + slot * (kSlotBytes >> 4), b_lo = w_lo0 + (uint32_t)cb * w_step;
#pragma unroll
          for (int ks = 0; ks < kBlockK / 16; ++ks) {
            if (ks < steps) {
              const uint32_t acc = (cb > 0 || ks > 0) ? 1u : 0u;
              umma_bf16_lo(d, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, p.idesc, acc);
              if (two_parts) umma_bf16_lo(d + (uint32_t)p.n_per_part, a_lo + 2 * ks, b_lo + part_lo + 2 * ks, desc_hi, p.idesc, acc);
            }
          }
          umma_commit(a_empty + 8 * slot);
          if (cb + 1 == p.ncb) umma_commit(acc_full + 8 * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quarter = warp & 3, half = warp >> 2;      // TMEM lanes [32 q, 32 q + 32): warps q and q + 4 take alternate chunks
    const bool do_relu = (p.flags & 1) != 0, do_gelu = (p.flags & 4) != 0, do_ln = (p.flags & 8) != 0;
    const bool do_tab = (p.flags & 16) != 0;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    auto add_bf16x16 = [&](const uint4 &ra, const uint4 &rb, float (&y)[16]) {
      const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        y[2 * i] += __uint_as_float(rw[i] << 16);
        y[2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
      }
    };
    auto store16 = [&](__nv_bfloat16 *dst, const float (&y)[16]) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
        o[i] = *reinterpret_cast<const uint32_t *>(&h);
      }
      reinterpret_cast<uint4 *>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<uint4 *>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
    };
    for (int it = 0; it < n_my; ++it) {
      const uint32_t buf = (uint32_t)it & 1u;
      const int64_t row = (int64_t)((int)blockIdx.x + it * (int)gridDim.x) * kTileM + quarter * 32 + lane;
      const bool row_ok = row < p.m;
      const uint32_t t_row = tmem_base + buf * 256u + ((uint32_t)(quarter * 32) << 16);
      // acc + bias for columns [col, col + 16) of this thread's row
      auto load16 = [&](int col, float (&y)[16]) {
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)col, v);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 b = *reinterpret_cast<const float4 *>(prm_s + col + 4 * q4);
          y[4 * q4 + 0] = __uint_as_float(v[4 * q4 + 0]) + b.x;
          y[4 * q4 + 1] = __uint_as_float(v[4 * q4 + 1]) + b.y;
          y[4 * q4 + 2] = __uint_as_float(v[4 * q4 + 2]) + b.z;
          y[4 * q4 + 3] = __uint_as_float(v[4 * q4 + 3]) + b.w;
        }
      };
      // row-wise additive inputs are fetched before the accumulator is ready and one chunk ahead afterwards
      const __nv_bfloat16 *rrow = (p.residual && row_ok) ? p.residual + row * p.n : nullptr;
      const __nv_bfloat16 *trow = (do_tab && row_ok) ? p.table + (int64_t)__ldg(p.tab_idx + row) * p.tab_cols : nullptr;
      uint4 nx[4] = {zero4, zero4, zero4, zero4};
      auto fetch = [&](int col) {
        if (trow && col < p.tab_cols) {
          nx[0] = __ldg(reinterpret_cast<const uint4 *>(trow + col));
          nx[1] = __ldg(reinterpret_cast<const uint4 *>(trow + col) + 1);
        } else {
          nx[0] = nx[1] = zero4;
        }
        if (rrow) {
          nx[2] = __ldg(reinterpret_cast<const uint4 *>(rrow + col));
          nx[3] = __ldg(reinterpret_cast<const uint4 *>(rrow + col) + 1);
        }
      };
      if (half * 16 < p.n) fetch(half * 16);
      mbar_wait(acc_full + 8 * buf, ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      float mean = 0.0f, rstd = 1.0f;
      if (do_ln) {
        // LayerNorm over the n columns: the two warps of a lane quarter own alternate 16-column chunks and exchange their
        // partial (sum, sum of squares) through shared memory at a 64-thread named barrier.
        float sum = 0.0f, sq = 0.0f;
        for (int col = half * 16; col < p.n; col += 32) {
          float y[16];
          load16(col, y);
#pragma unroll
          for (int i = 0; i < 16; ++i) { sum += y[i]; sq = fmaf(y[i], y[i], sq); }
        }
        part[warp * 32 + lane] = make_float2(sum, sq);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        const float2 other = part[(warp ^ 4) * 32 + lane];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // partner has read before the next tile overwrites
        sum += other.x;
        sq += other.y;
        mean = sum / (float)p.n;
        rstd = rsqrtf(fmaxf(sq / (float)p.n - mean * mean, 0.0f) + p.ln_eps);
      }
      for (int col = half * 16; col < p.n; col += 32) {
        const uint4 c0 = nx[0], c1 = nx[1], c2 = nx[2], c3 = nx[3];
        if (col + 32 < p.n) fetch(col + 32);
        float y[16];
        load16(col, y);
        if (row_ok) {
          add_bf16x16(c0, c1, y);                      // table term (zero when absent) joins the pre-activation
          if (do_ln) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 g = *reinterpret_cast<const float4 *>(prm_s + p.n + col + 4 * q4);
              const float4 be = *reinterpret_cast<const float4 *>(prm_s + 2 * p.n + col + 4 * q4);
              y[4 * q4 + 0] = fmaf((y[4 * q4 + 0] - mean) * rstd, g.x, be.x);
              y[4 * q4 + 1] = fmaf((y[4 * q4 + 1] - mean) * rstd, g.y, be.y);
              y[4 * q4 + 2] = fmaf((y[4 * q4 + 2] - mean) * rstd, g.z, be.z);
              y[4 * q4 + 3] = fmaf((y[4 * q4 + 3] - mean) * rstd, g.w, be.w);
            }
          }
          if (rrow) add_bf16x16(c2, c3, y);
          if (do_relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.0f);
          }
          if (do_gelu) {               // erf-form GELU, nn.GELU() default (point_transformer_layer.py:266)
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = gelu_erf_fast(y[i]);
          }
          store16(p.out + row * p.ldo + col, y);
        }
      }
      // this warp is done reading the accumulator: hand the buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * buf);
    }
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

}  // namespace lin
}  // namespace os3d

using namespace os3d;

// Returns 1 when the persistent kernel can take the problem (weights fit in shared memory next to >= 3 activation slots).
extern "C" int os3d_linear_tc_fits(int k, int n) {
  if (k <= 0 || k % 8 || n < 16 || n % 16 || n > 256) return 0;
  const int ncb = (int)cdiv(k, lin::kBlockK);
  const int tail = (2 * lin::kMaxSlots + 5) * 8 + 8 + 3 * n * 4 + 8 * 32 * 8 + 64;
  return 227 * 1024 - 1024 - tail - ncb * n * 128 >= 3 * lin::kSlotBytes;
}

extern "C" int os3d_linear_tc_bf16(const void *x, int64_t m, int k, int n, const void *w, const float *bias, int flags,
                                   const void *residual, const float *ln_gamma, const float *ln_beta, float ln_eps,
                                   const void *table, const int32_t *tab_idx, int tab_cols, void *out, int64_t ldo,
                                   void *stream) {
  if (!os3d_linear_tc_fits(k, n) || m < 0 || ((uintptr_t)x & 15) || (flags & ~(1 | 4 | 8 | 16)) ||
      ((flags & 8) && (!ln_gamma || !ln_beta || (flags & 16))) ||
      ((flags & 16) && (!table || !tab_idx || tab_cols % 16 || tab_cols > n)) || ldo < n || ldo % 8 ||
      ((uintptr_t)out & 15))
    return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  lin::encode_tiled_fn enc = lin::encode_tiled();
  if (!enc) return OS3D_ERR_BAD_ARG;
  lin::Params p;
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)m};
    const cuuint64_t gstr[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)lin::kBlockK, (cuuint32_t)lin::kTileM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&p.tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OS3D_ERR_BAD_ARG;
  }
  p.w_img = (const __nv_bfloat16 *)w;
  p.bias = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.ln_gamma = ln_gamma;
  p.ln_beta = ln_beta;
  p.ln_eps = ln_eps;
  p.table = (const __nv_bfloat16 *)table;
  p.tab_idx = tab_idx;
  p.tab_cols = tab_cols;
  p.out = (__nv_bfloat16 *)out;
  p.m = m;
  p.ldo = ldo;
  p.n_tiles = (int)cdiv(m, lin::kTileM);
  p.k16 = (k + 15) / 16 * 16;
  p.n = n;
  p.ncb = (int)cdiv(k, lin::kBlockK);
  p.flags = flags;
  p.n_parts = 1;
  p.n_per_part = n;
  p.idesc = ptx::make_idesc_bf16(lin::kTileM, n);
  const int tail = (2 * lin::kMaxSlots + 5) * 8 + 8 + 3 * n * 4 + 8 * 32 * 8 + 64;
  int slots = (227 * 1024 - 1024 - tail - p.ncb * n * 128) / lin::kSlotBytes;
  slots = slots > lin::kMaxSlots ? lin::kMaxSlots : slots;
  p.slots = slots;
  const int smem = 1024 + p.ncb * n * 128 + slots * lin::kSlotBytes + tail;
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: once per device, not once per process
  static bool configured[64] = {false};
  int cfg_dev = 0;
  OS3D_CUDA(cudaGetDevice(&cfg_dev));
  if (cfg_dev < 0 || cfg_dev >= 64 || !configured[cfg_dev]) {
    OS3D_CUDA(cudaFuncSetAttribute(lin::linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (cfg_dev >= 0 && cfg_dev < 64) configured[cfg_dev] = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  lin::linear_tc_kernel<<<grid, lin::kThreads, smem, (cudaStream_t)stream>>>(p);
  OS3D_LAUNCH_CHECK();
  return 0;
}
