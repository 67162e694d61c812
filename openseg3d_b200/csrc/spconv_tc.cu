// spconv_tc.cu -- stage 3, bf16 mode: output-stationary sparse convolution on the tcgen05 tensor cores (sm_100a).
//
//   out[r, :] = epilogue( sum_k  in[nbr[r, k], :] . W[k] )        bf16 in / bf16 out, fp32 accumulate in TMEM
//
// One CTA owns 128 output rows (UMMA M = 128, cta_group::1) and ALL output channels: the accumulator is a
// [128 lanes x Cout columns] fp32 tile in tensor memory, so a gathered input row is fetched once per CTA and nothing
// is ever scattered or atomically added.  The contraction dimension is the flattened (offset k, input channel) axis,
// 27*Cin long, cut into 64-element blocks (one 128-byte SWIZZLE_128B row per output row / output channel):
//
//   warps 0-7  producers : gather.  Thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 32 j; each chunk is one
//              cp.async (LDGSTS, zero-fill when the row has no neighbour at that offset) straight into the swizzled
//              K-major UMMA layout; 8 lanes cover one 128-byte row, so every global request is a full line.
//              Completion is signalled by cp.async.mbarrier.arrive.noinc on the stage's full barrier, so producers never
//              wait on their own loads (a per-stage fence.proxy.async in the producers costs a MEMBAR that drains the
//              in-flight LDGSTS: measured 1.5 us per K-block); the proxy fence is executed by the MMA thread instead.
//              Weights: the packed image is stored in global memory already swizzled, one contiguous [Cout x 128 B]
//              slab per K-block, so ONE cp.async.bulk (TMA bulk copy, mbarrier complete_tx) by one thread fills B.
//   warp 8     MMA issuer : one thread issues tcgen05.mma.kind::f16 (4 K-steps x N-parts per block), commits the
//              stage back to the producers (tcgen05.commit -> empty barrier) and finally signals the epilogue.
//   warps 0-7  epilogue  : tcgen05.ld 32x32b (lane = output row; warps w and w+4 split the columns), y = acc*scale + shift (+ residual) (ReLU), bf16,
//              32-byte vector stores.  Bias, eval-mode BatchNorm, the residual add and the ReLU of the reference's
//              SparseBasicBlock / ConvModule therefore never touch HBM as separate passes.
//   K-blocks whose offsets have no neighbour anywhere in the tile are skipped by all three roles (same enumeration).
//
// replaces: SubMConv3d / SparseConv3d / SparseInverseConv3d forward of spconv-cu113 (+ the BatchNorm1d / ReLU / add
// that follow them in seg3d/utils/spconv_utils.py:26-30 and seg3d/models/backbones/pointtransformer.py:47-66).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace tc {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                        // bf16 elements per K-block = one 128-byte swizzle row
constexpr int kATileBytes = kTileM * 128;          // 16 KB
constexpr int kProducerWarps = 8;                  // gather producers, then epilogue (two warps per TMEM lane quarter)
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = kProducerThreads + 32;    // + the MMA-issuing warp
constexpr int kMaxStages = 6;

struct Params {
  const __nv_bfloat16 *in;
  const int32_t *nbr;
  int64_t m_out;
  int cin;              // feature row pitch in elements (multiple of 8)
  int cout;             // multiple of 16, <= 512
  const __nv_bfloat16 *w_img;
  const int32_t *chunk_tab;   // [n_blocks * 8]: (offset k << 16) | first channel of the 16-byte chunk, -1 beyond K
  const uint32_t *blk_mask;   // [n_blocks]: bit mask of the kernel offsets a K-block touches
  int n_blocks;         // ceil(27 * cin / 64)
  int chunks_per_offset;  // cin / 8
  int total_chunks;     // 27 * cin / 8
  const float *scale, *shift;
  const __nv_bfloat16 *residual;
  int relu;
  __nv_bfloat16 *out;
  int tmem_cols;        // power of two >= cout, >= 32
  int n_parts;          // 1, or 2 when cout > 256
  int n_per_part;       // cout / n_parts (multiple of 16)
  uint32_t idesc;       // tcgen05 instruction descriptor (bf16 x bf16 -> f32, M=128, N=n_per_part, K-major A and B)
  int stages;
};

__global__ void __launch_bounds__(kThreads) spconv_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // carve: 1024-aligned A stages | B stages | neighbour tile | barriers
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const int b_tile_bytes = p.cout * 128;
  const uint32_t a_base = base;
  const uint32_t b_base = base + p.stages * kATileBytes;
  uint8_t *tail = smem + p.stages * (kATileBytes + b_tile_bytes);
  int32_t *nbr_s = reinterpret_cast<int32_t *>(tail);                          // [128][27]
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail + kTileM * OS3D_KVOL * 4);  // full[S], empty[S], accum
  uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 1);     // [0] tmem base, [1] offset mask
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages), accum_bar = smem_u32(bars + 2 * kMaxStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTileM;
  const int rows = (int)min((int64_t)kTileM, p.m_out - row0);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, kProducerThreads + 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    misc[1] = 0;
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == kProducerWarps) tmem_alloc(smem_u32(&misc[0]), (uint32_t)p.tmem_cols);
  {
    uint32_t mask = 0;
    for (int t = tid; t < kTileM * OS3D_KVOL; t += kThreads) {
      const int r = t / OS3D_KVOL;
      const int32_t v = r < rows ? __ldg(p.nbr + row0 * OS3D_KVOL + t) : -1;
      nbr_s[t] = v;
      if (v >= 0) mask |= 1u << (t - r * OS3D_KVOL);
    }
    mask = __reduce_or_sync(0xffffffffu, mask);
    if (lane == 0 && mask) atomicOr(&misc[1], mask);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  const uint32_t has_k = misc[1];

  if (warp < kProducerWarps) {
    // ================================ producers ================================
    // Per-thread invariants: chunk c of rows r0 + 16 j.  Everything that does not depend on the K-block is hoisted;
    // the (offset, channel) of a chunk comes from a per-layer table written next to the weight image, so the loop
    // body is: table load, then per row { LDS neighbour, multiply-add address, LDGSTS }.
    constexpr int kRowStep = kProducerThreads / 8, kRowsPerThread = kTileM / kRowStep;
    const int c = tid & 7, r0 = tid >> 3;
    const uint32_t dst0 = (uint32_t)(r0 * 128 + ((c ^ (r0 & 7)) << 4));      // row r0 + kRowStep j -> + j * kRowStep * 128
    const int32_t *nrow = nbr_s + r0 * OS3D_KVOL;                           // row r0 + kRowStep j -> + j * kRowStep * 27
    const char *in_bytes = reinterpret_cast<const char *>(p.in);
    const uint32_t row_bytes = (uint32_t)p.cin * 2u;
    const int32_t *tab = p.chunk_tab + c;
    int it = 0, stage = 0;
    uint32_t phase = 1;                                                     // empty barriers start "free"
    for (int blk = 0; blk < p.n_blocks; ++blk) {
      if (!(__ldg(p.blk_mask + blk) & has_k)) continue;
      mbar_wait(empty0 + 8 * stage, phase);
      if (tid == 0) {
        mbar_arrive_expect_tx(full0 + 8 * stage, (uint32_t)b_tile_bytes);
        bulk_g2s(b_base + stage * b_tile_bytes, p.w_img + (int64_t)blk * p.cout * kBlockK, (uint32_t)b_tile_bytes,
                 full0 + 8 * stage);
      }
      const int32_t t = __ldg(tab + blk * 8);
      const int koff = t >> 16;                        // -1 when the chunk lies beyond 27 * cin (t == -1)
      const uint32_t ch_bytes = (uint32_t)(t & 0xffff) * 2u;
      const uint32_t a_dst = a_base + stage * kATileBytes + dst0;
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        const int32_t n = t >= 0 ? nrow[j * kRowStep * OS3D_KVOL + koff] : -1;
        const char *src = in_bytes + ((uint64_t)((uint32_t)max(n, 0) * row_bytes) + ch_bytes);
        cp_async_16(a_dst + j * kRowStep * 128, src, n >= 0 ? 16u : 0u);
      }
      // asynchronous arrive: fires on the full barrier once this thread's copies above have landed -- the producer
      // never waits on its own loads and runs ahead until the ring is full
      cp_async_mbar_arrive_noinc(full0 + 8 * stage);
      ++it;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }

    // ================================ epilogue ================================
    if (it > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const int quarter = warp & 3;                       // TMEM lanes [32 q, 32 q + 32) are visible to warps q and q + 4
    const int row = quarter * 32 + lane;
    const bool row_ok = row < rows;
    __nv_bfloat16 *orow = p.out + (row0 + row) * p.cout;
    const bool pair_sum = (p.relu & 2) != 0;   // residual rows hold 2*cout channels; add r[2c] + r[2c+1] after the ReLU
    const bool do_relu = (p.relu & 1) != 0;
    const __nv_bfloat16 *rrow = p.residual ? p.residual + (row0 + row) * p.cout * (pair_sum ? 2 : 1) : nullptr;
    for (int col = (warp >> 2) * 16; col < p.cout; col += 16 * (kProducerWarps / 4)) {
      uint32_t v[16];
      if (it > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      if (row_ok) {
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(v[i]);
        if (p.scale) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.scale + col) + q);
            const float4 sh = __ldg(reinterpret_cast<const float4 *>(p.shift + col) + q);
            y[4 * q + 0] = fmaf(y[4 * q + 0], sc.x, sh.x);
            y[4 * q + 1] = fmaf(y[4 * q + 1], sc.y, sh.y);
            y[4 * q + 2] = fmaf(y[4 * q + 2], sc.z, sh.z);
            y[4 * q + 3] = fmaf(y[4 * q + 3], sc.w, sh.w);
          }
        }
        if (rrow && !pair_sum) {
          const uint4 ra = __ldg(reinterpret_cast<const uint4 *>(rrow + col));
          const uint4 rb = __ldg(reinterpret_cast<const uint4 *>(rrow + col) + 1);
          const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            y[2 * i] += __uint_as_float(rw[i] << 16);
            y[2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
          }
        }
        if (do_relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.0f);
        }
        if (rrow && pair_sum) {      // UpBlock: x_m + channel_reduction(cat)  (pointtransformer.py:89-110)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(rrow + 2 * col) + q);
            const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
              y[4 * q + i] += __uint_as_float(rw[i] << 16) + __uint_as_float(rw[i] & 0xffff0000u);
          }
        }
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
          o[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        reinterpret_cast<uint4 *>(orow + col)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<uint4 *>(orow + col)[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    tc_fence_before();
  } else {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int it = 0, stage = 0;
      uint32_t phase = 0;
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        if (!(__ldg(p.blk_mask + blk) & has_k)) continue;
        mbar_wait(full0 + 8 * stage, phase);
        fence_proxy_async();   // cp.async (generic-proxy) writes observed through the barrier -> visible to the MMA's async proxy
        tc_fence_after();
        const uint32_t a_stage = a_base + stage * kATileBytes;
        const uint32_t b_stage = b_base + stage * b_tile_bytes;
#pragma unroll
        for (int ks = 0; ks < kBlockK / 16; ++ks) {
          const uint64_t adesc = make_kmajor_sw128_desc(a_stage + ks * 32);
          for (int part = 0; part < p.n_parts; ++part) {
            const uint64_t bdesc = make_kmajor_sw128_desc(b_stage + part * p.n_per_part * 128 + ks * 32);
            umma_bf16(tmem_base + (uint32_t)(part * p.n_per_part), adesc, bdesc, p.idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8 * stage);  // frees the stage once the MMAs above have read it
        ++it;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (it > 0) umma_commit(accum_bar);  // accumulator complete -> epilogue
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// spconv 2.x weight [cout, 27, cin] f32  ->  bf16 UMMA image [n_blocks][cout][8 chunks, XOR-swizzled by row & 7][8]:
// exactly the bytes a K-block's B tile occupies in shared memory, so one bulk copy loads it.  Two small per-layer
// tables follow the image: chunk_tab[n_blocks * 8] and blk_mask[n_blocks] (see Params).
__global__ void pack_weight_img_kernel(const float *__restrict__ src, int cin, int cout, int cin_pad, int n_blocks,
                                       __nv_bfloat16 *__restrict__ dst, int32_t *__restrict__ chunk_tab,
                                       uint32_t *__restrict__ blk_mask) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n_blocks * cout * kBlockK;
  const int cpo = cin_pad / 8;
  const int total_chunks = OS3D_KVOL * cpo;
  if (t < n_blocks * 8) {
    const int g = (int)t;
    chunk_tab[g] = g < total_chunks ? (((g / cpo) << 16) | ((g % cpo) * 8)) : -1;
  }
  if (t < n_blocks) {
    const int first = ((int)t * 8) / cpo, last = min((int)t * 8 + 7, total_chunks - 1) / cpo;
    blk_mask[t] = (uint32_t)(((1ull << (last + 1)) - 1) & ~((1ull << first) - 1));
  }
  if (t >= total) return;
  const int e = (int)(t & 7);
  const int pc = (int)((t >> 3) & 7);
  const int n = (int)((t >> 6) % cout);
  const int blk = (int)((t >> 6) / cout);
  const int c = pc ^ (n & 7);  // logical chunk stored at physical chunk pc
  const int g = blk * 8 + c;
  float v = 0.0f;
  if (g < total_chunks) {
    const int koff = g / cpo, ch = (g - koff * cpo) * 8 + e;
    if (ch < cin) v = src[((int64_t)n * OS3D_KVOL + koff) * cin + ch];
  }
  dst[t] = __float2bfloat16(v);
}

}  // namespace tc
}  // namespace os3d

using namespace os3d;

extern "C" int os3d_spconv_bf16_packed_elems(int cin_pad, int cout, int64_t *elems) {
  if (cin_pad <= 0 || cin_pad % 8 || cout <= 0) return OS3D_ERR_BAD_ARG;
  const int64_t n_blocks = cdiv((int64_t)OS3D_KVOL * cin_pad, tc::kBlockK);
  *elems = n_blocks * cout * tc::kBlockK + (n_blocks * 9 * 4 + 64) / 2;   // image + chunk_tab + blk_mask (bf16 units)
  return 0;
}

extern "C" int os3d_pack_weight_bf16(const float *w_spconv, int cin, int cout, int cin_pad, void *w_packed,
                                     void *stream) {
  if (cin_pad % 8 || cin_pad < cin) return OS3D_ERR_BAD_ARG;
  const int n_blocks = (int)cdiv((int64_t)OS3D_KVOL * cin_pad, tc::kBlockK);
  const int64_t total = (int64_t)n_blocks * cout * tc::kBlockK;
  __nv_bfloat16 *img = (__nv_bfloat16 *)w_packed;
  int32_t *tab = reinterpret_cast<int32_t *>(img + total);               // total * 2 bytes is a multiple of 128
  tc::pack_weight_img_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w_spconv, cin, cout, cin_pad, n_blocks, img, tab, reinterpret_cast<uint32_t *>(tab + n_blocks * 8));
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_spconv_fwd_bf16(const void *in, const int32_t *nbr, int64_t m_out, int cin, int cout, const void *w,
                                    const float *scale, const float *shift, const void *residual, int relu, void *out,
                                    void *stream) {
  if (cin <= 0 || cin % 8 || cout < 16 || cout % 16 || cout > 512 || (cout > 256 && cout % 32) ||
      ((scale == nullptr) != (shift == nullptr)))
    return OS3D_ERR_BAD_ARG;
  if (m_out == 0) return 0;
  tc::Params p;
  p.in = (const __nv_bfloat16 *)in;
  p.nbr = nbr;
  p.m_out = m_out;
  p.cin = cin;
  p.cout = cout;
  p.w_img = (const __nv_bfloat16 *)w;
  p.n_blocks = (int)cdiv((int64_t)OS3D_KVOL * cin, tc::kBlockK);
  p.chunk_tab = reinterpret_cast<const int32_t *>(p.w_img + (int64_t)p.n_blocks * cout * tc::kBlockK);
  p.blk_mask = reinterpret_cast<const uint32_t *>(p.chunk_tab + p.n_blocks * 8);
  p.chunks_per_offset = cin / 8;
  p.total_chunks = OS3D_KVOL * cin / 8;
  p.scale = scale;
  p.shift = shift;
  p.residual = (const __nv_bfloat16 *)residual;
  p.relu = relu;
  p.out = (__nv_bfloat16 *)out;
  int cols = 32;
  while (cols < cout) cols <<= 1;
  p.tmem_cols = cols;
  p.n_parts = cout > 256 ? 2 : 1;
  p.n_per_part = cout / p.n_parts;
  // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major A and B, N>>3 at bit 17,
  // M>>4 at bit 24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_per_part >> 3) << 17) | ((uint32_t)(tc::kTileM >> 4) << 24);
  const int stage_bytes = tc::kATileBytes + cout * 128;
  const int tail = tc::kTileM * OS3D_KVOL * 4 + (2 * tc::kMaxStages + 1) * 8 + 64;
  // small tiles: 4 stages so two CTAs share an SM; large tiles: as many stages as fit one CTA per SM
  int stages = cout <= 48 ? 4 : cout <= 96 ? 3 : (227 * 1024 - 1024 - tail) / stage_bytes;
  stages = stages > tc::kMaxStages ? tc::kMaxStages : stages;
  if (stages < 2) return OS3D_ERR_BAD_ARG;
  p.stages = stages;
  const int smem = 1024 + stages * stage_bytes + tail;
  static int configured = 0;
  if (configured < smem) {
    OS3D_CUDA(cudaFuncSetAttribute(tc::spconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  tc::spconv_tc_kernel<<<(unsigned)cdiv(m_out, tc::kTileM), tc::kThreads, smem, (cudaStream_t)stream>>>(p);
  OS3D_LAUNCH_CHECK();
  return 0;
}
