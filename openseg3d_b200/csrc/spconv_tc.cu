// spconv_tc.cu -- stage 3, bf16 mode: output-stationary sparse convolution on the tcgen05 tensor cores (sm_100a).
//
//   out[r, :] = epilogue( sum_k  in[nbr[r, k], :] . W[k] )        bf16 in / bf16 out, fp32 accumulate in TMEM
//
// One CTA owns 128 output rows (UMMA M = 128, cta_group::1) and ALL output channels: the accumulator is a
// [128 lanes x Cout columns] fp32 tile in tensor memory, so a gathered input row is fetched once per CTA and nothing
// is ever scattered or atomically added.  The contraction dimension is the flattened (offset k, input channel) axis,
// 27*Cin long, cut into 64-element blocks (one 128-byte SWIZZLE_128B row per output row / output channel):
//
//   warps 0-3  producers : gather.  Thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 16 j; each chunk is one
//              cp.async (LDGSTS, zero-fill when the row has no neighbour at that offset) straight into the swizzled
//              K-major UMMA layout; 8 lanes cover one 128-byte row, so every global request is a full line.
//              Completion is tracked per stage with cp.async groups -> fence.proxy.async -> mbarrier arrive, two stages
//              behind the issue point so loads stay in flight.
//              Weights: the packed image is stored in global memory already swizzled, one contiguous [Cout x 128 B]
//              slab per K-block, so ONE cp.async.bulk (TMA bulk copy, mbarrier complete_tx) by one thread fills B.
//   warp 4     MMA issuer : one thread issues tcgen05.mma.kind::f16 (4 K-steps x N-parts per block), commits the
//              stage back to the producers (tcgen05.commit -> empty barrier) and finally signals the epilogue.
//   warps 0-3  epilogue  : tcgen05.ld 32x32b (lane = output row), y = acc*scale + shift (+ residual) (ReLU), bf16,
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
constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;
constexpr int kLag = 2;                            // producer arrives this many stages behind its cp.async issue
constexpr int kMaxStages = 6;

struct Params {
  const __nv_bfloat16 *in;
  const int32_t *nbr;
  int64_t m_out;
  int cin;              // feature row pitch in elements (multiple of 8)
  int cout;             // multiple of 16, <= 512
  const __nv_bfloat16 *w_img;
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

// offsets (bit mask over the 27 kernel offsets) that K-block `blk` touches
__device__ __forceinline__ uint32_t block_offsets(int blk, int cpo, int total_chunks) {
  const int first = (blk * 8) / cpo;
  const int last = min(blk * 8 + 7, total_chunks - 1) / cpo;
  return (uint32_t)(((1ull << (last + 1)) - 1) & ~((1ull << first) - 1));
}

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
  if (warp == 4) tmem_alloc(smem_u32(&misc[0]), (uint32_t)p.tmem_cols);
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
  const int cpo = p.chunks_per_offset;

  if (warp < 4) {
    // ================================ producers ================================
    const int c = tid & 7, r0 = tid >> 3;
    int it = 0;
    for (int blk = 0; blk < p.n_blocks; ++blk) {
      if (!(block_offsets(blk, cpo, p.total_chunks) & has_k)) continue;
      const int stage = it % p.stages;
      mbar_wait(empty0 + 8 * stage, ((it / p.stages) & 1) ^ 1);
      if (tid == 0) {
        mbar_arrive_expect_tx(full0 + 8 * stage, (uint32_t)b_tile_bytes);
        bulk_g2s(b_base + stage * b_tile_bytes, p.w_img + (int64_t)blk * p.cout * kBlockK, (uint32_t)b_tile_bytes,
                 full0 + 8 * stage);
      }
      const int g = blk * 8 + c;
      const bool chunk_ok = g < p.total_chunks;
      const int koff = chunk_ok ? g / cpo : 0;
      const int cc = g - koff * cpo;
      const uint32_t a_stage = a_base + stage * kATileBytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = r0 + 16 * j;
        const int32_t n = chunk_ok ? nbr_s[row * OS3D_KVOL + koff] : -1;
        const __nv_bfloat16 *src = p.in + (n >= 0 ? (int64_t)n * p.cin + cc * 8 : 0);
        cp_async_16(a_stage + row * 128 + ((c ^ (row & 7)) << 4), src, n >= 0 ? 16u : 0u);
      }
      cp_async_commit();
      if (it >= kLag) {
        cp_async_wait<kLag>();
        fence_proxy_async();
        mbar_arrive(full0 + 8 * ((it - kLag) % p.stages));
      }
      ++it;
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int d = max(0, it - kLag); d < it; ++d) mbar_arrive(full0 + 8 * (d % p.stages));

    // ================================ epilogue ================================
    if (it > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const int row = warp * 32 + lane;
    const bool row_ok = row < rows;
    __nv_bfloat16 *orow = p.out + (row0 + row) * p.cout;
    const __nv_bfloat16 *rrow = p.residual ? p.residual + (row0 + row) * p.cout : nullptr;
    for (int col = 0; col < p.cout; col += 16) {
      uint32_t v[16];
      if (it > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      if (row_ok) {
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          y[i] = __uint_as_float(v[i]);
          if (p.scale) y[i] = fmaf(y[i], __ldg(p.scale + col + i), __ldg(p.shift + col + i));
        }
        if (rrow) {
          const uint4 ra = __ldg(reinterpret_cast<const uint4 *>(rrow + col));
          const uint4 rb = __ldg(reinterpret_cast<const uint4 *>(rrow + col) + 1);
          const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            y[2 * i] += __uint_as_float(rw[i] << 16);
            y[2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
          }
        }
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = y[2 * i], b = y[2 * i + 1];
          if (p.relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
          const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
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
      int it = 0;
      for (int blk = 0; blk < p.n_blocks; ++blk) {
        if (!(block_offsets(blk, cpo, p.total_chunks) & has_k)) continue;
        const int stage = it % p.stages;
        mbar_wait(full0 + 8 * stage, (it / p.stages) & 1);
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
      }
      if (it > 0) umma_commit(accum_bar);  // accumulator complete -> epilogue
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// spconv 2.x weight [cout, 27, cin] f32  ->  bf16 UMMA image [n_blocks][cout][8 chunks, XOR-swizzled by row & 7][8]:
// exactly the bytes a K-block's B tile occupies in shared memory, so one bulk copy loads it.
__global__ void pack_weight_img_kernel(const float *__restrict__ src, int cin, int cout, int cin_pad, int n_blocks,
                                       __nv_bfloat16 *__restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)n_blocks * cout * kBlockK;
  if (t >= total) return;
  const int e = (int)(t & 7);
  const int pc = (int)((t >> 3) & 7);
  const int n = (int)((t >> 6) % cout);
  const int blk = (int)((t >> 6) / cout);
  const int c = pc ^ (n & 7);  // logical chunk stored at physical chunk pc
  const int cpo = cin_pad / 8;
  const int g = blk * 8 + c;
  float v = 0.0f;
  if (g < OS3D_KVOL * cpo) {
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
  *elems = cdiv((int64_t)OS3D_KVOL * cin_pad, tc::kBlockK) * cout * tc::kBlockK;
  return 0;
}

extern "C" int os3d_pack_weight_bf16(const float *w_spconv, int cin, int cout, int cin_pad, void *w_packed,
                                     void *stream) {
  if (cin_pad % 8 || cin_pad < cin) return OS3D_ERR_BAD_ARG;
  const int n_blocks = (int)cdiv((int64_t)OS3D_KVOL * cin_pad, tc::kBlockK);
  const int64_t total = (int64_t)n_blocks * cout * tc::kBlockK;
  tc::pack_weight_img_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w_spconv, cin, cout, cin_pad, n_blocks, (__nv_bfloat16 *)w_packed);
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
  // small tiles: 3 stages so two CTAs share an SM; large tiles: as many stages as fit one CTA per SM
  int stages = cout <= 96 ? 3 : (227 * 1024 - 1024 - tail) / stage_bytes;
  stages = stages > tc::kMaxStages ? tc::kMaxStages : stages;
  if (stages < tc::kLag + 1) return OS3D_ERR_BAD_ARG;
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
