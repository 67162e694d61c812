// qkv_tc.cu -- wide Linear layers of the SWFormer encoder on tcgen05, weights streamed, outputs staged and stored
// coalesced:
//
//   y[r, :] = x[r, :k] . W^T  (+ per-row table term | + bias)  ->  per-head L2 normalisation | GELU | nothing  ->  bf16
//
// Users: (1) the attention in-projection  q | k | v = in_proj(x + pos, x)  of CosineMultiheadAttention
// (seg3d/models/layers/cosine_msa.py:48-63, :152-153): one launch produces the head-padded, already L2-normalised q and k
// and v of a layer, with the position embedding folded in as a table ((x + pos) W^T = x W^T + (pos W^T)[pos_idx]) -- it
// replaces two library GEMMs, the position gather-add pass and the normalisation pass; (2) fc1 + GELU of the level-4 MLP
// (seg3d/models/layers/point_transformer_layer.py:260-276), whose hidden width (768) is beyond the fused MLP kernel.
//
// Shape of the problem: K = C (48..384) is small, N = 3 * heads * dp (384..1152) is wide, so the kernel is bound by its
// OUTPUT bytes.  One persistent CTA per SM walks 128-row tiles:
//   warp 8  producer : the x tile by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B = the UMMA K-major layout), all K-blocks of
//                      a tile resident (double-buffered across tiles when it fits); the weight image in chunks of NC
//                      output columns ([chunk][K-block][NC][128 B swizzled]: one cp.async.bulk per chunk) through a 2-slot
//                      ring -- the weights never fit in shared memory as a whole (level 3: 295 KB), they stream from L2.
//   warp 9  MMA      : per chunk ceil(K / 16) tcgen05.mma (M = 128, N = NC) into one of two TMEM accumulators.
//   warps 0-7 epilogue: thread = row (tcgen05.ld 32x32b); per chunk and head: add the table row / bias, normalise the
//                      head (the dp columns of a head are in ONE thread: no shuffles), or GELU; pack bf16 into a staging
//                      tile in shared memory; then all 256 threads store the tile with fully coalesced 16-byte vectors
//                      (a row-per-lane store touches 32 lines per instruction: that was what kept linear_tc.cu's table-mode
//                      projection behind the library GEMM).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace qkv {
// packed pairs of floats for the two-wide FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2)
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
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kSlotBytes = kTileM * 128;
constexpr int kEpiWarps = 16;                      // 4 per TMEM lane quarter
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kMaxChunks = 32;

struct alignas(64) Params {
  CUtensorMap tmap_x;            // x [m, k] bf16, box {64, 128}, SWIZZLE_128B, zero OOB fill
  CUtensorMap tmap_out;          // out [m, n] (pitch ldo) bf16, box {64, 128}, SWIZZLE_128B (tma_store only)
  const __nv_bfloat16 *w_img;    // [n_chunks][ncb][nc][64] swizzled weight image
  const float *bias;             // [n] or NULL (chunks of mode 0 / 2)
  const __nv_bfloat16 *table;    // [rows, tab_ld] bf16: additive row term of the mode-1 chunks (bias folded in)
  const int32_t *tab_idx;        // [m]
  int64_t tab_ld;
  __nv_bfloat16 *out;            // [m, ldo]
  int64_t m, ldo;
  int n_tiles, k16, ncb, nc, n_chunks, dp, a_bufs;
  int mode_n1;                   // chunks [0, mode_n1): mode 1 (table row, per-head L2 normalisation if `normalize`); the rest: mode_rest
  int normalize;
  int tma_store;                 // nc % 64 == 0: the staged chunk is written by cp.async.bulk.tensor stores
  int mode_rest;                 // 0 = + bias, 2 = + bias, GELU
  uint32_t idesc;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

template <int DP>
__global__ void __launch_bounds__(kThreads, 1) qkv_proj_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const int a_bytes = p.ncb * kSlotBytes;                  // one x tile, all K-blocks
  const int w_bytes = p.ncb * p.nc * 128;                  // one weight chunk
  const int st_pitch = p.nc * 2 + 16;                      // staging row pitch (bytes): + 16 keeps 16-byte row stores conflict free
  const uint32_t a_base = base;
  const uint32_t w_base = base + p.a_bufs * a_bytes;       // a_bytes, w_bytes are multiples of 1024 (nc % 8 == 0)
  uint8_t *stage = smem + p.a_bufs * a_bytes + 2 * w_bytes;   // [2] staging buffers (1024-byte aligned: a_bytes, w_bytes are)
  const int stage_bytes = p.tma_store ? (p.nc / 64) * kSlotBytes : ((kTileM * st_pitch + 1023) & ~1023);
  uint8_t *tail = stage + 2 * stage_bytes;
  tail += (16 - ((uintptr_t)tail & 15)) & 15;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);     // a_full[2] a_empty[2] w_full[2] w_empty[2] acc_full[2] acc_empty[2]
  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + 2), w_full = smem_u32(bars + 4), w_empty = smem_u32(bars + 6);
  const uint32_t acc_full = smem_u32(bars + 8), acc_empty = smem_u32(bars + 10);
  uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 12);   // [0] tmem base
  float *bias_s = reinterpret_cast<float *>(misc + 4);        // [n] bias (16-byte aligned)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full + 8 * b, 1);
      mbar_init(a_empty + 8 * b, 1);
      mbar_init(w_full + 8 * b, 1);
      mbar_init(w_empty + 8 * b, 1);
      mbar_init(acc_full + 8 * b, 1);
      mbar_init(acc_empty + 8 * b, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (p.bias)
    for (int i = tid; i < p.nc * p.n_chunks; i += kThreads) bias_s[i] = __ldg(p.bias + i);
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(&misc[0]), 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // tiles of this CTA

  if (warp == kEpiWarps) {
    // ================================ producer ================================
    if (lane == 0) {
      auto load_a = [&](int it) {
        const int ab = p.a_bufs == 2 ? (it & 1) : 0;
        const int use = p.a_bufs == 2 ? (it >> 1) : it;                  // how often this buffer was filled before
        if (use > 0) mbar_wait(a_empty + 8 * ab, (uint32_t)(use - 1) & 1u);
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * kTileM;
        mbar_arrive_expect_tx(a_full + 8 * ab, (uint32_t)a_bytes);
        for (int cb = 0; cb < p.ncb; ++cb)
          tma_load_2d(a_base + ab * a_bytes + cb * kSlotBytes, &p.tmap_x, cb * kBlockK, row0, a_full + 8 * ab);
      };
      if (n_my > 0) load_a(0);
      uint32_t g = 0;
      for (int it = 0; it < n_my; ++it) {
        for (int j = 0; j < p.n_chunks; ++j, ++g) {
          const uint32_t wb = g & 1u;
          if (g >= 2) mbar_wait(w_empty + 8 * wb, ((g >> 1) - 1u) & 1u);
          mbar_arrive_expect_tx(w_full + 8 * wb, (uint32_t)w_bytes);
          bulk_g2s(w_base + wb * w_bytes, p.w_img + (int64_t)j * p.ncb * p.nc * kBlockK, (uint32_t)w_bytes, w_full + 8 * wb);
          // the next tile's x follows this tile's first weight chunk (two x buffers), or waits for this tile's last MMA
          if (j == 0 && p.a_bufs == 2 && it + 1 < n_my) load_a(it + 1);
        }
        if (p.a_bufs == 1 && it + 1 < n_my) load_a(it + 1);
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + 1) {
    // ================================ MMA issuer ================================
    const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_kmajor_sw128_desc(a_base), w_lo0 = (uint32_t)make_kmajor_sw128_desc(w_base);
    const uint32_t cb_step = (uint32_t)(p.nc * 128) >> 4;
    const int last_steps = (p.k16 - (p.ncb - 1) * kBlockK) >> 4;
    uint32_t g = 0;
    for (int it = 0; it < n_my; ++it) {
      const int ab = p.a_bufs == 2 ? (it & 1) : 0;
      const int use = p.a_bufs == 2 ? (it >> 1) : it;
      mbar_wait(a_full + 8 * ab, (uint32_t)use & 1u);
      for (int j = 0; j < p.n_chunks; ++j, ++g) {
        const uint32_t wb = g & 1u;
        mbar_wait(w_full + 8 * wb, (g >> 1) & 1u);
        if (g >= 2) mbar_wait(acc_empty + 8 * wb, ((g >> 1) - 1u) & 1u);        // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + wb * 128u;
        if (elect_one()) {
          for (int cb = 0; cb < p.ncb; ++cb) {
            const int steps = cb + 1 < p.ncb ? kBlockK / 16 : last_steps;
            const uint32_t a_lo = a_lo0 + (uint32_t)((ab * a_bytes + cb * kSlotBytes) >> 4);
            const uint32_t b_lo = w_lo0 + (uint32_t)((wb * w_bytes) >> 4) + (uint32_t)cb * cb_step;
#pragma unroll
            for (int ks = 0; ks < kBlockK / 16; ++ks)
              if (ks < steps) umma_bf16_lo(d, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, p.idesc, (cb > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(w_empty + 8 * wb);
          umma_commit(acc_full + 8 * wb);
          if (j + 1 == p.n_chunks) umma_commit(a_empty + 8 * ab);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quarter = warp & 3, half = warp >> 2;      // TMEM lanes [32 q, 32 q + 32): warps q, q + 4, q + 8, q + 12 take heads round robin
    const int r_tile = quarter * 32 + lane;
    const int heads_per_chunk = p.nc / DP;
    constexpr int kMaxH = (128 / DP + 3) / 4;            // heads of a chunk per thread (nc <= 128)
    constexpr int kV = DP / 8;                           // 16-byte vectors per head
    const bool tma_out = p.tma_store != 0;               // nc % 64 == 0: staging tiles in the SWIZZLE_128B layout, stored by TMA
    // The additive row terms of this thread's heads (table rows: an L2 round trip each) are requested ONE CHUNK AHEAD
    // (fetched inside the chunk they were the kernel's pace: 15 % of all stall samples on the first use of the value).
    auto row_ptr = [&](int it2) -> const __nv_bfloat16 * {
      const int64_t r2 = (int64_t)((int)blockIdx.x + it2 * (int)gridDim.x) * kTileM + r_tile;
      return (p.table && r2 < p.m) ? p.table + (int64_t)__ldg(p.tab_idx + r2) * p.tab_ld : nullptr;
    };
    uint4 tpre[kMaxH * kV];
    auto load_tab = [&](const __nv_bfloat16 *tr, int j2) {
      const bool on = tr != nullptr && j2 < p.mode_n1;
#pragma unroll
      for (int hi = 0; hi < kMaxH; ++hi) {
        const int hd = half + 4 * hi;
#pragma unroll
        for (int c = 0; c < kV; ++c)
          tpre[hi * kV + c] = (on && hd < heads_per_chunk) ? __ldg(reinterpret_cast<const uint4 *>(tr + j2 * p.nc + hd * DP) + c)
                                                           : make_uint4(0, 0, 0, 0);
      }
    };
    const __nv_bfloat16 *trow = n_my > 0 ? row_ptr(0) : nullptr;
    load_tab(trow, 0);
    uint32_t g = 0;
    for (int it = 0; it < n_my; ++it) {
      const int64_t row0 = (int64_t)((int)blockIdx.x + it * (int)gridDim.x) * kTileM;
      const __nv_bfloat16 *trow_next = it + 1 < n_my ? row_ptr(it + 1) : nullptr;
      for (int j = 0; j < p.n_chunks; ++j, ++g) {
        const uint32_t wb = g & 1u;
        const int mode = j < p.mode_n1 ? 1 : p.mode_rest;
        const uint32_t t_row = tmem_base + wb * 128u + ((uint32_t)(quarter * 32) << 16);
        const bool has_tab = mode == 1 && trow != nullptr;
        uint4 tcur[kMaxH * kV];
#pragma unroll
        for (int i = 0; i < kMaxH * kV; ++i) tcur[i] = tpre[i];
        if (j + 1 < p.n_chunks) load_tab(trow, j + 1);
        else load_tab(trow_next, 0);
        uint8_t *stg = stage + wb * stage_bytes;
        mbar_wait(acc_full + 8 * wb, (g >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int hi = 0; hi < kMaxH; ++hi) {
          const int hd = half + 4 * hi;
          if (hd >= heads_per_chunk) continue;
          const int col = hd * DP;                       // first column of the head inside the chunk
          const int gcol = j * p.nc + col;               // ... inside the whole output row
          float y[DP];
          {
            uint32_t v[DP / 16][16];
#pragma unroll
            for (int c0 = 0; c0 < DP; c0 += 16) tmem_ld16(t_row + (uint32_t)(col + c0), v[c0 / 16]);
            tmem_ld_wait();
#pragma unroll
            for (int c0 = 0; c0 < DP; c0 += 16)
#pragma unroll
              for (int i = 0; i < 16; ++i) y[c0 + i] = __uint_as_float(v[c0 / 16][i]);
          }
          if (mode == 1) {
            if (has_tab) {
#pragma unroll
              for (int c = 0; c < kV; ++c) {
                const uint4 t4 = tcur[hi * kV + c];
                const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)            // two-wide FP32 add (FADD2)
                  unpack2(fadd2(pack2(y[c * 8 + 2 * i], y[c * 8 + 2 * i + 1]),
                                pack2(__uint_as_float(tw[i] << 16), __uint_as_float(tw[i] & 0xffff0000u))),
                          y[c * 8 + 2 * i], y[c * 8 + 2 * i + 1]);
              }
            } else if (p.bias) {
#pragma unroll
              for (int i = 0; i < DP; ++i) y[i] += bias_s[gcol + i];
            }
            if (p.normalize) {
              // F.normalize over the head (eps 1e-12), cosine_msa.py:152-153: y / max(|y|, 1e-12) = y * rsqrt(max(|y|^2, 1e-24));
              // FFMA2 / FMUL2 halve the per-element instruction count, rsqrt.approx replaces the IEEE sqrt + divide
              uint64_t ss2 = pack2(0.0f, 0.0f);
#pragma unroll
              for (int i = 0; i < DP; i += 2) {
                const uint64_t yy = pack2(y[i], y[i + 1]);
                ss2 = ffma2(yy, yy, ss2);
              }
              float ss_lo, ss_hi;
              unpack2(ss2, ss_lo, ss_hi);
              const float inv = rsqrtf(fmaxf(ss_lo + ss_hi, 1e-24f));
              const uint64_t inv2 = pack2(inv, inv);
#pragma unroll
              for (int i = 0; i < DP; i += 2) unpack2(fmul2(pack2(y[i], y[i + 1]), inv2), y[i], y[i + 1]);
            }
          } else {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < DP; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + gcol + i);
                y[i] += b4.x; y[i + 1] += b4.y; y[i + 2] += b4.z; y[i + 3] += b4.w;
              }
            }
            if (mode == 2) {
#pragma unroll
              for (int i = 0; i < DP; ++i) y[i] = gelu_erf_fast(y[i]);
            }
          }
#pragma unroll
          for (int c0 = 0; c0 < DP; c0 += 8) {
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[c0 + 2 * i], y[c0 + 2 * i + 1]);
              o[i] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            const int ch = (col + c0) >> 3;              // 16-byte chunk of the row inside the chunk's nc columns
            uint8_t *dst = tma_out ? stg + (ch >> 3) * kSlotBytes + r_tile * 128 + (((ch & 7) ^ (r_tile & 7)) << 4)
                                   : stg + r_tile * st_pitch + ch * 16;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        // this warp is done reading the accumulator: hand the buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + 8 * wb);
        if (tma_out) {
          // staging tile(s) complete -> ONE thread hands them to the TMA unit (rows past m are clipped by the tensor map);
          // the other staging buffer is free again once the previous chunk's store has read it (checked before the barrier)
          // (per TMEM lane quarter: its four warps own 32 rows of the tile -- a 128-thread named barrier and a 32-row store box
          // per quarter instead of one 512-thread barrier per chunk)
          fence_proxy_async();
          const bool issuer = half == 0 && lane == 0;
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
          if (issuer) {
            for (int t64 = 0; t64 < p.nc / 64; ++t64)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&p.tmap_out),
                           "r"(j * p.nc + t64 * 64), "r"((int)row0 + quarter * 32),
                           "r"(smem_u32(stg + t64 * kSlotBytes + quarter * 32 * 128))
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          // staging tile complete -> coalesced stores by all 512 epilogue threads -> staging free again
          asm volatile("bar.sync 1, 512;" ::: "memory");
          const int vec_per_row = p.nc / 8;              // 16-byte vectors per row of the chunk
          __nv_bfloat16 *obase = p.out + row0 * p.ldo + (int64_t)j * p.nc;
          int r = tid / vec_per_row, c = tid - r * vec_per_row;
          const int dr = (kEpiWarps * 32) / vec_per_row, dc = (kEpiWarps * 32) - dr * vec_per_row;
          for (; r < kTileM; r += dr) {
            if (row0 + r < p.m)
              *reinterpret_cast<uint4 *>(obase + (int64_t)r * p.ldo + c * 8) =
                  *reinterpret_cast<const uint4 *>(stg + r * st_pitch + c * 16);
            c += dc;
            if (c >= vec_per_row) { c -= vec_per_row; ++r; }
          }
          // (two staging buffers: the next chunk writes the other one; this one is rewritten after the next barrier)
        }
      }
      trow = trow_next;
    }
    if (tma_out && half == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
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

// chunk width / x buffers for (k, n, dp): the largest multiple of dp up to 128 that divides n and fits next to the x tile(s)
static bool plan(int k, int n, int n_norm, int dp, int *nc_out, int *a_bufs_out, int *smem_out) {
  if (k <= 0 || k % 8 || k > 512 || dp <= 0 || dp % 16 || dp > 128 || n <= 0 || n % dp || n_norm < 0 || n_norm > n || n_norm % dp)
    return false;
  const int ncb = (int)cdiv(k, kBlockK);
  const int budget = 227 * 1024 - 2048;
  // widest chunk first (more heads per chunk = more of the 16 epilogue warps busy, fewer hand-offs), then two x buffers
  for (int nc = 128 / dp * dp; nc >= dp; nc -= dp) {
    for (int a_bufs = 2; a_bufs >= 1; --a_bufs) {
      if (n % nc || n_norm % nc || n / nc > kMaxChunks) continue;
      const int stage_bytes = nc % 64 == 0 ? (nc / 64) * kSlotBytes : ((kTileM * (nc * 2 + 16) + 1023) & ~1023);
      const int bytes = a_bufs * ncb * kSlotBytes + 2 * ncb * nc * 128 + 2 * stage_bytes + 16 * 8 + 64 + n * 4;
      if (bytes <= budget) {
        *nc_out = nc;
        *a_bufs_out = a_bufs;
        *smem_out = bytes + 1024;
        return true;
      }
    }
  }
  return false;
}

}  // namespace qkv
}  // namespace os3d

using namespace os3d;

extern "C" int os3d_wide_linear_plan(int k, int n, int n_norm, int dp, int *nc) {
  int a_bufs = 0, smem = 0, c = 0;
  if (!qkv::plan(k, n, n_norm, dp, &c, &a_bufs, &smem)) return 0;
  if (nc) *nc = c;
  return 1;
}

extern "C" int os3d_wide_linear_bf16(const void *x, int64_t m, int k, int n, int dp, const void *w_img, const float *bias,
                                     const void *table, const int32_t *tab_idx, int64_t tab_ld, int n_norm, int normalize,
                                     int mode_rest, void *out, int64_t ldo, void *stream) {
  int nc = 0, a_bufs = 0, smem = 0;
  if (!qkv::plan(k, n, n_norm, dp, &nc, &a_bufs, &smem) || m < 0 || ((uintptr_t)x & 15) || ((uintptr_t)out & 15) || ldo < n || ldo % 8 ||
      (mode_rest != 0 && mode_rest != 2) ||
      (n_norm > 0 && table && (!tab_idx || tab_ld < n_norm || tab_ld % 8)))
    return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  qkv::encode_tiled_fn enc = qkv::encode_tiled();
  if (!enc) return OS3D_ERR_BAD_ARG;
  qkv::Params p;
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)m};
    const cuuint64_t gstr[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)qkv::kBlockK, (cuuint32_t)qkv::kTileM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&p.tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(x), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OS3D_ERR_BAD_ARG;
  }
  p.tma_store = nc % 64 == 0;
  if (p.tma_store) {
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)m};
    const cuuint64_t gstr[1] = {(cuuint64_t)ldo * 2};
    const cuuint32_t box[2] = {64u, 32u};                 // one TMEM lane quarter's rows
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&p.tmap_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return OS3D_ERR_BAD_ARG;
  } else {
    p.tmap_out = p.tmap_x;
  }
  p.w_img = (const __nv_bfloat16 *)w_img;
  p.bias = bias;
  p.table = (const __nv_bfloat16 *)table;
  p.tab_idx = tab_idx;
  p.tab_ld = tab_ld;
  p.out = (__nv_bfloat16 *)out;
  p.m = m;
  p.ldo = ldo;
  p.n_tiles = (int)cdiv(m, qkv::kTileM);
  p.k16 = (k + 15) / 16 * 16;
  p.ncb = (int)cdiv(k, qkv::kBlockK);
  p.nc = nc;
  p.n_chunks = n / nc;
  p.dp = dp;
  p.a_bufs = a_bufs;
  p.mode_n1 = n_norm / nc;
  p.mode_rest = mode_rest;
  p.normalize = normalize;
  p.idesc = ptx::make_idesc_bf16(qkv::kTileM, nc);
  int dev = 0, sms = 148;
  OS3D_CUDA(cudaGetDevice(&dev));
  OS3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static int configured_dev[64][4] = {{0}};
  const int di = dp == 16 ? 0 : dp == 32 ? 1 : dp == 48 ? 2 : 3;
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  cudaStream_t st = (cudaStream_t)stream;
#define OS3D_QKV_LAUNCH(DPV)                                                                                              \
  do {                                                                                                                    \
    if (dev < 0 || dev >= 64 || !configured_dev[dev][di]) {                                                               \
      OS3D_CUDA(cudaFuncSetAttribute(qkv::qkv_proj_tc_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      if (dev >= 0 && dev < 64) configured_dev[dev][di] = 1;                                                              \
    }                                                                                                                     \
    qkv::qkv_proj_tc_kernel<DPV><<<grid, qkv::kThreads, smem, st>>>(p);                                                   \
  } while (0)
  if (dp == 16) OS3D_QKV_LAUNCH(16);
  else if (dp == 32) OS3D_QKV_LAUNCH(32);
  else if (dp == 48) OS3D_QKV_LAUNCH(48);
  else if (dp == 64) OS3D_QKV_LAUNCH(64);
  else return OS3D_ERR_BAD_ARG;
#undef OS3D_QKV_LAUNCH
  OS3D_LAUNCH_CHECK();
  return 0;
}
