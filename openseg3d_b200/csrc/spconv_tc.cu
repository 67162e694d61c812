// spconv_tc.cu -- stage 3, bf16 mode: output-stationary sparse convolution on the tcgen05 tensor cores (sm_100a).
//
//   out[r, :] = epilogue( sum_k  in[nbr[r, k], :] . W[k] )        bf16 in / bf16 out, fp32 accumulate in TMEM
//
// What bounds this op on B200 is the L2 -> SM fabric, not the tensor pipe: every 128-row tile pulls its gathered rows
// (27 * Cin * 2 B * 128 * fill) AND a full copy of the weights (27 * Cin * Cout * 2 B) out of L2 (ncu of the first
// one-tile-per-CTA kernel on the L2 192->96 layer: 13.5 GB through the crossbar, 7.9 GB of it weights,
// profiles/r01b_spconv_l2_192_96_one_tile.txt).  So one CTA owns T consecutive 128-row tiles with T accumulators side
// by side in tensor memory (T * Cout <= 512 fp32 columns), and a weight K-block staged once in shared memory feeds all
// T tiles: weight traffic per output row drops T-fold.  The contraction axis is (offset k, channel block of 64).
//
//   warps 0-7 producers : gather.  Thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 32 j of a 128-row slot; each
//            chunk is one cp.async (LDGSTS, zero-fill when the row has no neighbour at that offset) straight into the
//            SWIZZLE_128B K-major UMMA layout.  Neighbour rows come from the offset-major table ([27][m_pad]) and are
//            prefetched one offset ahead into registers; each warp keeps 3 slots of copies in flight
//            (cp.async groups) and signals the oldest with ONE mbarrier arrive per warp.
//   warp 9  weights  : the
//            packed image is stored already swizzled, one contiguous [Cout x 128 B] slab per K-block = ONE
//            cp.async.bulk (TMA bulk copy, mbarrier complete_tx).
//            (TMA tile::gather4 was built and measured for this gather first: ~40 cycles per 4-row request per SM,
//            12.8 B/clk/SM, 1.3-1.9x slower than the LDGSTS gather on every layer -- DESIGN.md "Measured dead ends".)
//   warp 8  MMA      : one thread issues tcgen05.mma.kind::f16 (M = 128, N = Cout, K = 16 per instruction) into the
//            tile's accumulator; tcgen05.commit frees the A slot / the weight slot.
//   warps 0-7 epilogue (after their gather loop): tcgen05.ld, y = acc*scale + shift (+ residual) (ReLU)
//            (+ pair-summed residual), bf16, 32-byte vector stores.
//   (tile, offset) pairs with no neighbour anywhere in the tile are skipped by all roles (per-tile 27-bit masks).
//
// replaces: SubMConv3d / SparseConv3d / SparseInverseConv3d forward of spconv-cu113 (+ the BatchNorm1d / ReLU / add
// that follow them in seg3d/utils/spconv_utils.py:26-30 and seg3d/models/backbones/pointtransformer.py:47-66,89-110).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace os3d {
namespace tc {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                        // bf16 elements per K-block = one 128-byte swizzle row
constexpr int kATileBytes = kTileM * 128;          // 16 KB
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;     // 8 gather / epilogue warps + the MMA warp + the weight-loader warp
constexpr int kMaxTiles = 10;                      // accumulators per CTA
constexpr int kMaxSA = 8, kMaxSB = 4;             // ring depths (A tiles / weight slabs)

struct Params {
  const __nv_bfloat16 *in;    // [m_in, cin] bf16
  int cin;                    // row pitch in elements (multiple of 8)
  const int32_t *nbr_t;       // [27][m_pad] offset-major neighbour rows (-1 = none)
  const uint32_t *tile_mask;  // [m_pad / 128] offsets present in each 128-row tile
  const int32_t *perm;        // [m_out] output row computed by each row of the (mask-grouped) tile order; NULL = identity
  int64_t m_out, m_pad;
  int n_tiles;
  int cin16;                  // cin rounded up to 16 (MMA K granularity; the excess columns are zero on both operands)
  int cout, ncb;              // ncb = ceil(cin / 64) channel blocks per offset
  const __nv_bfloat16 *w_img; // [27 * ncb][cout][64] swizzled weight image
  const float *scale, *shift;
  const __nv_bfloat16 *residual;
  int flags;
  __nv_bfloat16 *out;
  int64_t ldo;                // output row pitch in elements (>= cout)
  int dense;                  // 1: the "kernel map" is the identity with one offset -- a plain Linear layer  (os3d_linear_bf16)
  const float *ln_gamma, *ln_beta;   // flags & 8: LayerNorm over the cout columns of y before the residual add
  float ln_eps;
  const __nv_bfloat16 *table; // flags & 16: y[r, c] += table[tab_idx[r], c] for c < tab_cols (position-embedding term)
  const int32_t *tab_idx;
  int tab_cols;
  int tmem_cols, n_parts, n_per_part;
  uint32_t idesc;
  int tiles_per_cta, sa, sb;
  int sa_log2;
  int issuers;                // MMA-issuing warps: 2 when at most four warps gather and the CTA owns >= 2 tiles (warp 7 issues the odd tiles)
};

__global__ void __launch_bounds__(kThreads, 2) spconv_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw);
  const int b_bytes = p.cout * 128;
  const uint32_t a_base = base;
  const uint32_t b_base = base + p.sa * kATileBytes;
  uint8_t *tail = smem + p.sa * kATileBytes + p.sb * b_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);   // a_full[kMaxSA] a_empty[kMaxSA] b_full[kMaxSB] b_empty[kMaxSB] accum
  uint32_t *misc = reinterpret_cast<uint32_t *>(bars + 2 * kMaxSA + 2 * kMaxSB + 1);   // [0] tmem base, [1..] tile masks
  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + kMaxSA);
  const uint32_t b_full = smem_u32(bars + 2 * kMaxSA), b_empty = smem_u32(bars + 2 * kMaxSA + kMaxSB);
  const uint32_t accum_bar = smem_u32(bars + 2 * kMaxSA + 2 * kMaxSB);
  uint32_t *masks_s = misc + 1;
  // epilogue parameters staged in shared memory: with the carve-out at its maximum there is no L1 left, so a __ldg per
  // 16-column chunk is a ~300-cycle trip to L2 in every epilogue iteration (ncu: 17 % of all stall samples)
  float *prm_s = reinterpret_cast<float *>(misc + 18);    // [scale | shift | gamma | beta], cout floats each (16-byte aligned)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  const int ntile = min(p.tiles_per_cta, p.n_tiles - tile0);

  if (tid == 0) {
    for (int s = 0; s < p.sa; ++s) { mbar_init(a_full + 8 * s, 32); mbar_init(a_empty + 8 * s, 1); }
    for (int s = 0; s < p.sb; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, (uint32_t)p.issuers); }
    mbar_init(accum_bar, (uint32_t)p.issuers);
    fence_barrier_init();
  }
  if (tid < kMaxTiles) masks_s[tid] = tid < ntile ? (p.dense ? 1u : __ldg(p.tile_mask + tile0 + tid)) : 0u;
  for (int i = tid; i < p.cout; i += kThreads) {
    prm_s[i] = p.scale ? __ldg(p.scale + i) : 1.0f;
    prm_s[p.cout + i] = p.shift ? __ldg(p.shift + i) : 0.0f;
    if (p.flags & 8) {
      prm_s[2 * p.cout + i] = __ldg(p.ln_gamma + i);
      prm_s[3 * p.cout + i] = __ldg(p.ln_beta + i);
    }
  }
  __syncthreads();
  if (warp == kEpiWarps) tmem_alloc(smem_u32(&misc[0]), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  uint32_t any = 0;
#pragma unroll
  for (int t = 0; t < kMaxTiles; ++t) any |= masks_s[t];

  // ---- the MMA-issuing loop (warp 8; with two issuers also warp 7, which takes the odd tiles of the CTA) ----
  auto issue_loop = [&](const uint32_t which) {
    {
      int sb_i = 0;
      uint32_t q = 0, b_ph = 0, started = 0, ready = 0;      // q: slots consumed so far
      // Two issuers = two independent pipelines: issuer w consumes the tiles t with (t & 1) == w out of ring slots
      // [w sa/2, (w + 1) sa/2), filled by the gather warps of the same half.  (One shared ring would let an issuer run a
      // whole revolution ahead of the other on a slot -- an mbarrier parity cannot tell that from "filled".)
      const bool two_issuers = p.issuers == 2;
      const uint32_t sa_shift = (uint32_t)p.sa_log2 - (two_issuers ? 1u : 0u), sa_mask = (1u << sa_shift) - 1u;
      const uint32_t slot0 = two_issuers ? which << sa_shift : 0u;
      const uint32_t tmask = two_issuers ? (which ? 0xaaaaaaaau : 0x55555555u) : 0xffffffffu;
      const uint32_t desc_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
      const uint32_t a_lo0 = (uint32_t)make_kmajor_sw128_desc(a_base), b_lo0 = (uint32_t)make_kmajor_sw128_desc(b_base);
      const uint32_t b_step = (uint32_t)b_bytes >> 4, part_lo = (uint32_t)(p.n_per_part * 128) >> 4;
      const bool two_parts = p.n_parts == 2;
      const uint32_t idesc = p.idesc, cout = (uint32_t)p.cout, npp = (uint32_t)p.n_per_part;
      const int last_steps = (p.cin16 - (p.ncb - 1) * kBlockK) >> 4;
      for (int k = 0; k < OS3D_KVOL; ++k) {
        if (!((any >> k) & 1u)) continue;
        uint32_t tiles_k = 0;                              // tiles of this CTA that have offset k
        for (int t = 0; t < ntile; ++t) tiles_k |= ((masks_s[t] >> k) & 1u) << t;
        tiles_k = __shfl_sync(0xffffffffu, tiles_k, 0);
        for (int cb = 0; cb < p.ncb; ++cb) {
          const int steps = cb + 1 < p.ncb ? kBlockK / 16 : last_steps;
          mbar_wait(b_full + 8 * sb_i, b_ph);
          const uint32_t b_lo = b_lo0 + sb_i * b_step;
          for (uint32_t rem = tiles_k & tmask; rem; rem &= rem - 1) {
            const uint32_t t = __ffs(rem) - 1;
            const uint32_t slot = slot0 + (q & sa_mask);
            if (!ready) mbar_wait(a_full + 8 * slot, (q >> sa_shift) & 1u);
            // probe the NEXT slot's barrier now: its round trip overlaps the MMAs issued below
            ready = mbar_test(a_full + 8 * (slot0 + ((q + 1) & sa_mask)), ((q + 1) >> sa_shift) & 1u);
            fence_proxy_async();   // LDGSTS (generic proxy) writes observed through the barrier -> visible to the MMA's async proxy
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + slot * (kATileBytes >> 4);
            const uint32_t d = tmem_base + t * cout;
            const uint32_t acc0 = (started >> t) & 1u;
            if (elect_one()) {
              umma_bf16_lo(d, a_lo, b_lo, desc_hi, idesc, acc0);
              if (two_parts) umma_bf16_lo(d + npp, a_lo, b_lo + part_lo, desc_hi, idesc, acc0);
#pragma unroll
              for (int ks = 1; ks < kBlockK / 16; ++ks) {
                if (ks < steps) {
                  umma_bf16_lo(d, a_lo + 2 * ks, b_lo + 2 * ks, desc_hi, idesc, 1u);
                  if (two_parts) umma_bf16_lo(d + npp, a_lo + 2 * ks, b_lo + part_lo + 2 * ks, desc_hi, idesc, 1u);
                }
              }
              umma_commit(a_empty + 8 * slot);
            }
            __syncwarp();
            started |= 1u << t;
            ++q;
          }
          if (elect_one()) umma_commit(b_empty + 8 * sb_i);
          __syncwarp();
          if (++sb_i == p.sb) { sb_i = 0; b_ph ^= 1; }
        }
      }
      if (elect_one()) {
        if (started) umma_commit(accum_bar);
        else mbar_arrive(accum_bar);
      }
    }
  };

  if (warp < kEpiWarps) {
    // ================================ producers ================================
    // A full / empty mbarrier round trip costs ~300 cycles per thread (tools/bench_mbar.cu), whatever the ring depth:
    // with all warps cooperating on every slot that was the pace of the whole gather.  So each warp fills WHOLE slots
    // on its own: slot u of the (offset, channel block, tile) sequence goes to warp u mod n_prod and always lands in
    // ring slot u mod n_prod, with that slot's own barrier pair.  Lane l owns 16-byte chunk (l & 7) of rows
    // 32 (l >> 3) + i, i = 0..31: 8 lanes cover one 128-byte row, so each global request is a full line.  The 128
    // neighbour rows of a slot are ONE coalesced int4 per lane from the offset-major table (prefetched a slot ahead);
    // the row of iteration i is then a warp shuffle away.  Completion: cp.async.mbarrier.arrive.noinc (32 per slot).
    if (warp < p.sa) {
      const uint32_t c = lane & 7, g = lane >> 3;
      // (two issuers: the warps of each half of the ring serve one issuer's tiles -- even tiles, odd tiles)
      const bool two_issuers = p.issuers == 2;
      const uint32_t nprod_mask = ((uint32_t)p.sa >> (two_issuers ? 1 : 0)) - 1u;
      const uint32_t slot = (uint32_t)warp, slot_g = (uint32_t)warp & nprod_mask;          // ring slot; index inside the half
      const uint32_t tmask = two_issuers ? (((uint32_t)warp > nprod_mask) ? 0xaaaaaaaau : 0x55555555u) : 0xffffffffu;
      const uint32_t dst_base = a_base + slot * kATileBytes + g * 32 * 128;
      const uint32_t full_bar = a_full + 8 * slot, empty_bar = a_empty + 8 * slot;
      const char *in_bytes = reinterpret_cast<const char *>(p.in);
      const uint32_t row_bytes = (uint32_t)p.cin * 2u;
      const int32_t *nt = p.nbr_t + (int64_t)tile0 * kTileM + 4 * lane;
      // Iterator over this warp's slots.  Per (offset, channel block) the CTA's slots are the set bits of tiles_k in
      // ascending tile order; the warp owns those whose running index is congruent to its ring slot.
      int ik = -1, icb = p.ncb - 1, it = 0;                  // the first advance moves to offset 0, block 0
      uint32_t u0 = 0, tiles_k = 0, cnt = 0, r = 0;           // u0: slots before this (k, cb); r: my next rank in it
      auto next_mine = [&]() -> bool {
        while (true) {
          if (r < cnt) {
            it = __fns(tiles_k, 0, r + 1);
            r += nprod_mask + 1;
            return true;
          }
          u0 += cnt;
          if (++icb >= p.ncb) {
            icb = 0;
            do {
              if (++ik >= OS3D_KVOL) return false;
            } while (!((any >> ik) & 1u));
            tiles_k = 0;
            for (int t = 0; t < ntile; ++t) tiles_k |= ((masks_s[t] >> ik) & 1u) << t;
            tiles_k &= tmask;
            cnt = __popc(tiles_k);
          }
          r = (slot_g - u0) & nprod_mask;
        }
      };
      auto fetch_rows = [&]() -> int4 {
        if (!p.dense) return __ldg(reinterpret_cast<const int4 *>(nt + (int64_t)ik * p.m_pad + it * kTileM));
        const int64_t r = (int64_t)(tile0 + it) * kTileM + 4 * lane;      // Linear: row r of the output reads row r
        return make_int4(r < p.m_out ? (int)r : -1, r + 1 < p.m_out ? (int)r + 1 : -1, r + 2 < p.m_out ? (int)r + 2 : -1,
                         r + 3 < p.m_out ? (int)r + 3 : -1);
      };
      // neighbour rows are prefetched TWO slots ahead: one slot's work is shorter than a trip to L2 / HBM for the table
      int4 v0 = make_int4(-1, -1, -1, -1), v1 = v0;
      int cb0 = 0, cb1 = 0;
      bool have0 = next_mine();
      if (have0) { cb0 = icb; v0 = fetch_rows(); }
      bool have1 = have0 && next_mine();
      if (have1) { cb1 = icb; v1 = fetch_rows(); }
      uint32_t ph = 1;                                     // empty barriers start "free"
      while (have0) {
        const int4 cur = v0;
        const uint32_t ch = (uint32_t)cb0 * kBlockK + c * 8;
        const bool ch_ok = ch < (uint32_t)p.cin;            // channels >= cin of the last channel block: zero-fill
        const char *src0 = in_bytes + ch * 2u;
        have0 = have1; v0 = v1; cb0 = cb1;
        have1 = have1 && next_mine();
        if (have1) { cb1 = icb; v1 = fetch_rows(); }
        mbar_wait(empty_bar, ph);
        ph ^= 1;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int comp = (i & 3) == 0 ? cur.x : (i & 3) == 1 ? cur.y : (i & 3) == 2 ? cur.z : cur.w;
          const int n = __shfl_sync(0xffffffffu, comp, (lane & 24) + (i >> 2));
          cp_async_16(dst_base + i * 128 + ((c ^ (uint32_t)(i & 7)) << 4), src0 + (uint64_t)((uint32_t)max(n, 0) * row_bytes),
                      (ch_ok && n >= 0) ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(full_bar);
      }
    } else if (p.issuers == 2 && warp == kEpiWarps - 1) {
      issue_loop(1);          // second MMA issuer: the serial per-slot work (wait, proxy fence, issue, commit) of the odd tiles
      __syncwarp();
    }
  } else if (warp == kEpiWarps + 1) {
    // ================================ weight loader ================================
    // its own warp, so that waiting for a free weight slot never delays the gather
    if (lane == 0) {
      int sb_i = 0;
      uint32_t b_ph = 1;
      for (int k = 0; k < OS3D_KVOL; ++k) {
        if (!((any >> k) & 1u)) continue;
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(b_empty + 8 * sb_i, b_ph);
          mbar_arrive_expect_tx(b_full + 8 * sb_i, (uint32_t)b_bytes);
          bulk_g2s(b_base + sb_i * b_bytes, p.w_img + (int64_t)(k * p.ncb + cb) * p.cout * kBlockK, (uint32_t)b_bytes,
                   b_full + 8 * sb_i);
          if (++sb_i == p.sb) { sb_i = 0; b_ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ MMA issuer ================================
    // One thread feeds the tensor pipe, so its instruction count per slot IS the pipeline's pace for small N (ncu of
    // the first version: 130 instructions per slot, tensor pipe 11 % busy, producers blocked on empty slots).  The loop
    // is therefore stripped to: barrier wait, fence, 1-8 UTCHMMA whose descriptors differ only in their low word, commit.
    // The WHOLE warp runs this loop on warp-uniform values; only the tcgen05 instructions are under elect.sync.
    issue_loop(0);
    __syncwarp();
  }
  if (warp < kEpiWarps) {
    // ================================ epilogue ================================
    // The gather warps wait for the last MMA on the mbarrier; the warps that gathered nothing have been parked on a
    // hardware named barrier since the start of the main loop (a spinning mbarrier wait re-issues try_wait + branch
    // every ~18 cycles on the scheduler it shares with a gather warp: ncu counted 54 % of the kernel's executed warp
    // instructions as barrier polls) and are released by the gather warps' arrival.
    if (warp < p.sa) {
      mbar_wait(accum_bar, 0);
      tc_fence_before();
    }
    asm volatile("bar.sync 5, %0;" ::"n"(kEpiWarps * 32) : "memory");      // (ids 1-4: the quarter pairs of the LayerNorm epilogue)
    tc_fence_after();
    const int quarter = warp & 3;                       // TMEM lanes [32 q, 32 q + 32) are visible to warps q and q + 4
    const bool pair_sum = (p.flags & 2) != 0;           // residual rows hold 2*cout channels; add r[2c] + r[2c+1] after the ReLU
    const bool do_relu = (p.flags & 1) != 0, do_gelu = (p.flags & 4) != 0, do_ln = (p.flags & 8) != 0;
    const bool do_tab = (p.flags & 16) != 0;
    // y[0..16) = acc * scale + shift for columns [col, col + 16) of this thread's row of tile t
    auto load16 = [&](int t, int col, bool has, float (&y)[16]) {
      uint32_t v[16];
      if (has) {
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * p.cout + col), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 sc = *reinterpret_cast<const float4 *>(prm_s + col + 4 * q);
        const float4 sh = *reinterpret_cast<const float4 *>(prm_s + p.cout + col + 4 * q);
        y[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), sc.x, sh.x);
        y[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), sc.y, sh.y);
        y[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), sc.z, sh.z);
        y[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), sc.w, sh.w);
      }
    };
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
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const int half = warp >> 2;                          // the two warps of a lane quarter take alternate 16-column chunks
    if (do_ln) {
      // out = residual + LayerNorm(y) * gamma + beta.  A thread owns a row (lane = row) of alternate 16-column chunks;
      // the two warps of a lane quarter exchange their partial (sum, sum of squares) through shared memory (the A
      // ring is free by now) at a 64-thread named barrier.  Two passes over the accumulator in TMEM; the residual
      // rows (global, uncoalesced, L2 latency) are fetched one chunk ahead.
      // (EncoderLayer: x + LN(attn(x)), x + LN(mlp(x)), point_transformer_layer.py:288-298.)
      float2 *part = reinterpret_cast<float2 *>(smem);                 // [8 warps][32 lanes]
      for (int t = 0; t < ntile; ++t) {
        const bool has = masks_s[t] != 0u;
        const int64_t vrow = (int64_t)(tile0 + t) * kTileM + quarter * 32 + lane;
        const bool row_ok = vrow < p.m_out;
        const int64_t row = (row_ok && p.perm) ? (int64_t)__ldg(p.perm + vrow) : vrow;
        const __nv_bfloat16 *rrow = (p.residual && row_ok) ? p.residual + row * p.cout : nullptr;
        uint4 na = zero4, nb = zero4;
        if (rrow && half * 16 < p.cout) {
          na = __ldg(reinterpret_cast<const uint4 *>(rrow + half * 16));
          nb = __ldg(reinterpret_cast<const uint4 *>(rrow + half * 16) + 1);
        }
        float sum = 0.0f, sq = 0.0f;
        for (int col = half * 16; col < p.cout; col += 32) {
          float y[16];
          load16(t, col, has, y);
#pragma unroll
          for (int i = 0; i < 16; ++i) { sum += y[i]; sq = fmaf(y[i], y[i], sq); }
        }
        part[warp * 32 + lane] = make_float2(sum, sq);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        const float2 other = part[(warp ^ 4) * 32 + lane];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // partner has read before the next tile overwrites
        sum += other.x;
        sq += other.y;
        const float mean = sum / (float)p.cout;
        const float rstd = rsqrtf(fmaxf(sq / (float)p.cout - mean * mean, 0.0f) + p.ln_eps);
        for (int col = half * 16; col < p.cout; col += 32) {
          const uint4 ra = na, rb = nb;
          if (rrow && col + 32 < p.cout) {
            na = __ldg(reinterpret_cast<const uint4 *>(rrow + col + 32));
            nb = __ldg(reinterpret_cast<const uint4 *>(rrow + col + 32) + 1);
          }
          float y[16];
          load16(t, col, has, y);
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 g = *reinterpret_cast<const float4 *>(prm_s + 2 * p.cout + col + 4 * q);
              const float4 be = *reinterpret_cast<const float4 *>(prm_s + 3 * p.cout + col + 4 * q);
              y[4 * q + 0] = fmaf((y[4 * q + 0] - mean) * rstd, g.x, be.x);
              y[4 * q + 1] = fmaf((y[4 * q + 1] - mean) * rstd, g.y, be.y);
              y[4 * q + 2] = fmaf((y[4 * q + 2] - mean) * rstd, g.z, be.z);
              y[4 * q + 3] = fmaf((y[4 * q + 3] - mean) * rstd, g.w, be.w);
            }
            if (rrow) add_bf16x16(ra, rb, y);
            store16(p.out + row * p.ldo + col, y);
          }
        }
      }
    } else {
      for (int t = 0; t < ntile; ++t) {
        const bool has = masks_s[t] != 0u;
        const int64_t vrow = (int64_t)(tile0 + t) * kTileM + quarter * 32 + lane;
        const bool row_ok = vrow < p.m_out;
        const int64_t row = (row_ok && p.perm) ? (int64_t)__ldg(p.perm + vrow) : vrow;
        __nv_bfloat16 *orow = p.out + row * p.ldo;
        const __nv_bfloat16 *rrow = (p.residual && row_ok) ? p.residual + row * p.cout * (pair_sum ? 2 : 1) : nullptr;
        const __nv_bfloat16 *trow = (do_tab && row_ok) ? p.table + (int64_t)__ldg(p.tab_idx + row) * p.tab_cols : nullptr;
        // additive row terms (table row before the activation, plain residual) are fetched one chunk ahead
        uint4 nx[4] = {zero4, zero4, zero4, zero4};
        auto fetch = [&](int col) {
          if (trow && col < p.tab_cols) {
            nx[0] = __ldg(reinterpret_cast<const uint4 *>(trow + col));
            nx[1] = __ldg(reinterpret_cast<const uint4 *>(trow + col) + 1);
          } else {
            nx[0] = nx[1] = zero4;
          }
          if (rrow && !pair_sum) {
            nx[2] = __ldg(reinterpret_cast<const uint4 *>(rrow + col));
            nx[3] = __ldg(reinterpret_cast<const uint4 *>(rrow + col) + 1);
          }
        };
        if (half * 16 < p.cout) fetch(half * 16);
        for (int col = half * 16; col < p.cout; col += 32) {
          const uint4 c0 = nx[0], c1 = nx[1], c2 = nx[2], c3 = nx[3];
          if (col + 32 < p.cout) fetch(col + 32);
          float y[16];
          load16(t, col, has, y);
          if (row_ok) {
            add_bf16x16(c0, c1, y);
            if (rrow && !pair_sum) add_bf16x16(c2, c3, y);
            if (do_relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.0f);
            }
            if (do_gelu) {               // exact (erf) GELU, nn.GELU() default (point_transformer_layer.py:266)
#pragma unroll
              for (int i = 0; i < 16; ++i) y[i] = 0.5f * y[i] * (1.0f + erff(y[i] * 0.70710678118654752f));
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
            store16(orow + col, y);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---- tile order of a kernel map --------------------------------------------------------------------------------
// An output-stationary tile does the MMAs of every offset that ANY of its 128 rows has, so rows with the same set of
// neighbour offsets should share tiles.  In storage order a tile of the synthetic Waymo frames has 24.9 (L1 subm) /
// 26.6 (inverse convs) of the 27 offsets; sorted by mask (similar masks end up adjacent) inside blocks of consecutive rows
// it has 13.8 / 3.4 with 64 k-row blocks and the plain mask as key, 10.8 / 3.4 with 256 k-row blocks and the
// frequency-ranked bit order below (the inverse conv's rows fall into the 8 parity classes either way).  Blocks, not a
// global sort, keep the gather local: a tile's neighbours stay within a window of rows that the concurrent CTAs share
// in L2.  The sort is one cub radix sort of 32-bit keys (block << 27 | mask) with the row index as value.
constexpr int kOrderRowsPerCta = 256;

// Significance of the 27 offsets in the sort key, least significant first: the centre, the in-plane axis neighbours, the
// in-plane diagonals, the two vertical neighbours, the vertical edges, the corners -- i.e. by falling frequency in lidar
// maps (a submanifold map is symmetric, k and 26 - k are equally frequent; the ranking is the same at every level).  With
// the RARE offsets in the high key bits, rows that lack them end up in the same tiles and those tiles skip the offsets:
// 13.97 -> 12.35 active offsets per 128-row tile at level 1, 19.81 -> 19.30 at level 2, strided maps 9.19 -> 8.83 /
// 15.49 -> 14.75, inverse maps unchanged at 3.4 (CPU simulation on the oracle's maps of two frames, tools/sim_tile_order.py).
__constant__ int8_t kOrderBit[OS3D_KVOL] = {13, 10, 12, 14, 16, 9, 11, 15, 17, 4, 22, 1, 3, 5, 7, 19, 21, 23, 25, 0, 2, 6, 8, 18, 20, 24, 26};

// sort key of every row: (block of consecutive rows) << 27 | 27-bit offset mask, bits permuted as above; coalesced reads of
// the row-major map
__global__ void __launch_bounds__(256) order_keys_kernel(const int32_t *__restrict__ nbr, int64_t m, int block_shift,
                                                         uint32_t *__restrict__ keys, int32_t *__restrict__ rows) {
  __shared__ uint32_t bits_s[kOrderRowsPerCta];
  const int64_t row0 = (int64_t)blockIdx.x * kOrderRowsPerCta;
  bits_s[threadIdx.x] = 0;
  __syncthreads();
  const int n = (int)min((int64_t)kOrderRowsPerCta, m - row0);
  for (int t = threadIdx.x; t < n * OS3D_KVOL; t += 256) {
    const int r = t / OS3D_KVOL;
    if (__ldg(nbr + row0 * OS3D_KVOL + t) >= 0) atomicOr(&bits_s[r], 1u << (t - r * OS3D_KVOL));
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    const int64_t r = row0 + threadIdx.x;
    const uint32_t mask = bits_s[threadIdx.x];
    uint32_t key = 0;
#pragma unroll
    for (int j = 0; j < OS3D_KVOL; ++j) key |= ((mask >> kOrderBit[j]) & 1u) << j;
    keys[r] = ((uint32_t)(r >> block_shift) << OS3D_KVOL) | key;
    rows[r] = (int32_t)r;
  }
}

// [m, 27] row-major kernel map -> offset-major [27][m_pad] in tile order (row v of the tile order is output row perm[v];
// m_pad = 128 * n_tiles, padding rows -1) + per-tile masks.
__global__ void __launch_bounds__(256) kernel_map_tiles_kernel(const int32_t *__restrict__ nbr, int64_t m, int64_t m_pad,
                                                               const int32_t *__restrict__ perm,
                                                               int32_t *__restrict__ nbr_t,
                                                               uint32_t *__restrict__ tile_mask) {
  __shared__ int32_t s[kTileM * OS3D_KVOL];
  __shared__ int32_t rows_s[kTileM];
  __shared__ uint32_t msk;
  const int64_t row0 = (int64_t)blockIdx.x * kTileM;
  if (threadIdx.x == 0) msk = 0;
  if (threadIdx.x < kTileM) {
    const int64_t v = row0 + threadIdx.x;
    rows_s[threadIdx.x] = v < m ? (perm ? __ldg(perm + v) : (int32_t)v) : -1;
  }
  __syncthreads();
  uint32_t mask = 0;
  for (int t = threadIdx.x; t < kTileM * OS3D_KVOL; t += 256) {
    const int r = t / OS3D_KVOL, k = t - r * OS3D_KVOL;
    const int32_t src = rows_s[r];
    const int32_t v = src >= 0 ? __ldg(nbr + (int64_t)src * OS3D_KVOL + k) : -1;
    s[t] = v;
    if (v >= 0) mask |= 1u << k;
  }
  mask = __reduce_or_sync(0xffffffffu, mask);
  if ((threadIdx.x & 31) == 0 && mask) atomicOr(&msk, mask);
  __syncthreads();
  for (int t = threadIdx.x; t < kTileM * OS3D_KVOL; t += 256) {
    const int k = t >> 7, r = t & 127;
    nbr_t[(int64_t)k * m_pad + row0 + r] = s[r * OS3D_KVOL + k];
  }
  if (threadIdx.x == 0) tile_mask[blockIdx.x] = msk;
}

// spconv 2.x weight [cout, 27, cin] f32 -> bf16 UMMA image [27 * ncb][cout][8 chunks, XOR-swizzled by row & 7][8]:
// exactly the bytes a weight K-block occupies in shared memory, so one bulk copy loads it.  K-blocks never straddle
// kernel offsets (the last channel block of an offset is zero-padded to 64).
__global__ void pack_weight_img_kernel(const float *__restrict__ src, int cin, int cout, int ncb, int kvol,
                                       __nv_bfloat16 *__restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)kvol * ncb * cout * kBlockK;
  if (t >= total) return;
  const int e = (int)(t & 7);
  const int pc = (int)((t >> 3) & 7);
  const int n = (int)((t >> 6) % cout);
  const int blk = (int)((t >> 6) / cout);
  const int c = pc ^ (n & 7);  // logical chunk stored at physical chunk pc
  const int k = blk / ncb, ch = (blk - k * ncb) * kBlockK + c * 8 + e;
  dst[t] = __float2bfloat16(ch < cin ? src[((int64_t)n * kvol + k) * cin + ch] : 0.0f);
}

}  // namespace tc
}  // namespace os3d

using namespace os3d;

// Blocks of >= 262144 rows, at most 32 of them (5 key bits above the mask).  Measured on the 8-frame batch (conv time per
// step): 64 k rows 11.71 ms, 256 k rows 11.37, 512 k 11.39, 1 M 11.40 -- larger blocks group the masks better (level 1:
// 12.35 -> 10.75 active offsets per tile) until the gathers of concurrent CTAs stop sharing L2 lines (the 384-channel
// inverse conv loses 5 % at 1 M rows).
static int order_block_shift(int64_t m) {
  int shift = 18;
  { const char *e = getenv("OS3D_ORDER_BLOCK_SHIFT"); if (e && atoi(e) >= 10 && atoi(e) <= 30) shift = atoi(e); }
  while ((m >> shift) >= 32) ++shift;
  return shift;
}

extern "C" int os3d_kernel_map_order_scratch(int64_t m, int64_t *temp_bytes) {
  if (m < 0 || m > 0x7fffffff) return OS3D_ERR_BAD_ARG;
  size_t bytes = 0;
  OS3D_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (const int32_t *)nullptr, (int32_t *)nullptr, (int)m, 0, 32));
  *temp_bytes = (int64_t)bytes;
  return 0;
}

extern "C" int os3d_kernel_map_order(const int32_t *nbr, int64_t m, uint32_t *keys, uint32_t *keys_sorted, int32_t *rows,
                                     int32_t *perm, void *temp, int64_t temp_bytes, void *stream) {
  if (m < 0 || m > 0x7fffffff) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int shift = order_block_shift(m);
  tc::order_keys_kernel<<<(unsigned)cdiv(m, tc::kOrderRowsPerCta), 256, 0, st>>>(nbr, m, shift, keys, rows);
  OS3D_LAUNCH_CHECK();
  size_t bytes = (size_t)temp_bytes;
  OS3D_CUDA(cub::DeviceRadixSort::SortPairs(temp, bytes, keys, keys_sorted, rows, perm, (int)m, 0, 32, st));
  return 0;
}

extern "C" int os3d_kernel_map_tiles(const int32_t *nbr, int64_t m, const int32_t *perm, int32_t *nbr_t,
                                     uint32_t *tile_mask, void *stream) {
  if (m < 0) return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  const int64_t n_tiles = cdiv(m, tc::kTileM);
  tc::kernel_map_tiles_kernel<<<(unsigned)n_tiles, 256, 0, (cudaStream_t)stream>>>(nbr, m, n_tiles * tc::kTileM, perm,
                                                                                   nbr_t, tile_mask);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_spconv_bf16_packed_elems(int cin, int cout, int64_t *elems) {
  if (cin <= 0 || cin % 8 || cout <= 0) return OS3D_ERR_BAD_ARG;
  *elems = (int64_t)OS3D_KVOL * cdiv(cin, tc::kBlockK) * cout * tc::kBlockK;
  return 0;
}

extern "C" int os3d_pack_weight_bf16(const float *w_spconv, int cin, int cout, void *w_packed, void *stream) {
  if (cin <= 0 || cout <= 0) return OS3D_ERR_BAD_ARG;
  const int ncb = (int)cdiv(cin, tc::kBlockK);
  const int64_t total = (int64_t)OS3D_KVOL * ncb * cout * tc::kBlockK;
  tc::pack_weight_img_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w_spconv, cin, cout, ncb, OS3D_KVOL, (__nv_bfloat16 *)w_packed);
  OS3D_LAUNCH_CHECK();
  return 0;
}

// Shared launch: geometry of the accumulators / rings for one (m_out, cin, cout) problem.
static int launch_tc(tc::Params &p, cudaStream_t stream) {
  const int cout = p.cout;
  p.n_parts = cout > 256 ? 2 : 1;
  p.n_per_part = cout / p.n_parts;
  p.idesc = ptx::make_idesc_bf16(tc::kTileM, p.n_per_part);
  // Occupancy first: what paces this kernel for small Cout is the serial work of ONE MMA-issuing warp per CTA
  // (barrier round trips, proxy fence, issue), so two or three co-resident CTAs beat one CTA with a deep ring
  // (measured sweep: gpurun_out/sweep_cfg.log).  Per CTA: T accumulators (T * Cout <= 512 / ctas TMEM
  // columns), a weight ring of sb slabs and an A ring of sa (power of two) 16 KB slots within 227 KB / ctas.
  int ctas = cout <= 256 ? 2 : 1;
  { const char *e = getenv("OS3D_SPCONV_CTAS");   // tuning override: co-resident CTAs per SM the geometry is sized for
    if (e && (atoi(e) == 1 || (atoi(e) == 2 && cout <= 256))) ctas = atoi(e); }
  int tiles = (512 / ctas) / cout;
  tiles = tiles > 4 ? 4 : tiles < 1 ? 1 : tiles;      // 4, not 5: an even split between the two issuing warps (48 -> 48: 0.52 -> 0.50 ms)
  while (tiles > 1 && cdiv(p.n_tiles, tiles) < 2 * 148 * ctas) --tiles;
  { const char *e = getenv("OS3D_SPCONV_TILES");   // tuning / test override of the accumulators per CTA
    if (e && atoi(e) > 0) tiles = min(atoi(e), min(512 / cout, tc::kMaxTiles)); }
  p.tiles_per_cta = tiles;
  int cols = 32;
  while (cols < tiles * cout) cols <<= 1;
  p.tmem_cols = cols;
  const int b_bytes = cout * 128;
  const int tail = (2 * tc::kMaxSA + 2 * tc::kMaxSB + 1) * 8 + 16 * 4 + 4 * cout * 4 + 64;   // barriers, misc, epilogue parameters
  const int budget = (227 * 1024) / ctas - 1024 - tail - (ctas > 1 ? 1024 : 0);   // 1 KB per CTA is reserved by the driver
  int sb = 3;
  while (sb > 2 && budget - sb * b_bytes < 2 * tc::kATileBytes) --sb;
  int sa = (budget - sb * b_bytes) / tc::kATileBytes;
  { const char *e = getenv("OS3D_SPCONV_SA"); if (e && atoi(e) > 0 && atoi(e) < sa) sa = atoi(e); }
  int sa_log2 = 3;                                           // ring depth: a power of two (index / phase from one counter)
  while (sa_log2 > 0 && (1 << sa_log2) > sa) --sa_log2;
  sa = 1 << sa_log2;
  if (sa < 2 || (budget - sb * b_bytes) < sa * tc::kATileBytes) return OS3D_ERR_BAD_ARG;
  p.sa = sa;
  p.sa_log2 = sa_log2;
  p.sb = sb;
  // One issuing warp's serial per-slot work (barrier wait, proxy fence, UTCHMMAs, commit: ~500 cycles) paces the layers with
  // several small tiles per CTA; when warps 4-7 do not gather, warp 7 issues the odd tiles.  (OS3D_SPCONV_ISSUERS=1: off)
  p.issuers = (sa <= 4 && tiles >= 2) ? 2 : 1;
  { const char *e = getenv("OS3D_SPCONV_ISSUERS"); if (e && atoi(e) == 1) p.issuers = 1; }
  const int smem = 1024 + sa * tc::kATileBytes + sb * b_bytes + tail;
  // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute: once per device, not once per process
  static bool configured[64] = {false};
  int cfg_dev = 0;
  OS3D_CUDA(cudaGetDevice(&cfg_dev));
  if (cfg_dev < 0 || cfg_dev >= 64 || !configured[cfg_dev]) {
    OS3D_CUDA(cudaFuncSetAttribute(tc::spconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (cfg_dev >= 0 && cfg_dev < 64) configured[cfg_dev] = true;
  }
  tc::spconv_tc_kernel<<<(unsigned)cdiv(p.n_tiles, tiles), tc::kThreads, smem, stream>>>(p);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_spconv_fwd_bf16_ld(const void *in, int64_t m_in, const int32_t *nbr_t, const uint32_t *tile_mask,
                                       const int32_t *perm, int64_t m_out, int cin, int cout, const void *w,
                                       const float *scale, const float *shift, const void *residual, int flags, void *out,
                                       int64_t ldo, void *stream) {
  if (cin <= 0 || cin % 8 || cout < 16 || cout % 16 || cout > 512 || (cout > 256 && cout % 32) ||
      ((scale == nullptr) != (shift == nullptr)) || m_in <= 0 || ((uintptr_t)in & 15) || ldo < cout || ldo % 8 ||
      ((uintptr_t)out & 15))
    return OS3D_ERR_BAD_ARG;
  if (m_out == 0) return 0;
  tc::Params p;
  p.in = (const __nv_bfloat16 *)in;
  p.cin = cin;
  p.nbr_t = nbr_t;
  p.tile_mask = tile_mask;
  p.perm = perm;
  p.m_out = m_out;
  p.n_tiles = (int)cdiv(m_out, tc::kTileM);
  p.m_pad = (int64_t)p.n_tiles * tc::kTileM;
  p.cin16 = (cin + 15) / 16 * 16;
  p.cout = cout;
  p.ncb = (int)cdiv(cin, tc::kBlockK);
  p.w_img = (const __nv_bfloat16 *)w;
  p.scale = scale;
  p.shift = shift;
  p.residual = (const __nv_bfloat16 *)residual;
  p.flags = flags;
  p.out = (__nv_bfloat16 *)out;
  p.ldo = ldo;
  p.dense = 0;
  p.ln_gamma = p.ln_beta = nullptr;
  p.ln_eps = 0.0f;
  p.table = nullptr;
  p.tab_idx = nullptr;
  p.tab_cols = 0;
  return launch_tc(p, (cudaStream_t)stream);
}

extern "C" int os3d_spconv_fwd_bf16(const void *in, int64_t m_in, const int32_t *nbr_t, const uint32_t *tile_mask,
                                    const int32_t *perm, int64_t m_out, int cin, int cout, const void *w, const float *scale,
                                    const float *shift, const void *residual, int flags, void *out, void *stream) {
  return os3d_spconv_fwd_bf16_ld(in, m_in, nbr_t, tile_mask, perm, m_out, cin, cout, w, scale, shift, residual, flags, out,
                                 cout, stream);
}

extern "C" int os3d_linear_bf16_packed_elems(int k, int n, int64_t *elems) {
  if (k <= 0 || k % 8 || n <= 0) return OS3D_ERR_BAD_ARG;
  *elems = cdiv(k, tc::kBlockK) * n * tc::kBlockK;
  return 0;
}

extern "C" int os3d_pack_linear_bf16(const float *w, int k, int n, void *w_packed, void *stream) {
  if (k <= 0 || n <= 0) return OS3D_ERR_BAD_ARG;
  const int ncb = (int)cdiv(k, tc::kBlockK);
  const int64_t total = (int64_t)ncb * n * tc::kBlockK;
  tc::pack_weight_img_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, k, n, ncb, 1,
                                                                                         (__nv_bfloat16 *)w_packed);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_linear_bf16(const void *x, int64_t m, int k, int n, const void *w, const float *bias, int flags,
                                const void *residual, const float *ln_gamma, const float *ln_beta, float ln_eps,
                                const void *table, const int32_t *tab_idx, int tab_cols, void *out, int64_t ldo,
                                void *stream) {
  if (k <= 0 || k % 8 || n < 16 || n % 16 || n > 512 || (n > 256 && n % 32) || m < 0 || ((uintptr_t)x & 15) ||
      (flags & ~(1 | 4 | 8 | 16)) || ((flags & 8) && (!ln_gamma || !ln_beta || (flags & 16))) ||
      ((flags & 16) && (!table || !tab_idx || tab_cols % 16 || tab_cols > n)) || ldo < n || ldo % 8 ||
      ((uintptr_t)out & 15))
    return OS3D_ERR_BAD_ARG;
  if (m == 0) return 0;
  tc::Params p;
  p.in = (const __nv_bfloat16 *)x;
  p.cin = k;
  p.nbr_t = nullptr;
  p.tile_mask = nullptr;
  p.perm = nullptr;
  p.m_out = m;
  p.n_tiles = (int)cdiv(m, tc::kTileM);
  p.m_pad = (int64_t)p.n_tiles * tc::kTileM;
  p.cin16 = (k + 15) / 16 * 16;
  p.cout = n;
  p.ncb = (int)cdiv(k, tc::kBlockK);
  p.w_img = (const __nv_bfloat16 *)w;
  p.scale = nullptr;
  p.shift = bias;
  p.residual = (const __nv_bfloat16 *)residual;
  p.flags = flags;
  p.out = (__nv_bfloat16 *)out;
  p.ldo = ldo;
  p.dense = 1;
  p.ln_gamma = ln_gamma;
  p.ln_beta = ln_beta;
  p.ln_eps = ln_eps;
  p.table = (const __nv_bfloat16 *)table;
  p.tab_idx = tab_idx;
  p.tab_cols = tab_cols;
  return launch_tc(p, (cudaStream_t)stream);
}
