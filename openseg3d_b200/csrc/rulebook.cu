// rulebook.cu -- stage 2: kernel maps for SubMConv3d / SparseConv3d / SparseInverseConv3d (3x3x3), integer only.
// The reference gets these from spconv (seg3d/utils/spconv_utils.py:16-22; call sites
// seg3d/models/backbones/pointtransformer.py:26,31,133,159-166); semantics per SURVEY.md Appendix A.
//
// Design (B200-first, not spconv's): every conv is OUTPUT-STATIONARY -- the map is a dense [M_out, 27] table of input
// rows (-1 = no neighbour) so the convolution kernel needs no scatter atomics and one table serves every conv that
// shares an indice_key.  Sites live in a 16-byte-slot open-addressing table (one 128-bit load per probe, the whole
// table is L2 resident: 2M slots = 32 MB for a batch of 8 frames).  Strided output sites are produced already sorted
// (ascending linear index) by marking a bitmap over the dense output grid and ranking set bits with a popcount
// prefix -- no sort, no dedupe pass; the same (bitmap, prefix) pair is an O(1) site->row lookup.
#include "common.cuh"

namespace os3d {

__device__ __forceinline__ int64_t lin4(int b, int z, int y, int x, int sz, int sy, int sx) {
  return (((int64_t)b * sz + z) * sy + y) * (int64_t)sx + x;
}

__global__ void hash_build_kernel(const int4 *__restrict__ idx, int64_t m, int sz, int sy, int sx, os3d_slot_t *table,
                                  uint64_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = __ldg(idx + i);  // (b, z, y, x)
  const int64_t s = table_insert(table, mask, lin4(c.x, c.y, c.z, c.w, sz, sy, sx));
  table[s].val = (int32_t)i;  // sites are unique, so exactly one writer per slot
}

// The submanifold map is symmetric: nbr[i, k] = j  <=>  nbr[j, 26 - k] = i.  One thread per (row, k <= 13) probes the
// site table (half the random L2 accesses of probing all 27 offsets -- the probes are what this kernel costs) and
// writes both entries; the table is pre-filled with -1.  The 4-int coordinate load is a broadcast inside the 14
// consecutive threads of a row.
constexpr int kSubmHalf = 14;       // offsets 0..13 (13 = centre)
__global__ void subm_table_kernel(const int4 *__restrict__ idx, int64_t m, int sz, int sy, int sx,
                                  const os3d_slot_t *__restrict__ table, uint64_t mask, int32_t *__restrict__ nbr,
                                  int32_t *__restrict__ pair_count) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int found = 0;
  if (t < m * kSubmHalf) {
    const int64_t i = t / kSubmHalf;
    const int k = (int)(t - i * kSubmHalf);
    const int4 c = __ldg(idx + i);
    if (k == 13) {
      nbr[i * OS3D_KVOL + 13] = (int32_t)i;  // centre offset is the identity map
      found = 1;
    } else {
      const int z = c.y + k / 9 - 1, y = c.z + (k / 3) % 3 - 1, x = c.w + k % 3 - 1;
      const int32_t j = (z >= 0 && z < sz && y >= 0 && y < sy && x >= 0 && x < sx)
                            ? table_find(table, mask, lin4(c.x, z, y, x, sz, sy, sx)) : -1;
      if (j >= 0) {
        nbr[i * OS3D_KVOL + k] = j;
        nbr[(int64_t)j * OS3D_KVOL + (26 - k)] = (int32_t)i;
        found = 2;
      }
    }
  }
  // pair count (diagnostic): one atomic per CTA
  const int both = __syncthreads_count(found == 2), one = __syncthreads_count(found == 1);
  if (threadIdx.x == 0 && (both | one)) atomicAdd(pair_count, 2 * both + one);
}

// ---- strided conv: output sites --------------------------------------------------------------------
// k = 3, stride 2, pad 1:  o = (i + 1 - k) / 2 for the k in {0,1,2} that make it an integer in range.
// Per dim an input coordinate reaches 1 (even i: k = 1) or 2 (odd i: k = 0, 2) outputs.
__device__ __forceinline__ int strided_outs(int i, int osz, int *o, int *k) {
  int n = 0;
  if (i & 1) {
    const int a = (i + 1) >> 1;  // k = 0
    if (a < osz) { o[n] = a; k[n] = 0; ++n; }
    const int b = (i - 1) >> 1;  // k = 2
    if (b < osz) { o[n] = b; k[n] = 2; ++n; }
  } else {
    const int a = i >> 1;  // k = 1
    if (a < osz) { o[n] = a; k[n] = 1; ++n; }
  }
  return n;
}

__global__ void strided_mark_kernel(const int4 *__restrict__ idx, int64_t m, int oz, int oy, int ox,
                                    uint32_t *__restrict__ bitmap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = __ldg(idx + i);
  int zo[2], yo[2], xo[2], kk[2];
  const int nz = strided_outs(c.y, oz, zo, kk), ny = strided_outs(c.z, oy, yo, kk), nx = strided_outs(c.w, ox, xo, kk);
  for (int a = 0; a < nz; ++a)
    for (int b = 0; b < ny; ++b)
      for (int d = 0; d < nx; ++d) {
        const int64_t l = lin4(c.x, zo[a], yo[b], xo[d], oz, oy, ox);
        atomicOr(bitmap + (l >> 5), 1u << (l & 31));
      }
}

__global__ void __launch_bounds__(kScanThreads) bitmap_count_kernel(const uint32_t *__restrict__ bitmap, int64_t n_words,
                                                                     int32_t *__restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int cnt = 0;
  if (base + kScanItems <= n_words) {
    const uint4 w = __ldg(reinterpret_cast<const uint4 *>(bitmap + base));
    cnt = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
  } else {
    for (int t = 0; t < kScanItems; ++t) if (base + t < n_words) cnt += __popc(bitmap[base + t]);
  }
  int total;
  block_excl_scan_256(cnt, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) bitmap_expand_kernel(const uint32_t *__restrict__ bitmap, int64_t n_words,
                                                                      const int32_t *__restrict__ block_sums,
                                                                      int64_t n_blocks, int oz, int oy, int ox,
                                                                      int32_t *__restrict__ word_prefix,
                                                                      int32_t *__restrict__ out_idx, int64_t cap_out,
                                                                      int32_t *__restrict__ num_out) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t w[kScanItems];
  int cnt = 0;
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    w[t] = base + t < n_words ? __ldg(bitmap + base + t) : 0u;
    cnt += __popc(w[t]);
  }
  int total;
  int ex = block_excl_scan_256(cnt, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    if (base + t >= n_words) break;
    word_prefix[base + t] = ex;
    uint32_t bits = w[t];
    while (bits) {
      const int bpos = __ffs(bits) - 1;
      bits &= bits - 1;
      int64_t l = ((base + t) << 5) + bpos;
      const int x = (int)(l % ox); l /= ox;
      const int y = (int)(l % oy); l /= oy;
      const int z = (int)(l % oz); l /= oz;
      if (ex < cap_out) reinterpret_cast<int4 *>(out_idx)[ex] = make_int4((int)l, z, y, x);
      ++ex;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *num_out = block_sums[n_blocks];
}

// Both tables of a strided conv from the input side, one thread per (input row i, k): the output site is
// o = (i + 1 - k) / 2 when integral and in range, its row the rank of its bit in the bitmap (an O(1), cache-local
// lookup -- no hash probes).  The pair (k: i -> o) is written to inv_nbr[i, k] (coalesced) and to fwd_nbr[o, k]
// (scattered; fwd_nbr is pre-filled with -1).  Probing the input hash table from the output side, as the first
// version did, cost 27 random L2 accesses per output row.
__global__ void strided_pairs_kernel(const int4 *__restrict__ idx, int64_t m, int oz, int oy, int ox,
                                     const uint32_t *__restrict__ bitmap, const int32_t *__restrict__ word_prefix,
                                     int32_t *__restrict__ inv_nbr, int32_t *__restrict__ fwd_nbr,
                                     int32_t *__restrict__ pair_count) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int found = 0;
  if (t < m * OS3D_KVOL) {
  const int64_t i = t / OS3D_KVOL;
  const int k = (int)(t - i * OS3D_KVOL);
  const int4 c = __ldg(idx + i);
  const int tz = c.y + 1 - k / 9, ty = c.z + 1 - (k / 3) % 3, tx = c.w + 1 - k % 3;
  int32_t r = -1;
  if (tz >= 0 && ty >= 0 && tx >= 0 && !((tz | ty | tx) & 1)) {
    const int z = tz >> 1, y = ty >> 1, x = tx >> 1;
    if (z < oz && y < oy && x < ox) {
      const int64_t l = lin4(c.x, z, y, x, oz, oy, ox);
      const uint32_t w = __ldg(bitmap + (l >> 5));
      const uint32_t bit = 1u << (l & 31);
      if (w & bit) r = __ldg(word_prefix + (l >> 5)) + __popc(w & (bit - 1));
    }
  }
  inv_nbr[t] = r;
  if (r >= 0) {
    fwd_nbr[(int64_t)r * OS3D_KVOL + k] = (int32_t)i;
    found = 1;
  }
  }
  // one atomic per CTA: a same-address atomic per warp (a million of them) costs more than the lookups
  const int cta_pairs = __syncthreads_count(found);
  if (threadIdx.x == 0 && cta_pairs) atomicAdd(pair_count, cta_pairs);
}

}  // namespace os3d

using namespace os3d;

extern "C" int os3d_hash_build(const int32_t *idx, int64_t m, int sz, int sy, int sx, os3d_slot_t *table, int64_t cap,
                               void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if ((cap & (cap - 1)) || cap < 2 * m || cap < 2) return OS3D_ERR_BAD_ARG;
  OS3D_CUDA(cudaMemsetAsync(table, 0xff, sizeof(os3d_slot_t) * (size_t)cap, st));
  if (m > 0)
    hash_build_kernel<<<(unsigned)cdiv(m, 256), 256, 0, st>>>((const int4 *)idx, m, sz, sy, sx, table, (uint64_t)cap - 1);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_subm_table(const int32_t *idx, int64_t m, int sz, int sy, int sx, const os3d_slot_t *table,
                               int64_t cap, int32_t *nbr, int32_t *pair_count, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OS3D_CUDA(cudaMemsetAsync(pair_count, 0, sizeof(int32_t), st));
  if (m > 0) {
    OS3D_CUDA(cudaMemsetAsync(nbr, 0xff, (size_t)m * OS3D_KVOL * sizeof(int32_t), st));      // -1 = no neighbour
    subm_table_kernel<<<(unsigned)cdiv(m * kSubmHalf, 256), 256, 0, st>>>((const int4 *)idx, m, sz, sy, sx, table,
                                                                         (uint64_t)cap - 1, nbr, pair_count);
  }
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_strided_sites(const int32_t *idx, int64_t m, int batch, int oz, int oy, int ox, uint32_t *bitmap,
                                  int64_t n_words, int32_t *word_prefix, int32_t *block_sums, int64_t n_blocks,
                                  int32_t *out_idx, int64_t cap_out, int32_t *num_out, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // n_words = ceil(cells / 32) rounded up to a multiple of 4 (16-byte loads)
  if (n_words != cdiv(cdiv((int64_t)batch * oz * oy * ox, 32), 4) * 4 || n_blocks != cdiv(n_words, kScanTile))
    return OS3D_ERR_BAD_ARG;
  OS3D_CUDA(cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * (size_t)n_words, st));
  if (m > 0) strided_mark_kernel<<<(unsigned)cdiv(m, 256), 256, 0, st>>>((const int4 *)idx, m, oz, oy, ox, bitmap);
  bitmap_count_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(bitmap, n_words, block_sums);
  scan_block_sums_kernel<<<1, kScanThreads, 0, st>>>(block_sums, n_blocks);
  bitmap_expand_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(bitmap, n_words, block_sums, n_blocks, oz, oy, ox,
                                                                    word_prefix, out_idx, cap_out, num_out);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_strided_tables(const int32_t *idx, int64_t m, int sz, int sy, int sx, const os3d_slot_t *table,
                                   int64_t cap, const int32_t *out_idx, int64_t m_out, int oz, int oy, int ox,
                                   const uint32_t *bitmap, const int32_t *word_prefix, int32_t *fwd_nbr,
                                   int32_t *inv_nbr, int32_t *pair_count, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OS3D_CUDA(cudaMemsetAsync(pair_count, 0, sizeof(int32_t), st));
  if (m_out > 0) OS3D_CUDA(cudaMemsetAsync(fwd_nbr, 0xff, (size_t)m_out * OS3D_KVOL * sizeof(int32_t), st));
  if (m > 0)
    strided_pairs_kernel<<<(unsigned)cdiv(m * OS3D_KVOL, 256), 256, 0, st>>>((const int4 *)idx, m, oz, oy, ox, bitmap,
                                                                            word_prefix, inv_nbr, fwd_nbr, pair_count);
  OS3D_LAUNCH_CHECK();
  return 0;
}
