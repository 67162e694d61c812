// window.cu -- stage 4, integer part: sparse window partition without host synchronisation.
// replaces get_window_coors (seg3d/utils/swformer_utils.py:109-154), get_inner_win_inds
// (seg3d/ops/ingroup_inds/src/ingroup_inds_cuda.cu:12-52: one D2H sync + cudaMalloc/cudaFree per call),
// batching_single_shift (seg3d/models/layers/point_transformer_layer.py:71-87) and make_continuous_inds /
// get_flat2win_inds (swformer_utils.py:8-31,158-171: unique + sort + two .item() syncs per level).
//
// Window ids live in a small dense range (batch * nwin_x*nwin_y*nwin_z, 1.5 M for 8 Cartesian frames at level 1), so
// the partition is a counting sort over that range:
//   1. assign : window id + in-window coords per voxel, occupancy histogram (RED.ADD)
//   2. count  : per-block sums of 5 lanes over the dense windows {tokens, non-empty windows of level 0..3}
//   3. scan   : single-block scan of the block sums
//   4. apply  : token offset of each window, its level and its rank among that level's windows -> segment table
//   5. fill   : voxel rows into their window's segment (arrival order)
//   6. sort   : each segment ascending by voxel row (warp per window) -> deterministic, stable in-window rank
// Every pass is a coalesced stream over voxels or dense windows; nothing returns to the host.
#include "common.cuh"

namespace os3d {

constexpr int kLanes = 5;  // tokens, windows@level0..3

__device__ __forceinline__ int level_of(const os3d_window_cfg_t &cfg, int cnt) {
  int l = -1;
  for (int i = 0; i < cfg.n_levels; ++i)
    if (cnt >= cfg.lvl_lo[i] && cnt < cfg.lvl_hi[i]) l = i;
  return l;
}

__global__ void win_assign_kernel(const int4 *__restrict__ idx, int64_t m, os3d_window_cfg_t cfg,
                                  int64_t *__restrict__ win_id, int32_t *__restrict__ in_win,
                                  int32_t *__restrict__ win_count, int32_t *__restrict__ pos_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int4 c = __ldg(idx + i);  // (b, z, y, x)
  const int sx = c.w + cfg.shift_x, sy = c.z + cfg.shift_y, sz = c.y + cfg.shift_z;
  const int wx = sx / cfg.win_x, wy = sy / cfg.win_y, wz = sz / cfg.win_z;
  const int64_t per_sample = (int64_t)cfg.nwin_x * cfg.nwin_y * cfg.nwin_z;
  const int64_t w = c.x * per_sample + (int64_t)wx * cfg.nwin_y * cfg.nwin_z + (int64_t)wy * cfg.nwin_z + wz;
  win_id[i] = w;
  in_win[i * 3 + 0] = sz - wz * cfg.win_z;
  in_win[i * 3 + 1] = sy - wy * cfg.win_y;
  in_win[i * 3 + 2] = sx - wx * cfg.win_x;
  // row of the window's position-embedding table: (z * win_y + y) * win_x + x
  if (pos_idx) pos_idx[i] = ((sz - wz * cfg.win_z) * cfg.win_y + (sy - wy * cfg.win_y)) * cfg.win_x + (sx - wx * cfg.win_x);
  atomicAdd(win_count + w, 1);
}

__global__ void group_hist_kernel(const int64_t *__restrict__ group, int64_t n, int32_t *__restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(count + __ldg(group + i), 1);
}

__device__ __forceinline__ void lanes_of(const os3d_window_cfg_t &cfg, int cnt, int v[kLanes]) {
  const int l = cnt > 0 ? level_of(cfg, cnt) : -1;
  v[0] = l >= 0 ? cnt : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) v[1 + i] = (l == i);
}

__global__ void __launch_bounds__(kScanThreads) win_count_kernel(const int32_t *__restrict__ win_count, int64_t n_win,
                                                                  os3d_window_cfg_t cfg,
                                                                  int32_t *__restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int s[kLanes] = {0, 0, 0, 0, 0};
  for (int t = 0; t < kScanItems; ++t) {
    if (base + t >= n_win) break;
    int v[kLanes];
    lanes_of(cfg, __ldg(win_count + base + t), v);
#pragma unroll
    for (int l = 0; l < kLanes; ++l) s[l] += v[l];
  }
#pragma unroll
  for (int l = 0; l < kLanes; ++l) {
    int total;
    block_excl_scan_256(s[l], &total);
    if (threadIdx.x == 0) block_sums[(int64_t)blockIdx.x * kLanes + l] = total;
  }
}

// win_meta[w] = {token offset, level, slot in the level-major segment table}; resets win_count to 0 (fill cursor).
__global__ void __launch_bounds__(kScanThreads) win_apply_kernel(int32_t *__restrict__ win_count, int64_t n_win,
                                                                  os3d_window_cfg_t cfg,
                                                                  const int32_t *__restrict__ block_sums,
                                                                  int64_t n_blocks, int32_t *__restrict__ win_meta,
                                                                  int32_t *__restrict__ seg_start,
                                                                  int32_t *__restrict__ seg_len,
                                                                  int32_t *__restrict__ level_info, int64_t m) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int cnt[kScanItems], v[kScanItems][kLanes], s[kLanes] = {0, 0, 0, 0, 0}, ex[kLanes];
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    cnt[t] = base + t < n_win ? win_count[base + t] : 0;
    lanes_of(cfg, cnt[t], v[t]);
#pragma unroll
    for (int l = 0; l < kLanes; ++l) s[l] += v[t][l];
  }
#pragma unroll
  for (int l = 0; l < kLanes; ++l) {
    int total;
    ex[l] = block_excl_scan_256(s[l], &total) + block_sums[(int64_t)blockIdx.x * kLanes + l];
  }
  const int32_t *tot = block_sums + n_blocks * kLanes;  // grand totals per lane
  int first[4];
  first[0] = 0;
#pragma unroll
  for (int l = 1; l < 4; ++l) first[l] = first[l - 1] + tot[l];  // tot[1+l-1]
#pragma unroll
  for (int t = 0; t < kScanItems; ++t) {
    if (base + t >= n_win) break;
    const int64_t w = base + t;
    int lvl = -1;
#pragma unroll
    for (int l = 0; l < 4; ++l) if (v[t][1 + l]) lvl = l;
    if (lvl >= 0) {
      const int slot = first[lvl] + ex[1 + lvl];
      win_meta[w * 3 + 0] = ex[0];
      win_meta[w * 3 + 1] = lvl;
      win_meta[w * 3 + 2] = slot;
      seg_start[slot] = ex[0];
      seg_len[slot] = cnt[t];
    } else {
      win_meta[w * 3 + 0] = -1;  // empty, or occupancy outside every batching range (tokens dropped)
      win_meta[w * 3 + 1] = -1;
      win_meta[w * 3 + 2] = -1;
    }
    if (cnt[t]) win_count[w] = 0;
#pragma unroll
    for (int l = 0; l < kLanes; ++l) ex[l] += v[t][l];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int n_all = 0;
    for (int l = 0; l < 4; ++l) {
      level_info[l] = tot[1 + l];
      level_info[4 + l] = first[l];
      n_all += tot[1 + l];
    }
    level_info[12] = (int32_t)(m - tot[0]);  // tokens in windows outside every batching range
    level_info[13] = n_all;
    level_info[14] = tot[0];
    level_info[15] = 0;
    seg_start[n_all] = tot[0];  // sentinel: seg r spans [seg_start[r], seg_start[r+1]) only inside a level; see below
  }
}

__global__ void win_fill_kernel(const int64_t *__restrict__ win_id, int64_t m, const int32_t *__restrict__ win_meta,
                                int32_t *__restrict__ cursor, int32_t *__restrict__ order) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t w = __ldg(win_id + i);
  const int32_t off = __ldg(win_meta + w * 3);
  if (off < 0) return;
  order[off + atomicAdd(cursor + w, 1)] = (int32_t)i;
}

// One warp per NON-EMPTY window (segment slot), warps striding over the level-major segment table: sort the segment
// ascending by voxel row (rank counting in shared memory), write the per-voxel level / window rank / in-window rank and the
// number of over-capacity tokens.  (The first version launched one warp per grid CELL -- 1.3 M windows at level 1 of an
// 8-frame batch, 17 k of them occupied: 0.22 ms of empty warps per call.)
constexpr int kSortWarps = 8;
constexpr int kMaxSeg = 1024;  // >= any max_tokens the reference uses (800)

__global__ void __launch_bounds__(kSortWarps * 32) win_sort_kernel(const int32_t *__restrict__ seg_start,
                                                                   const int32_t *__restrict__ seg_len,
                                                                   os3d_window_cfg_t cfg, int32_t *__restrict__ order,
                                                                   int32_t *__restrict__ level,
                                                                   int32_t *__restrict__ win_rank,
                                                                   int32_t *__restrict__ inner,
                                                                   int2 *__restrict__ pos_seg,
                                                                   int32_t *__restrict__ level_info) {
  __shared__ int32_t buf[kSortWarps][kMaxSeg];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_slots = level_info[13];
  int first[4], has[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) { first[l] = level_info[4 + l]; has[l] = level_info[l] > 0; }
  int dropped = 0;
  for (int slot = blockIdx.x * kSortWarps + wid; slot < n_slots; slot += gridDim.x * kSortWarps) {
  int lvl = 0;
#pragma unroll
  for (int l = 1; l < 4; ++l) if (has[l] && slot >= first[l]) lvl = l;
  const int first_of_level = first[lvl];
  const int32_t off = __ldg(seg_start + slot);
  const int n = __ldg(seg_len + slot);
  int32_t *seg = order + off;
  for (int t = lane; t < n; t += 32) pos_seg[off + t] = make_int2(off, n);  // (window start, length) per position
  __syncwarp();
  const int cap = cfg.lvl_tokens[lvl];
  if (n <= 64) {
    for (int t = lane; t < n; t += 32) buf[wid][t] = seg[t];
    __syncwarp();
    for (int t = lane; t < n; t += 32) {
      const int32_t mine = buf[wid][t];
      int r = 0;
      for (int j = 0; j < n; ++j) r += buf[wid][j] < mine;
      seg[r] = mine;
      level[mine] = lvl;
      win_rank[mine] = slot - first_of_level;
      inner[mine] = r;
      dropped += r >= cap;
    }
  } else if (n <= kMaxSeg) {
    // larger windows: bitonic sort of the (unique) voxel rows in shared memory -- O(n log^2 n / 32) steps for the warp
    // instead of the O(n^2 / 32) of rank counting, which made the few 500-800 token windows of the coarse levels the
    // critical path of the whole launch (one warp, 20k iterations)
    int pow2 = 128;
    while (pow2 < n) pow2 <<= 1;
    int32_t *b = buf[wid];
    for (int t = lane; t < pow2; t += 32) b[t] = t < n ? seg[t] : 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= pow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < (pow2 >> 1); i += 32) {
          const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
          const int32_t x = b[lo], y = b[hi];
          const bool up = (lo & k) == 0;
          if ((x > y) == up) { b[lo] = y; b[hi] = x; }
        }
        __syncwarp();
      }
    }
    for (int t = lane; t < n; t += 32) {
      const int32_t mine = b[t];
      seg[t] = mine;
      level[mine] = lvl;
      win_rank[mine] = slot - first_of_level;
      inner[mine] = t;
      dropped += t >= cap;
    }
  } else {  // longer than the staging buffer (never for windows: max_tokens <= 800): rank straight from global
            // memory; the segment itself stays in arrival order (attention is order independent)
    for (int t = lane; t < n; t += 32) {
      const int32_t mine = seg[t];
      int r = 0;
      for (int j = 0; j < n; ++j) r += seg[j] < mine;
      level[mine] = lvl; win_rank[mine] = slot - first_of_level; inner[mine] = r;
      dropped += r >= cap;
    }
  }
  __syncwarp();               // the staging buffer is reused by the next segment of this warp
  }
  dropped = __reduce_add_sync(0xffffffffu, dropped);
  if (lane == 0 && dropped) atomicAdd(level_info + 15, dropped);
}

__global__ void mark_unassigned_kernel(const int64_t *__restrict__ win_id, int64_t m, const int32_t *__restrict__ win_meta,
                                       int32_t *__restrict__ level, int32_t *__restrict__ win_rank,
                                       int32_t *__restrict__ inner) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  if (__ldg(win_meta + __ldg(win_id + i) * 3) < 0) { level[i] = -1; win_rank[i] = -1; inner[i] = -1; }
}

// ---- positional embedding ----------------------------------------------------------------------------
// get_pos_embed, point_transformer_layer.py:152-207 (3-D windows, normalize_pos=False):
//   channel j of axis a (a: 0 = x, 1 = y, 2 = z; pos_length = c/3 channels each):
//     angle = (coord_a - win_a/2) / temperature^(2*floor(j/2)/pos_length);  even j -> sin, odd j -> cos
template <typename T>
__global__ void __launch_bounds__(256) pos_embed_kernel(const int32_t *__restrict__ in_win, int64_t m, int c, int pos_len,
                                                         float hx, float hy, float hz, float temperature,
                                                         T *__restrict__ out) {
  // one thread per (row, channel pair): channels (2q, 2q+1) of an axis share the angle -> sin, cos.
  // inv_freq depends only on the pair index inside an axis: a shared table, filled once per block.
  __shared__ float inv_freq[256];
  const int half = (pos_len + 1) / 2;
  for (int q = threadIdx.x; q < half; q += blockDim.x)
    inv_freq[q] = powf(temperature, (float)(2 * q) / (float)pos_len);
  __syncthreads();
  const int pairs_per_row = (c + 1) / 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * pairs_per_row) return;
  const int64_t i = t / pairs_per_row;
  const int ch0 = (int)(t - i * pairs_per_row) * 2;
  float v[2] = {0.0f, 0.0f};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int ch = ch0 + e;
    const int a = ch / pos_len;
    if (ch < c && a < 3) {
      const int j = ch - a * pos_len;
      // in_win is (z, y, x); the embedding's axis order is x, y, z
      const float coord = (float)__ldg(in_win + i * 3 + (2 - a)) - (a == 0 ? hx : (a == 1 ? hy : hz));
      const float ang = coord / inv_freq[j >> 1];
      v[e] = (j & 1) ? cosf(ang) : sinf(ang);
    }
  }
  if constexpr (sizeof(T) == 4) {
    out[i * c + ch0] = v[0];
    if (ch0 + 1 < c) out[i * c + ch0 + 1] = v[1];
  } else {
    if (!(c & 1)) {
      *reinterpret_cast<__nv_bfloat162 *>(out + i * c + ch0) = __floats2bfloat162_rn(v[0], v[1]);
    } else {
      out[i * c + ch0] = __float2bfloat16(v[0]);
      if (ch0 + 1 < c) out[i * c + ch0 + 1] = __float2bfloat16(v[1]);
    }
  }
}

}  // namespace os3d

using namespace os3d;

static int partition_common(const int64_t *win_id, int64_t m, int64_t n_win, const os3d_window_cfg_t &cfg,
                            int32_t *win_count, int32_t *win_meta, int32_t *block_sums, int64_t n_blocks,
                            int32_t *level, int32_t *win_rank, int32_t *inner, int32_t *order, int32_t *seg_start,
                            int32_t *seg_len, int32_t *pos_seg, int32_t *level_info, cudaStream_t st) {
  const unsigned gv = (unsigned)cdiv(m, 256);
  win_count_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(win_count, n_win, cfg, block_sums);
  scan_block_sums_multi_kernel<<<1, kScanThreads, 0, st>>>(block_sums, n_blocks, kLanes);
  win_apply_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(win_count, n_win, cfg, block_sums, n_blocks, win_meta,
                                                                seg_start, seg_len, level_info, m);
  win_fill_kernel<<<gv, 256, 0, st>>>(win_id, m, win_meta, win_count, order);
  mark_unassigned_kernel<<<gv, 256, 0, st>>>(win_id, m, win_meta, level, win_rank, inner);
  // at most one segment per voxel; 148 SMs x 8 blocks of 8 warps stride over the occupied ones
  const unsigned sort_blocks = (unsigned)std::min<int64_t>(cdiv(std::min<int64_t>(n_win, m), kSortWarps), 148 * 8);
  win_sort_kernel<<<sort_blocks, kSortWarps * 32, 0, st>>>(seg_start, seg_len, cfg, order, level, win_rank, inner,
                                                           (int2 *)pos_seg, level_info);
  OS3D_LAUNCH_CHECK();
  return 0;
}

extern "C" int os3d_window_partition(const int32_t *idx, int64_t m, int batch, const os3d_window_cfg_t *cfg_p,
                                     int32_t *win_count, int32_t *win_meta, int32_t *block_sums, int64_t n_blocks,
                                     int64_t *win_id, int32_t *in_win, int32_t *level, int32_t *win_rank,
                                     int32_t *inner, int32_t *order, int32_t *seg_start, int32_t *seg_len,
                                     int32_t *pos_seg, int32_t *level_info, int32_t *pos_idx, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const os3d_window_cfg_t cfg = *cfg_p;
  const int64_t n_win = (int64_t)batch * cfg.nwin_x * cfg.nwin_y * cfg.nwin_z;
  if (cfg.n_levels < 1 || cfg.n_levels > OS3D_MAX_LEVELS || n_blocks != cdiv(n_win, kScanTile)) return OS3D_ERR_BAD_ARG;
  OS3D_CUDA(cudaMemsetAsync(win_count, 0, sizeof(int32_t) * (size_t)n_win, st));
  OS3D_CUDA(cudaMemsetAsync(level_info, 0, sizeof(int32_t) * 16, st));
  if (m == 0) return 0;
  win_assign_kernel<<<(unsigned)cdiv(m, 256), 256, 0, st>>>((const int4 *)idx, m, cfg, win_id, in_win, win_count, pos_idx);
  return partition_common(win_id, m, n_win, cfg, win_count, win_meta, block_sums, n_blocks, level, win_rank, inner, order,
                          seg_start, seg_len, pos_seg, level_info, st);
}

extern "C" int os3d_group_partition(const int64_t *group, int64_t n, int64_t n_groups, const os3d_window_cfg_t *cfg_p,
                                    int32_t *count, int32_t *meta, int32_t *block_sums, int64_t n_blocks,
                                    int32_t *level, int32_t *group_rank, int32_t *inner, int32_t *order,
                                    int32_t *seg_start, int32_t *seg_len, int32_t *pos_seg, int32_t *level_info,
                                    void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const os3d_window_cfg_t cfg = *cfg_p;
  if (cfg.n_levels < 1 || cfg.n_levels > OS3D_MAX_LEVELS || n_blocks != cdiv(n_groups, kScanTile)) return OS3D_ERR_BAD_ARG;
  OS3D_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)n_groups, st));
  OS3D_CUDA(cudaMemsetAsync(level_info, 0, sizeof(int32_t) * 16, st));
  if (n == 0) return 0;
  group_hist_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(group, n, count);
  return partition_common(group, n, n_groups, cfg, count, meta, block_sums, n_blocks, level, group_rank, inner, order,
                          seg_start, seg_len, pos_seg, level_info, st);
}

extern "C" int os3d_pos_embed(const int32_t *in_win, int64_t m, int c, int win_x, int win_y, int win_z,
                              float temperature, int elem_size, void *out, void *stream) {
  if (m == 0) return 0;
  const int pos_len = c / 3;
  if (pos_len <= 0) return OS3D_ERR_BAD_ARG;
  if (pos_len > 512) return OS3D_ERR_BAD_ARG;
  const unsigned g = (unsigned)cdiv(m * ((c + 1) / 2), 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_size == 4)
    pos_embed_kernel<float><<<g, 256, 0, st>>>(in_win, m, c, pos_len, win_x / 2.0f, win_y / 2.0f, win_z / 2.0f,
                                               temperature, (float *)out);
  else if (elem_size == 2)
    pos_embed_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(in_win, m, c, pos_len, win_x / 2.0f, win_y / 2.0f, win_z / 2.0f,
                                                       temperature, (__nv_bfloat16 *)out);
  else
    return OS3D_ERR_BAD_ARG;
  OS3D_LAUNCH_CHECK();
  return 0;
}
