/*
 * os3d.h -- C ABI of libos3d.so, the B200 (sm_100a) implementation of OpenSeg3D's voxel-backbone hot path.
 *
 * This is the drop-in boundary: plain device pointers, sizes and a CUDA stream; no torch types.  The Python
 * package openseg3d_b200 binds these with ctypes and mirrors the reference's module / operator API on top
 * (INTEGRATION.md shows the binding a reference maintainer would add).  Every entry point
 *   - is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates, never synchronises;
 *   - takes caller-owned scratch (sizes documented per call; the Python side gets it from torch's
 *     caching allocator);
 *   - returns 0 on success or a cudaError_t / negative os3d error code (os3d_error_string()).
 *
 * "replaces:" lines cite the reference interface (paths relative to the OpenSeg3D repository) that each
 * entry point stands in for.
 */
#ifndef OS3D_H_
#define OS3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OS3D_ERR_BAD_ARG (-2)
#define OS3D_KVOL 27 /* 3x3x3 kernel offsets, k = (kz*3+ky)*3+kx */

/* One 16-byte open-addressing slot: linear voxel index -> row. key == -1 means empty. */
typedef struct { int64_t key; int32_t val; int32_t aux; } os3d_slot_t;

const char *os3d_error_string(int code);
int os3d_version(void);

/* ---------------------------------------------------------------- stage 1: voxelize / pool / gather --- */

/* Scratch sizes for os3d_voxelize (bytes).  hash_cap is returned through *hash_cap (power of two >= 2n). */
int os3d_voxelize_scratch(int64_t n, int64_t *hash_cap, int64_t *n_blocks);

/* Dynamic voxelization with first-occurrence voxel ids.
 * replaces: VoxelGenerator.generate / points_to_voxel (seg3d/core/voxel/voxel_generator.py:24-26,55-153), run per
 *           frame inside WaymoDataset.prepare_data (seg3d/datasets/waymo_dataset.py:275), plus the batch column and
 *           cumulative id offsets collate_batch adds (waymo_dataset.py:347-365).
 * points  [n, stride] f32 rows (batch, x, y, z, ...) when has_batch, else (x, y, z, ...); frames contiguous, ascending.
 * lo[3], vs[3]: range minimum and voxel size (f32, by value); grid[3] = (X, Y, Z) as the reference rounds it.
 * table   [hash_cap] slots, slot_of [n] int32, block_sums [n_blocks + 1] int32 -- scratch.
 * coors   [n, 4] int32 out (b, z, y, x), first *num_voxels rows valid;  pvid [n] int64 out (-1 = outside range).
 * num_voxels: device int32[1].
 */
int os3d_voxelize(const float *points, int64_t n, int stride, int has_batch, float lo_x, float lo_y, float lo_z,
                  float vs_x, float vs_y, float vs_z, int grid_x, int grid_y, int grid_z, os3d_slot_t *table,
                  int64_t hash_cap, int32_t *slot_of, int32_t *block_sums, int64_t n_blocks, int32_t *coors,
                  int64_t *pvid, int32_t *num_voxels, void *stream);

/* Cylinder rows on device: out[i] = (rho, phi, z, x, y, rest...) from (x, y, z, rest...).
 * replaces: cart2polar + concatenate (seg3d/utils/pointops_utils.py:8-11, seg3d/datasets/waymo_dataset.py:270-273).
 * Not bit-exact with numpy's atan2 (last ulp); the bit-exact path feeds host-computed polar rows instead. */
int os3d_cart2polar_rows(const float *in, int64_t n, int in_stride, int has_batch, float *out, void *stream);

/* Point -> voxel scatter reductions; ids < 0 or >= m are skipped.
 * replaces: VFE.forward -> torch_scatter.scatter(reduce='max'|'mean') (seg3d/models/voxel_encoders/vfe.py:24-25),
 *           voxel_max_pooling / voxel_avg_pooling (seg3d/ops/voxel_pooling/voxel_pooling.py:62-79,10-41 and
 *           src/voxel_pooling_cuda.cu:11-24).
 * feats [n, c] f32; ids [n] int64; out [m, c] f32 (fully overwritten).  fix_empty: rows nothing was scattered to
 * become 0 (torch_scatter semantics) at the price of one more pass; counts [m] int32 scratch (mean only; also an
 * output: points per voxel).  counts_in (avg pooling with caller counts) may be NULL. */
int os3d_scatter_max_f32(const float *feats, const int64_t *ids, int64_t n, int c, float *out, int64_t m, int fix_empty,
                         void *stream);
/* the same maximum for bf16 features (c % 8 == 0), bf16 voxel features out: feats [n, c] bf16 read as they are (no fp32
 * copy), ids int32 or int64 (ids_are_i64), acc [m, c] f32 scratch, out [m, c] bf16.  Exact: every maximum is one of the
 * bf16 inputs.  fix_empty as above (fused into the bf16 write). */
int os3d_scatter_max_bf16(const void *feats, const void *ids, int ids_are_i64, int64_t n, int c, float *acc, void *out,
                          int64_t m, int fix_empty, void *stream);
/* the same result without atomics: point rows sorted by voxel id (cub radix sort, ceil(log2(m + 1)) key bits), each
 * run of equal ids reduced by one lane group and written once.  keys / keys_sorted [n] uint32, rows / rows_sorted [n]
 * int32, temp: os3d_scatter_max_sorted_scratch(n, m) bytes.  c % 8 == 0, c <= 2048. */
int os3d_scatter_max_sorted_scratch(int64_t n, int64_t m, int64_t *temp_bytes);
int os3d_scatter_max_sorted_bf16(const void *feats, const void *ids, int ids_are_i64, int64_t n, int c, uint32_t *keys,
                                 uint32_t *keys_sorted, int32_t *rows, int32_t *rows_sorted, void *temp,
                                 int64_t temp_bytes, void *out, int64_t m, int fix_empty, void *stream);
int os3d_scatter_mean_f32(const float *feats, const int64_t *ids, int64_t n, int c, float *out, int32_t *counts,
                          const int32_t *counts_in, int64_t m, void *stream);
/* scatter mean of bf16 rows into a FEW destination rows (m <= 64: the SE layer's per-frame pooling, se_layer.py:17-23):
 * feats [n, c] bf16 read as they are, c % 8 == 0, ids int32 / int64 (sorted runs make it fast, any order is correct),
 * out [m, c] f32, counts [m] int32 (output). */
int os3d_scatter_mean_small_bf16(const void *feats, const void *ids, int ids_are_i64, int64_t n, int c, float *out,
                                 int32_t *counts, int64_t m, void *stream);
/* backward of scatter_max, torch_scatter semantics: the gradient of out[v, c] goes to ONE argmax row (the lowest point
 * index among ties); arg: int32 [m * c] scratch.  mean: grad / count. */
int os3d_scatter_max_bwd_f32(const float *grad_out, const float *feats, const float *out, const int64_t *ids, int64_t n,
                             int c, int64_t m, int32_t *arg, float *grad_in, void *stream);
int os3d_scatter_mean_bwd_f32(const float *grad_out, const int64_t *ids, const int32_t *counts, int64_t n, int c,
                              int64_t m, float *grad_in, void *stream);

/* Voxel -> point gather: out[i] = feats[ids[i]] or 0 when ids[i] < 0.  elem_size 4 (f32) or 2 (bf16).
 * replaces: voxel_to_point (seg3d/ops/voxel_to_point/voxel_to_point.py:5-17). */
int os3d_gather_rows(const void *feats, const int64_t *ids, int64_t n, int c, int elem_size, void *out, void *stream);
/* backward of the gather = scatter-add of rows (f32). */
int os3d_scatter_add_rows_f32(const float *grad_out, const int64_t *ids, int64_t n, int c, float *grad_feats, int64_t m,
                              void *stream);

/* ---------------------------------------------------------------- stage 2: kernel maps (rulebooks) --- */

/* Insert rows of idx [m,4] (b,z,y,x) into an empty-initialised table (this call initialises it).
 * replaces: spconv's hash-table build inside SubMConv3d / SparseConv3d (spconv-cu113, external). */
int os3d_hash_build(const int32_t *idx, int64_t m, int sz, int sy, int sx, os3d_slot_t *table, int64_t cap, void *stream);

/* Submanifold 3x3x3 neighbour table: nbr[i*27+k] = row j with coord[j] = coord[i] + (k - centre), else -1.
 * pair_count: device int32[1] (number of valid pairs; diagnostic / FLOP accounting).
 * replaces: spconv SubMConv3d indice-pair generation (call sites seg3d/models/backbones/pointtransformer.py:26,31,133). */
int os3d_subm_table(const int32_t *idx, int64_t m, int sz, int sy, int sx, const os3d_slot_t *table, int64_t cap,
                    int32_t *nbr, int32_t *pair_count, void *stream);

/* Strided 3x3x3 / stride 2 / pad 1 output sites in ascending linear order.
 * bitmap: [n_words] uint32 scratch over the dense output grid of the whole batch (n_words = ceil(B*oz*oy*ox/32) rounded up to a multiple of 4);
 * word_prefix: [n_words] int32 out (exclusive popcount prefix -> O(1) site->row rank);  block_sums: [n_blocks+1] scratch.
 * out_idx [<= cap_out, 4] int32;  num_out: device int32[1].
 * replaces: spconv SparseConv3d output-index generation (seg3d/utils/spconv_utils.py:19; pointtransformer.py:159-166). */
int os3d_strided_sites(const int32_t *idx, int64_t m, int batch, int oz, int oy, int ox, uint32_t *bitmap,
                       int64_t n_words, int32_t *word_prefix, int32_t *block_sums, int64_t n_blocks, int32_t *out_idx,
                       int64_t cap_out, int32_t *num_out, void *stream);

/* Output-stationary tables of the strided conv and of its inverse:
 *   fwd_nbr[o*27+k] = input row i with i = 2o - 1 + k (else -1)       SparseConv3d
 *   inv_nbr[i*27+k] = output row o with the same pair (else -1)       SparseInverseConv3d
 * replaces: spconv indice pairs of SparseConv3d and their swapped reuse by SparseInverseConv3d (spconv_utils.py:19,22). */
int os3d_strided_tables(const int32_t *idx, int64_t m, int sz, int sy, int sx, const os3d_slot_t *table, int64_t cap,
                        const int32_t *out_idx, int64_t m_out, int oz, int oy, int ox, const uint32_t *bitmap,
                        const int32_t *word_prefix, int32_t *fwd_nbr, int32_t *inv_nbr, int32_t *pair_count, void *stream);

/* ---------------------------------------------------------------- stage 3: sparse convolution --- */

/* out[r, :] = bias + sum_k  in[nbr[r*27+k], :] . W[k]      (gather - GEMM, output-stationary: no scatter atomics)
 * w: [27, cin, cout] (repacked from spconv's [cout, kz, ky, kx, cin] by os3d_pack_weight_*).
 * replaces: SubMConv3d / SparseConv3d / SparseInverseConv3d forward (spconv-cu113; seg3d/utils/spconv_utils.py:16-22). */
int os3d_spconv_fwd_f32(const float *in, const int32_t *nbr, int64_t m_out, int cin, int cout, const float *w,
                        const float *bias, float *out, void *stream);
/* Tile order of a kernel map: perm [m] lists the output rows sorted by their set of neighbour offsets (the 27-bit mask
 * with the rare offsets -- corners, vertical edges -- in the high key bits, so similar sets are adjacent) inside blocks of
 * >= 262144 consecutive rows, so that the rows of a 128-row tile share offsets and the tensor-core kernel skips the others
 * (storage order: ~25 of 27 offsets per tile; sorted: 3.4 for inverse convs, 10.8 for the level-1 submanifold map).  Blocks
 * keep the gather L2-local.  Results do not depend on the order.  keys, keys_sorted, rows: [m] scratch; temp: os3d_kernel_map_order_scratch(m) bytes (radix sort). */
int os3d_kernel_map_order_scratch(int64_t m, int64_t *temp_bytes);
int os3d_kernel_map_order(const int32_t *nbr, int64_t m, uint32_t *keys, uint32_t *keys_sorted, int32_t *rows,
                          int32_t *perm, void *temp, int64_t temp_bytes, void *stream);
/* Tile form of a kernel map for the tensor-core path: nbr [m, 27] -> nbr_t [27][m_pad] (offset-major, m_pad = m rounded
 * up to 128, padding rows -1; row v holds the neighbours of output row perm[v], or of row v when perm is NULL) and
 * tile_mask [m_pad / 128] (bit k set when some row of the 128-row tile has a neighbour at offset k).  Built once per
 * kernel map and shared by every conv that uses the map. */
int os3d_kernel_map_tiles(const int32_t *nbr, int64_t m, const int32_t *perm, int32_t *nbr_t, uint32_t *tile_mask,
                          void *stream);
/* bf16 in / bf16 out, f32 accumulate in TMEM on the tcgen05 tensor cores; the gathered operand rows are fetched by
 * 16-byte cp.async (LDGSTS, zero-fill for absent neighbours) straight into the SWIZZLE_128B UMMA shared-memory layout,
 * the weight K-blocks by cp.async.bulk (TMA tile::gather4 was built, measured 1.0-4.3x slower and dropped: DESIGN.md
 * section 3.1).  Fused epilogue:
 * y = acc*scale[c]+shift[c] (bias and folded BatchNorm), optional residual add [m_out, cout] bf16, optional ReLU.
 * scale/shift (both or neither) and residual may be NULL.  flags: bit 0 = ReLU; bit 1 = the residual has 2*cout
 * channels per row and residual[r, 2c] + residual[r, 2c+1] is added AFTER the ReLU (UpBlock's
 * x_m + channel_reduction(cat), pointtransformer.py:89-110).
 * in: [m_in, cin] bf16 rows, 16-byte aligned, cin % 8 == 0 (callers zero-pad); cout % 16 == 0, cout <= 512
 * (cout % 32 == 0 above 256).  nbr_t / tile_mask / perm: os3d_kernel_map_tiles of the map with that perm (NULL =
 * storage order).  w: os3d_pack_weight_bf16 image.
 * replaces: SubMConv3d / SparseConv3d / SparseInverseConv3d forward (spconv-cu113; seg3d/utils/spconv_utils.py:16-22)
 *           and the BatchNorm1d / ReLU / residual add after them (spconv_utils.py:26-30, pointtransformer.py:47-66). */
int os3d_spconv_fwd_bf16(const void *in, int64_t m_in, const int32_t *nbr_t, const uint32_t *tile_mask,
                         const int32_t *perm, int64_t m_out, int cin, int cout, const void *w, const float *scale,
                         const float *shift, const void *residual, int flags, void *out, void *stream);
/* The same with an output row pitch `ldo` (elements, >= cout, multiple of 8): the rows of `out` may be a column slice of a
 * wider row-major buffer -- the decoder writes both halves of torch.cat([x_bottom, x_trans]) (pointtransformer.py:105)
 * straight into one buffer instead of concatenating. */
int os3d_spconv_fwd_bf16_ld(const void *in, int64_t m_in, const int32_t *nbr_t, const uint32_t *tile_mask,
                            const int32_t *perm, int64_t m_out, int cin, int cout, const void *w, const float *scale,
                            const float *shift, const void *residual, int flags, void *out, int64_t ldo, void *stream);
/* spconv 2.x weight [cout, kz, ky, kx, cin] f32 -> kernel layouts.  f32: [27, cin, cout].  bf16: the shared-memory image
 * of the UMMA B operand, [27 * ceil(cin/64)][cout][128 B swizzled] (os3d_spconv_bf16_packed_elems elements). */
int os3d_spconv_bf16_packed_elems(int cin, int cout, int64_t *elems);
int os3d_pack_weight_f32(const float *w_spconv, int cin, int cout, float *w_packed, void *stream);
int os3d_pack_weight_bf16(const float *w_spconv, int cin, int cout, void *w_packed, void *stream);

/* Linear layer on the same tensor-core kernel (the identity kernel map with one offset), with the epilogues the
 * SWFormer encoder layer needs fused in:  out[r, :n] = epi( x[r, :k] . W^T + bias )
 *   flags bit 0 ReLU; bit 2 exact (erf) GELU; bit 3 out = residual + LayerNorm_n(y) * ln_gamma + ln_beta;
 *   bit 4 y[r, c] += table[tab_idx[r], c] for c < tab_cols (the position-embedding term of q = k = (x + pos) W^T:
 *   pos W^T is a [window volume, n] table).
 * x: [m, k] bf16 (k % 8 == 0), w: os3d_pack_linear_bf16 image of the nn.Linear weight [n, k] f32, bias [n] f32 or NULL,
 * n % 16 == 0, n <= 512; out rows have pitch ldo >= n elements.  residual [m, n] bf16, table [*, tab_cols] bf16.
 * replaces: the nn.Linear / GELU / LayerNorm / residual chain of EncoderLayer.forward and cosine_multi_head_attention_forward's
 *           projections (seg3d/models/layers/point_transformer_layer.py:260-298, seg3d/models/layers/cosine_msa.py:48-63,403). */
int os3d_linear_bf16_packed_elems(int k, int n, int64_t *elems);
int os3d_pack_linear_bf16(const float *w, int k, int n, void *w_packed, void *stream);
int os3d_linear_bf16(const void *x, int64_t m, int k, int n, const void *w, const float *bias, int flags,
                     const void *residual, const float *ln_gamma, const float *ln_beta, float ln_eps, const void *table,
                     const int32_t *tab_idx, int tab_cols, void *out, int64_t ldo, void *stream);

/* The same operation on the persistent Linear kernel (linear_tc.cu): weights resident in shared memory, activations by
 * TMA tile loads, double-buffered accumulators so the epilogue of one row tile overlaps the loads and MMAs of the next.
 * Same arguments as os3d_linear_bf16; needs n <= 256 and the weight image to fit next to the activation ring
 * (os3d_linear_tc_fits(k, n) != 0) -- callers fall back to os3d_linear_bf16 otherwise. */
int os3d_linear_tc_fits(int k, int n);
int os3d_linear_tc_bf16(const void *x, int64_t m, int k, int n, const void *w, const float *bias, int flags,
                        const void *residual, const float *ln_gamma, const float *ln_beta, float ln_eps, const void *table,
                        const int32_t *tab_idx, int tab_cols, void *out, int64_t ldo, void *stream);

/* Wide Linear layers with streamed weights and staged, coalesced stores (qkv_tc.cu):
 *   y = x . W^T,  x [m, k] bf16 (k % 8 == 0, k <= 512),  W [n, k],  out [m, n] bf16 with row pitch ldo.
 * The n output columns are processed in chunks of nc columns (os3d_wide_linear_plan(k, n, n_norm, dp, &nc) != 0 when the
 * problem fits; nc is a multiple of dp that divides n and n_norm).  Columns [0, n_norm): y += table[tab_idx[r], col] (table
 * bf16 [*, tab_ld]; NULL: y += bias) and, when `normalize`, every group of dp columns (one attention head) is L2-normalised
 * (F.normalize, eps 1e-12).  Columns [n_norm, n): y += bias (NULL: nothing), then GELU (erf) when mode_rest == 2 (0: none).
 * w_img: per chunk c the os3d_pack_linear_bf16 image of W[c*nc : (c+1)*nc, :], chunks concatenated.
 * replaces: the q / k / v in-projections of cosine_multi_head_attention_forward with q = k = x + pos, v = x
 *           (seg3d/models/layers/cosine_msa.py:48-63; point_transformer_layer.py:248) together with F.normalize of q and
 *           k (cosine_msa.py:152-153); and MLP.fc1 + GELU (point_transformer_layer.py:260-276) where the hidden width
 *           is beyond os3d_swformer_mlp_bf16. */
int os3d_wide_linear_plan(int k, int n, int n_norm, int dp, int *nc);
int os3d_wide_linear_bf16(const void *x, int64_t m, int k, int n, int dp, const void *w_img, const float *bias,
                          const void *table, const int32_t *tab_idx, int64_t tab_ld, int n_norm, int normalize,
                          int mode_rest, void *out, int64_t ldo, void *stream);

/* A chain of 2..4 Linear layers in one persistent kernel (mlp_tc.cu), activations kept in shared / tensor memory:
 *   y = L_{n-1}(... act_0(L_0(a0))),  L_l(a) = a . W_l^T + b_l,  act: 0 none, 1 ReLU, 2 exact (erf) GELU.
 * a0 is either the bf16 matrix x [m, layers[0].k] (pitch ldx elements; x32 = NULL) or the output of an fp32 front layer
 * computed in the kernel from x32 [m, k32 <= 16] (pitch ld32 floats): a0 = act32(x32 . w32 + b32), w32 [k32][64] fp32,
 * 64 wide (x = NULL).  The last layer may add `residual` [m, n] (pitch ldr) after an optional LayerNorm (gamma / beta
 * non-NULL): out = residual + LN(L(a)), as in EncoderLayer.  out: bf16 (or fp32 when out_f32) [m, n_out] with pitch ldo
 * elements; n_out <= layers[last].n drops zero-padded output columns.
 * layers[l].w is an os3d_pack_linear_bf16 image of the [n, k] weight; widths are multiples of 16 up to 256 and
 * layers[l].k == layers[l-1].n.  os3d_mlp_chain_fits() != 0 when all weight images fit in shared memory.
 * replaces: point_encoder / fusion_encoder / classifier of Segformer with eval-mode BatchNorm folded
 *           (seg3d/models/segmentors/segformer.py:21-32,58-76,105-116) and the MLP + norm2 of EncoderLayer
 *           (seg3d/models/layers/point_transformer_layer.py:260-298). */
typedef struct {
  const void *w;       /* packed bf16 weight image */
  const float *bias;   /* [n] or NULL */
  int k, n, act;
} os3d_mlp_layer;
int os3d_mlp_chain_fits(const os3d_mlp_layer *layers, int n_layers, int has_front);
int os3d_mlp_chain_bf16(const void *x, int64_t m, int64_t ldx, const float *x32, int64_t ld32, int k32, const float *w32,
                        const float *b32, int act32, const os3d_mlp_layer *layers, int n_layers, const void *residual,
                        int64_t ldr, const float *ln_gamma, const float *ln_beta, float ln_eps, void *out, int64_t ldo,
                        int n_out, int out_f32, void *stream);

/* The SWFormer MLP with its norm and residual as one persistent kernel (mlp2_tc.cu), weights streamed from L2 in chunks
 * of the hidden dimension so that C = 192 (2 x 147 KB of weights) runs too:
 *   out[m, c] = x + LayerNorm(fc2(GELU(fc1(x))))      x, out bf16 [m, c] contiguous; w1 / w2: os3d_pack_linear_bf16 images
 * of fc1.weight [h, c] and fc2.weight [c, h]; exact (erf) GELU.  c % 16 == 0, c <= 192; h % 64 == 0 or h <= 128
 * (os3d_swformer_mlp_fits).
 * replaces: MLP + norm2 + residual of EncoderLayer.forward (seg3d/models/layers/point_transformer_layer.py:260-298). */
int os3d_swformer_mlp_fits(int c, int h);
int os3d_swformer_mlp_bf16(const void *x, int64_t m, int c, int h, const void *w1, const float *b1, const void *w2,
                           const float *b2, const float *ln_gamma, const float *ln_beta, float ln_eps, void *out,
                           void *stream);

/* ---------------------------------------------------------------- stage 4: window partition + attention --- */

#define OS3D_MAX_LEVELS 4
typedef struct {
  int sparse_x, sparse_y, sparse_z;      /* voxel grid of this stage (x, y, z) */
  int win_x, win_y, win_z;               /* window shape                        */
  int nwin_x, nwin_y, nwin_z;            /* ceil(sparse/win) + 1                */
  int shift_x, shift_y, shift_z;         /* per get_window_coors                */
  int n_levels;
  int lvl_lo[OS3D_MAX_LEVELS], lvl_hi[OS3D_MAX_LEVELS], lvl_tokens[OS3D_MAX_LEVELS];
} os3d_window_cfg_t;

/* Window partition of one shift, no host synchronisation.
 * replaces: get_window_coors (seg3d/utils/swformer_utils.py:109-154), get_inner_win_inds
 *           (seg3d/ops/ingroup_inds/src/ingroup_inds_cuda.cu:12-52), batching_single_shift
 *           (seg3d/models/layers/point_transformer_layer.py:71-87), make_continuous_inds / get_flat2win_inds
 *           (swformer_utils.py:8-31,158-171).
 * idx [m,4] int32.  With n_win = batch * nwin_x*nwin_y*nwin_z dense window ids and n_blocks = ceil(n_win / 1024):
 * scratch  win_count [n_win] int32, win_meta [n_win * 3] int32, block_sums [(n_blocks + 1) * 5] int32.
 * Per-voxel outputs: win_id [m] int64 (batch_win_inds), in_win [m,3] int32 (z,y,x), level [m] int32 (-1 = occupancy
 *   outside every batching range), win_rank [m] int32 (rank of the voxel's window among the windows of its level,
 *   ascending window id = make_continuous_inds), inner [m] int32 (stable rank inside the window; the reference's
 *   atomicAdd rank is a race, this is its deterministic member).  flat2window slot = win_rank * max_tokens + inner.
 * Segment outputs: order [m] int32 = voxel rows grouped by window (ascending window id, ascending row inside);
 *   seg_start / seg_len [<= m + 1] int32: one entry per non-empty window, level-major (all level-0 windows first),
 *   ascending window id inside a level;  pos_seg [m, 2] int32: (start, length) in `order` of the window that owns
 *   each grouped position (what the tensor-core attention walks);
 *   level_info: device int32[16] = { n_windows[4], first_window[4], 0,0,0,0, tokens_outside_ranges, n_windows_total,
 *   n_tokens_assigned, tokens_over_capacity }.  (The reference drops over-capacity tokens and then cannot continue
 *   -- SURVEY.md Appendix C; callers treat a non-zero count as an error.)
 *   pos_idx [m] int32 (may be NULL): row of each voxel in the window's position-embedding table,
 *   (z * win_y + y) * win_x + x of its in-window coordinates. */
int os3d_window_partition(const int32_t *idx, int64_t m, int batch, const os3d_window_cfg_t *cfg, int32_t *win_count,
                          int32_t *win_meta, int32_t *block_sums, int64_t n_blocks, int64_t *win_id, int32_t *in_win,
                          int32_t *level, int32_t *win_rank, int32_t *inner, int32_t *order, int32_t *seg_start,
                          int32_t *seg_len, int32_t *pos_seg, int32_t *level_info, int32_t *pos_idx, void *stream);

/* The same partition over caller-supplied group ids in [0, n_groups) (cfg supplies only the batching levels).
 * replaces: get_inner_win_inds as a stand-alone op (seg3d/ops/ingroup_inds/ingroup_inds.py:7-20): `inner` is the rank. */
int os3d_group_partition(const int64_t *group, int64_t n, int64_t n_groups, const os3d_window_cfg_t *cfg, int32_t *count,
                         int32_t *meta, int32_t *block_sums, int64_t n_blocks, int32_t *level, int32_t *group_rank,
                         int32_t *inner, int32_t *order, int32_t *seg_start, int32_t *seg_len, int32_t *pos_seg,
                         int32_t *level_info, void *stream);

/* Sinusoidal window position embedding, flat [m, c] (f32 or bf16 by elem_size).
 * replaces: SparseWindowPartitionLayer.get_pos_embed (point_transformer_layer.py:152-207). */
int os3d_pos_embed(const int32_t *in_win, int64_t m, int c, int win_x, int win_y, int win_z, float temperature,
                   int elem_size, void *out, void *stream);
/* out[r, :] = x[r, :] * (bias + table[idx[r], :]): per-point application of the per-frame squeeze-excite gate
 * (table f32 [frames, c], idx int64 [m] = batch index; bias 1 folds the segmentor's residual x + x * gate).
 * replaces: `x * y[indices]` of FlattenSELayer.forward (se_layer.py:29-30) and the add in segformer.py:134. */
int os3d_scale_rows_by_table(const void *x, const float *table, const int64_t *idx, int64_t m, int c, float bias,
                             int elem_size, void *out, void *stream);
/* out = GELU(x) (erf form) on n bf16 elements (n % 8 == 0; out may alias x).  erf by Abramowitz-Stegun 7.1.26,
 * |error| <= 1.5e-7: at most one bf16 ulp from the exact erf GELU (98.7 % of results bit-identical after rounding).
 * replaces: MLP.act = nn.GELU() between fc1 and fc2 (point_transformer_layer.py:260-276). */
int os3d_gelu_bf16(const void *x, int64_t n, void *out, void *stream);
/* out[r, :] = x[r, :] + table[idx[r], :] (f32 or bf16 by elem_size, c % 8 == 0): q = k = x + pos with the position
 * embedding as a [window volume, c] table indexed by the in-window position.
 * replaces: the flat2window'ed pos-embed add of WindowAttention.forward (point_transformer_layer.py:248). */
int os3d_add_table_rows(const void *x, const void *table, const int32_t *idx, int64_t m, int c, int elem_size, void *out,
                        void *stream);

/* In-place L2 normalisation of every head slice of q and k rows (F.normalize, eps 1e-12).
 * replaces: cosine_msa.py:152-153. */
int os3d_qk_normalize(void *q, void *k, int64_t ld, int64_t m, int c, int heads, int elem_size, void *stream);

/* Variable-length cosine window attention over the segments of os3d_window_partition.
 * q, k: [m] rows with pitch ld (elements), already head-normalised; v: pitch ldv; heads*d == c.  Per (window, head):
 *   softmax( q k^T / max(tau, tau_min) ) v        -- no padding, no masks, no score tensor in memory.
 * tau: device f32[1];  lvl_tokens: host int[4] (max_tokens per batching level, sizes the work decomposition).
 * out [m, c] in the original voxel order.  elem_size 4 (f32) or 2 (bf16).
 * replaces: flat2window + CosineMultiheadAttention core + window2flat (swformer_utils.py:34-85,
 *           seg3d/models/layers/cosine_msa.py:115-177, point_transformer_layer.py:233-258).
 * drop_p > 0 (training): attention dropout on the softmax weights (cosine_msa.py:173-174) with a keep-mask hashed from
 * (seed, head, query row, key row) -- reproducible from the seed, regenerated by the backward. */
int os3d_window_attention(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m, int c, int heads,
                          const int32_t *order, const int32_t *seg_start, const int32_t *seg_len,
                          const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min,
                          float drop_p, uint64_t seed, int elem_size, void *out, void *stream);

/* Backward of os3d_window_attention.  q, k (normalised), v, o (the forward's output), go (gradient of o), gq, gk, gv:
 * [m, c] rows of pitch c.  stats: scratch f32 [m * heads * 3].  g_inv_tau: device f32[1], zeroed by the caller, receives
 * d loss / d (1 / max(tau, tau_min)).  Two kernels: query-stationary (row statistics, gq, g_inv_tau) and key-stationary
 * (gk, gv) -- no atomics on the row gradients.
 * replaces: autograd through torch.bmm / softmax / dropout of _scaled_cosine_attention (cosine_msa.py:152-176). */
int os3d_window_attention_bwd(const void *q, const void *k, const void *v, const void *o, const void *go, int64_t m, int c,
                              int heads, const int32_t *order, const int32_t *seg_start, const int32_t *seg_len,
                              const int32_t *level_info, const int *lvl_tokens, const float *tau, float tau_min,
                              float drop_p, uint64_t seed, int elem_size, float *stats, void *gq, void *gk, void *gv,
                              float *g_inv_tau, void *stream);

/* The same attention on the tcgen05 tensor cores (bf16): QK^T and PV as UMMA tiles with TMEM accumulators, q / k
 * normalisation folded into the gather (do NOT call os3d_qk_normalize first).  Head-padded layout: head h occupies
 * columns [h*dp, (h+1)*dp) of q / k / v / out rows with dp = head_dim rounded up to 16 (16, 32 or 48; pad columns
 * zero -- the Python layer pads the projection weights).  pos_seg from os3d_window_partition. */
int os3d_window_attention_bf16_tc(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m,
                                  int heads, int dp, const int32_t *order, const int32_t *pos_seg,
                                  const int32_t *level_info, const float *tau, float tau_min, void *out, int64_t ldo,
                                  void *stream);

/* The same kernel for q, k that are ALREADY L2-normalised per head (os3d_wide_linear_bf16 with normalize = 1): the gather
 * is a plain copy. */
int os3d_window_attention_bf16_tc_prenorm(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m,
                                          int heads, int dp, const int32_t *order, const int32_t *pos_seg,
                                          const int32_t *level_info, const float *tau, float tau_min, void *out,
                                          int64_t ldo, void *stream);
/* the training forward of the same kernel: q / k pre-normalised, attention dropout (0 <= drop_p < 1) applied to the
 * normalised weights with the keep mask of os3d_window_attention (a hash of seed, head, query row, key row), which
 * os3d_window_attention_bwd regenerates from the same seed. */
int os3d_window_attention_bf16_tc_drop(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m,
                                       int heads, int dp, const int32_t *order, const int32_t *pos_seg,
                                       const int32_t *level_info, const float *tau, float tau_min, float drop_p,
                                       uint64_t seed, void *out, int64_t ldo, void *stream);

/* Second tensor-core design of the same attention (attention_v2.cu): one CTA per (128-query tile, group of heads whose
 * slices make 96-128 columns), dedicated loader / MMA-issuer / softmax warps with mbarrier hand-offs, whole row slices
 * fetched once for all heads of the group.  Same arguments and head-padded layout as os3d_window_attention_bf16_tc, but
 * q and k must ALREADY be L2-normalised per head (os3d_qk_normalize with c = heads * dp, or the projection kernel's
 * epilogue), and the softmax uses the fixed maximum 1 (unit vectors bound the scores): the caller must make sure that
 * log2(e) / max(tau, tau_min) <= 60, i.e. max(tau, tau_min) >= 0.02405 -- below that use os3d_window_attention_bf16_tc
 * (online maximum).  heads * dp must split into groups of 128 columns (dp = 16, 32) or 96 (dp = 48). */
int os3d_window_attention_bf16_v2(const void *q, const void *k, const void *v, int64_t ld, int64_t ldv, int64_t m,
                                  int heads, int dp, const int32_t *order, const int32_t *pos_seg,
                                  const int32_t *level_info, const float *tau, float tau_min, void *out, int64_t ldo,
                                  void *stream);

/* out = resid + LayerNorm(x) * w + b over rows of c elements (c % 8 == 0); resid may be NULL; w, b f32.
 * replaces: norm1 / norm2 + the residual adds of EncoderLayer.forward (point_transformer_layer.py:288-298). */
int os3d_layernorm_residual(const void *x, const void *resid, const float *w, const float *b, int64_t m, int c,
                            float eps, int elem_size, void *out, void *stream);

/* ---------------------------------------------------------------- loss side --- */

/* k nearest points (squared distances, ascending) of every query among the points of its own batch segment.
 * xyz [n, 3], new_xyz [m, 3] f32; offset / new_offset [n_seg] int32: cumulative segment ends of xyz / new_xyz;
 * idx [m, nsample] int32, dist2 [m, nsample] f32 (squared); 1 <= nsample <= 100.  Exact brute force with the
 * reference's heap semantics (ties keep the lower index).
 * replaces: knn_query_ext.knn_query_cuda (seg3d/ops/knn_query/src/knn_query_cuda.cu:67-133; tools/train.py:103). */
int os3d_knn_query(const float *xyz, const float *new_xyz, int64_t m, int nsample, const int32_t *offset,
                   const int32_t *new_offset, int n_seg, int32_t *idx, float *dist2, void *stream);

/* Majority label of every voxel over its points (ties -> the lowest label, voxels without a labelled point -> ignore).
 * pvid [n] int64 point -> voxel ids (-1 = outside), labels [n] uint8 in [0, 30] or == ignore (31 <= ignore <= 255);
 * hist [m * 32] int32 scratch, bad: device int32[1] (set to 1 when a label outside that set was seen), out [m] uint8.
 * replaces: WaymoDataset.prepare_voxel_labels (seg3d/datasets/waymo_dataset.py:213-246). */
int os3d_voxel_majority_labels(const int64_t *pvid, const uint8_t *labels, int64_t n, int64_t m, int ignore, int32_t *hist,
                               int32_t *bad, uint8_t *out, void *stream);

/* Predicted class per point: out[i] = argmax_j x[i, j] as uint8 (ties -> the lowest j).
 * replaces: torch.argmax(point_out, dim=1) of tools/test.py:58.  x [n, c] bf16 (elem_size 2) or f32 (4), c <= 256. */
int os3d_argmax_rows(const void *x, int64_t n, int c, int elem_size, uint8_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OS3D_H_ */
