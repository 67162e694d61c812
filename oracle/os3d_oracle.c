/*
 * os3d_oracle.c -- CPU oracle for the integer stages of the OpenSeg3D voxel-backbone hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under openseg3d_b200/ may import, link or call this file;
 * it is the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  It is a plain, serial restatement of the reference's algorithms.
 *
 *  - os3d_oracle_voxelize  : reference seg3d/core/voxel/voxel_generator.py:99-153
 *                            (_points_to_voxel_reverse_kernel) + the per-frame id offsets that
 *                            collate_batch applies (seg3d/datasets/waymo_dataset.py:347-365).
 *                            Parity PINNED: checked against the reference's own numba code run in the
 *                            build container (tests/golden/voxelize_*.npz, made by
 *                            tests/golden/make_golden.py).
 *  - os3d_oracle_subm_map / os3d_oracle_strided_map
 *                          : kernel-map (rulebook) semantics of spconv 2.x SubMConv3d / SparseConv3d /
 *                            SparseInverseConv3d as used at reference seg3d/utils/spconv_utils.py:16-22 and
 *                            seg3d/models/backbones/pointtransformer.py:26,31,133,159-166.  spconv is an
 *                            un-vendored, unversioned dependency (requirements.txt:4 "spconv-cu113") that is
 *                            absent from /root/reference and from this image => PARITY UNPINNED against
 *                            spconv itself; the semantics are pinned instead against torch's dense conv3d /
 *                            conv_transpose3d in tests/test_oracle_spconv.py.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Stage 1a: dynamic voxelization, first-occurrence voxel ids.
 * points: [n, stride] float32, row = (batch, x, y, z, ...) if has_batch else (x, y, z, ...).
 * Points of one frame must be contiguous and frames ascending (what collate_batch produces).
 * range[6] = xyz min, xyz max (f32); vsize[3] (f32).  All arithmetic is IEEE float32, exactly as numba
 * evaluates `np.floor((points[i, j] - coors_range[j]) / voxel_size[j])` on float32 arrays
 * (voxel_generator.py:139) and grid = round((hi-lo)/vs) (voxel_generator.py:129-132).
 * Outputs: coors [<=n, 4] int32 (b, z, y, x); pvid [n] int64 (-1 = dropped); returns voxel count.
 * ---------------------------------------------------------------------------------------------- */
int64_t os3d_oracle_voxelize(const float *points, int64_t n, int stride, int has_batch, const float *range,
                             const float *vsize, int32_t *coors, int64_t *pvid, int32_t *grid_out) {
  int32_t grid[3];
  for (int j = 0; j < 3; ++j) {
    volatile float g = (range[3 + j] - range[j]) / vsize[j];
    grid[j] = (int32_t)rintf(g); /* np.round = round-half-even */
    if (grid_out) grid_out[j] = grid[j];
  }
  const int64_t cells = (int64_t)grid[0] * grid[1] * grid[2];
  int32_t *lut = (int32_t *)malloc(sizeof(int32_t) * (size_t)cells);
  if (!lut) return -1;
  int64_t num = 0;     /* voxels emitted so far over all frames          */
  int64_t base = 0;    /* voxel-id offset of the current frame           */
  int cur_b = -1;
  const int off = has_batch ? 1 : 0;
  for (int64_t i = 0; i < n; ++i) {
    const float *p = points + i * stride;
    int b = has_batch ? (int)p[0] : 0;
    if (b != cur_b) { /* new frame: fresh lookup grid (coor_to_voxelidx = -ones, :83) */
      memset(lut, 0xff, sizeof(int32_t) * (size_t)cells);
      cur_b = b;
      base = num;
    }
    int32_t c[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      volatile float d = p[off + j] - range[j];
      volatile float q = d / vsize[j];
      float f = floorf(q);
      if (f < 0.0f || f >= (float)grid[j]) { failed = 1; break; }
      c[j] = (int32_t)f;
    }
    if (failed) { pvid[i] = -1; continue; }
    int64_t cell = ((int64_t)c[2] * grid[1] + c[1]) * grid[0] + c[0]; /* (z, y, x) */
    int32_t v = lut[cell];
    if (v == -1) {
      v = (int32_t)(num - base);
      lut[cell] = v;
      int32_t *o = coors + num * 4;
      o[0] = b; o[1] = c[2]; o[2] = c[1]; o[3] = c[0];
      ++num;
    }
    pvid[i] = base + v;
  }
  free(lut);
  return num;
}

/* ------------------------------------------------------------------------------------------------
 * Open-addressing map (linear voxel index -> row) for the rulebook oracle.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t *keys; int32_t *vals; uint64_t mask; } map_t;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
static int map_init(map_t *m, int64_t n) {
  uint64_t cap = 16;
  while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
  m->keys = (int64_t *)malloc(sizeof(int64_t) * cap);
  m->vals = (int32_t *)malloc(sizeof(int32_t) * cap);
  if (!m->keys || !m->vals) return -1;
  for (uint64_t i = 0; i < cap; ++i) m->keys[i] = -1;
  m->mask = cap - 1;
  return 0;
}
static void map_free(map_t *m) { free(m->keys); free(m->vals); }
static int32_t map_get(const map_t *m, int64_t key) {
  uint64_t h = mix64((uint64_t)key) & m->mask;
  while (m->keys[h] != -1) {
    if (m->keys[h] == key) return m->vals[h];
    h = (h + 1) & m->mask;
  }
  return -1;
}
/* insert if absent; returns the stored value */
static int32_t map_put(map_t *m, int64_t key, int32_t val) {
  uint64_t h = mix64((uint64_t)key) & m->mask;
  while (m->keys[h] != -1) {
    if (m->keys[h] == key) return m->vals[h];
    h = (h + 1) & m->mask;
  }
  m->keys[h] = key; m->vals[h] = val;
  return val;
}
static int64_t lin(const int32_t *c, const int32_t *shape) { /* c = (b,z,y,x), shape = (Z,Y,X) */
  return (((int64_t)c[0] * shape[0] + c[1]) * shape[1] + c[2]) * shape[2] + c[3];
}

/* ------------------------------------------------------------------------------------------------
 * Stage 2a: submanifold 3x3x3 (pad 1, stride 1, dilation 1) neighbour table.
 * nbr[i*27 + k] = j such that coord[j] = coord[i] + (k - centre), else -1;  k = (kz*3+ky)*3+kx.
 * Returns the number of (k, j -> i) pairs.
 * ---------------------------------------------------------------------------------------------- */
int64_t os3d_oracle_subm_map(const int32_t *idx, int64_t m, const int32_t *shape, int32_t *nbr) {
  map_t mp;
  if (map_init(&mp, m)) return -1;
  for (int64_t i = 0; i < m; ++i) map_put(&mp, lin(idx + 4 * i, shape), (int32_t)i);
  int64_t pairs = 0;
  for (int64_t i = 0; i < m; ++i) {
    const int32_t *c = idx + 4 * i;
    for (int kz = 0; kz < 3; ++kz) for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
      int k = (kz * 3 + ky) * 3 + kx;
      int32_t q[4] = {c[0], c[1] + kz - 1, c[2] + ky - 1, c[3] + kx - 1};
      int32_t j = -1;
      if (q[1] >= 0 && q[1] < shape[0] && q[2] >= 0 && q[2] < shape[1] && q[3] >= 0 && q[3] < shape[2])
        j = map_get(&mp, lin(q, shape));
      nbr[i * 27 + k] = j;
      pairs += (j >= 0);
    }
  }
  map_free(&mp);
  return pairs;
}

static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------------
 * Stage 2b: strided 3x3x3 conv, stride 2, pad 1 (SparseConv3d) -- output sites, and both
 * output-stationary tables:
 *   out_shape[d] = (shape[d] + 2 - 3) / 2 + 1
 *   site o active iff an active input i and offset k satisfy  i = 2*o - 1 + k,  0 <= o < out_shape
 *   out_idx: canonical order = ascending ((b*Z+z)*Y+y)*X+x over out_shape      [<= 8*m rows, caller sizes it]
 *   fwd_nbr[o*27+k] = input row i (or -1)      -> SparseConv3d        out[o] += W[k] in[i]
 *   inv_nbr[i*27+k] = output row o (or -1)     -> SparseInverseConv3d out[i] += Winv[k] in[o]
 * Returns the number of output sites; *pairs_out = number of (k, i, o) pairs.
 * ---------------------------------------------------------------------------------------------- */
int64_t os3d_oracle_strided_map(const int32_t *idx, int64_t m, const int32_t *shape, int32_t *out_shape,
                                int32_t *out_idx, int32_t *fwd_nbr, int32_t *inv_nbr, int64_t *pairs_out) {
  for (int d = 0; d < 3; ++d) out_shape[d] = (shape[d] + 2 - 3) / 2 + 1;
  int64_t *cand = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m * 8 + 1));
  if (!cand) return -1;
  int64_t nc = 0;
  for (int64_t i = 0; i < m; ++i) {
    const int32_t *c = idx + 4 * i;
    for (int k = 0; k < 27; ++k) {
      int kk[3] = {k / 9, (k / 3) % 3, k % 3};
      int32_t o[4] = {c[0], 0, 0, 0};
      int ok = 1;
      for (int d = 0; d < 3; ++d) {
        int t = c[1 + d] + 1 - kk[d];
        if (t < 0 || (t & 1)) { ok = 0; break; }
        t >>= 1;
        if (t >= out_shape[d]) { ok = 0; break; }
        o[1 + d] = t;
      }
      if (ok) cand[nc++] = lin(o, out_shape);
    }
  }
  qsort(cand, (size_t)nc, sizeof(int64_t), cmp_i64);
  int64_t mo = 0;
  for (int64_t t = 0; t < nc; ++t)
    if (t == 0 || cand[t] != cand[t - 1]) cand[mo++] = cand[t];
  map_t mp;
  if (map_init(&mp, mo)) { free(cand); return -1; }
  for (int64_t r = 0; r < mo; ++r) {
    int64_t key = cand[r];
    map_put(&mp, key, (int32_t)r);
    int32_t *o = out_idx + 4 * r;
    o[3] = (int32_t)(key % out_shape[2]); key /= out_shape[2];
    o[2] = (int32_t)(key % out_shape[1]); key /= out_shape[1];
    o[1] = (int32_t)(key % out_shape[0]); key /= out_shape[0];
    o[0] = (int32_t)key;
  }
  for (int64_t t = 0; t < mo * 27; ++t) fwd_nbr[t] = -1;
  int64_t pairs = 0;
  for (int64_t i = 0; i < m; ++i) {
    const int32_t *c = idx + 4 * i;
    for (int k = 0; k < 27; ++k) {
      int kk[3] = {k / 9, (k / 3) % 3, k % 3};
      int32_t o[4] = {c[0], 0, 0, 0};
      int ok = 1;
      for (int d = 0; d < 3; ++d) {
        int t = c[1 + d] + 1 - kk[d];
        if (t < 0 || (t & 1)) { ok = 0; break; }
        t >>= 1;
        if (t >= out_shape[d]) { ok = 0; break; }
        o[1 + d] = t;
      }
      int32_t r = ok ? map_get(&mp, lin(o, out_shape)) : -1;
      inv_nbr[i * 27 + k] = r;
      if (r >= 0) { fwd_nbr[(int64_t)r * 27 + k] = (int32_t)i; ++pairs; }
    }
  }
  if (pairs_out) *pairs_out = pairs;
  map_free(&mp);
  free(cand);
  return mo;
}

/* ------------------------------------------------------------------------------------------------
 * Stage 3: output-stationary sparse convolution, float32 in / float64 accumulate.
 *   out[r, :] = bias + sum_k  W[:, k, :] . in[nbr[r*27+k], :]      W layout [Cout, 27, Cin] (spconv 2.x
 *   [Cout, kz, ky, kx, Cin] flattened).  Used only on small cases; the big cases use the torch-CPU
 *   restatement in oracle/oracle.py, which is itself checked against this loop.
 * ---------------------------------------------------------------------------------------------- */
void os3d_oracle_spconv(const float *in, const int32_t *nbr, int64_t m_out, int cin, int cout, const float *w,
                        const float *bias, float *out) {
  double *acc = (double *)malloc(sizeof(double) * (size_t)cout);
  for (int64_t r = 0; r < m_out; ++r) {
    for (int o = 0; o < cout; ++o) acc[o] = bias ? (double)bias[o] : 0.0;
    for (int k = 0; k < 27; ++k) {
      int32_t j = nbr[r * 27 + k];
      if (j < 0) continue;
      const float *x = in + (int64_t)j * cin;
      for (int o = 0; o < cout; ++o) {
        const float *wr = w + ((int64_t)o * 27 + k) * cin;
        double s = 0.0;
        for (int c = 0; c < cin; ++c) s += (double)wr[c] * (double)x[c];
        acc[o] += s;
      }
    }
    for (int o = 0; o < cout; ++o) out[r * cout + o] = (float)acc[o];
  }
  free(acc);
}

/* ------------------------------------------------------------------------------------------------
 * Stage 4 (integer part): stable within-group rank.  The reference's ingroup_inds kernel
 * (seg3d/ops/ingroup_inds/src/ingroup_inds_cuda.cu:12-25) returns an arrival-order rank (atomicAdd race);
 * any permutation inside a group is a valid reference output.  The oracle (and the CUDA path) pick the
 * deterministic member of that set: rank = number of earlier elements with the same group id.
 * ---------------------------------------------------------------------------------------------- */
void os3d_oracle_ingroup_rank(const int64_t *group, int64_t n, int64_t *rank) {
  int64_t mx = -1;
  for (int64_t i = 0; i < n; ++i) if (group[i] > mx) mx = group[i];
  int64_t *cnt = (int64_t *)calloc((size_t)(mx + 2), sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) rank[i] = cnt[group[i]]++;
  free(cnt);
}
