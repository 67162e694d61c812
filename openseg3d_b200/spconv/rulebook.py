"""Kernel maps (rulebooks) built on the GPU by libos3d (stage 2)."""
import torch

from .. import _lib


def _pow2_cap(m):
    cap = 1024
    while cap < 2 * m:
        cap <<= 1
    return cap


def build_site_table(indices, spatial_shape):
    """Open-addressing table linear (b,z,y,x) index -> row.  Returns (table int64 [2*cap], cap)."""
    m = indices.shape[0]
    cap = _pow2_cap(m)
    table = torch.empty(cap * 2, dtype=torch.int64, device=indices.device)
    z, y, x = spatial_shape
    _lib.call('os3d_hash_build', indices, m, z, y, x, table, cap)
    return table, cap


class SubmRulebook(object):
    """nbr [M, 27] int32: row of the neighbour at offset k (or -1); shared by every SubMConv3d of one indice_key."""
    kind = 'subm'

    def __init__(self, nbr, pair_count, indices, spatial_shape):
        self.nbr, self.pair_count, self.indices, self.spatial_shape = nbr, pair_count, indices, spatial_shape

    @property
    def num_pairs(self):
        return int(self.pair_count.item())


class StridedRulebook(object):
    """SparseConv3d(k3, s2, p1): out_indices (ascending linear index), fwd_nbr [M_out, 27], inv_nbr [M_in, 27]."""
    kind = 'strided'

    def __init__(self, in_indices, in_shape, out_indices, out_shape, fwd_nbr, inv_nbr, pair_count):
        self.in_indices, self.in_shape = in_indices, in_shape
        self.out_indices, self.out_shape = out_indices, out_shape
        self.fwd_nbr, self.inv_nbr, self.pair_count = fwd_nbr, inv_nbr, pair_count

    @property
    def num_pairs(self):
        return int(self.pair_count.item())


def _table_of(x):
    key = id(x.indices)
    hit = x._site_table.get(key)
    if hit is None or hit[2] is not x.indices:
        table, cap = build_site_table(x.indices, x.spatial_shape)
        hit = (table, cap, x.indices)
        x._site_table[key] = hit
    return hit[0], hit[1]


def build_subm_rulebook(x):
    _lib.require_cuda(x.indices)
    m = x.indices.shape[0]
    table, cap = _table_of(x)
    nbr = torch.empty((m, 27), dtype=torch.int32, device=x.indices.device)
    pairs = torch.empty(1, dtype=torch.int32, device=x.indices.device)     # zeroed by the entry point
    z, y, xx = x.spatial_shape
    _lib.call('os3d_subm_table', x.indices, m, z, y, xx, table, cap, nbr, pairs, work=lambda: m * (16 + 27 * 4))
    return SubmRulebook(nbr, pairs, x.indices, list(x.spatial_shape))


def build_strided_rulebook(x):
    """kernel 3, stride 2, padding 1.  One device->host read (the number of output sites) sizes the result."""
    _lib.require_cuda(x.indices)
    dev = x.indices.device
    m = x.indices.shape[0]
    sz, sy, sx = x.spatial_shape
    oz, oy, ox = [(s + 2 - 3) // 2 + 1 for s in (sz, sy, sx)]
    cells = x.batch_size * oz * oy * ox
    n_words = ((cells + 31) // 32 + 3) // 4 * 4
    n_blocks = (n_words + 1023) // 1024
    bitmap = torch.empty(n_words, dtype=torch.int32, device=dev)
    prefix = torch.empty(n_words, dtype=torch.int32, device=dev)
    block_sums = torch.empty(n_blocks + 1, dtype=torch.int32, device=dev)
    cap_out = min(8 * m, cells) + 1
    out_idx = torch.empty((cap_out, 4), dtype=torch.int32, device=dev)
    num = torch.empty(1, dtype=torch.int32, device=dev)                # written by the entry point
    _lib.call('os3d_strided_sites', x.indices, m, x.batch_size, oz, oy, ox, bitmap, n_words, prefix, block_sums, n_blocks,
              out_idx, cap_out, num)
    table, cap = _table_of(x)
    m_out = int(num.item())
    out_idx = out_idx[:m_out]
    fwd = torch.empty((m_out, 27), dtype=torch.int32, device=dev)
    inv = torch.empty((m, 27), dtype=torch.int32, device=dev)
    pairs = torch.empty(1, dtype=torch.int32, device=dev)              # zeroed by the entry point
    _lib.call('os3d_strided_tables', x.indices, m, sz, sy, sx, table, cap, out_idx, m_out, oz, oy, ox, bitmap, prefix, fwd,
              inv, pairs, work=lambda: (m + m_out) * (16 + 27 * 4))
    return StridedRulebook(x.indices, [sz, sy, sx], out_idx, [oz, oy, ox], fwd, inv, pairs)
