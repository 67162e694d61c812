import torch


class SparseConvTensor(object):
    """features [M, C], indices [M, 4] int32 (batch, z, y, x), spatial_shape [Z, Y, X], batch_size.

    ``indice_dict`` caches kernel maps by ``indice_key`` exactly as the reference's modules expect (a rulebook is
    built by the first conv that names a key and reused by every later conv with that key, including the
    SparseInverseConv3d that swaps it); ``replace_feature`` returns a new tensor sharing that cache."""

    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None, _site_table=None):
        if indices.dtype != torch.int32:
            indices = indices.int()
        self.features = features
        self.indices = indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self._site_table = _site_table if _site_table is not None else {}      # id(indices) -> hash table

    def replace_feature(self, feature):
        return SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.indice_dict,
                                self._site_table)

    def find_indice_pair(self, key):
        return None if key is None else self.indice_dict.get(key)

    @property
    def spatial_size(self):
        n = 1
        for s in self.spatial_shape:
            n *= s
        return n

    def dense(self, channels_first=True):
        z, y, x = self.spatial_shape
        out = torch.zeros((self.batch_size, z, y, x, self.features.shape[1]), dtype=self.features.dtype,
                          device=self.features.device)
        i = self.indices.long()
        out[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out
