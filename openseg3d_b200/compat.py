"""Import swap: make an UNMODIFIED OpenSeg3D checkout run its hot path on libos3d.

    import openseg3d_b200.compat as compat
    compat.install()                 # before anything imports seg3d.models / seg3d.ops
    from seg3d.models.segmentors import Segformer          # the reference's own class, on the new ops

``install()`` registers, in ``sys.modules``:
  * ``spconv`` / ``spconv.pytorch``  -> openseg3d_b200.spconv           (seg3d/utils/spconv_utils.py:1,
                                                                         seg3d/models/backbones/pointtransformer.py:6)
  * ``torch_scatter``                -> a module whose ``scatter(src, index, dim=0, reduce=)`` is os3d_scatter_{max,mean}
                                        (seg3d/models/voxel_encoders/vfe.py:3, seg3d/models/layers/se_layer.py:3)
  * ``seg3d.ops``                    -> openseg3d_b200.ops              (seg3d/ops/__init__.py:1-6; the reference package
                                        imports four compiled torch extensions that do not exist for sm_100)
  * ``easydict``                     -> a minimal EasyDict, only when the real package is absent (seg3d/utils/config.py:2)
With ``layers=True`` (default) the reference's own window-partition / attention classes are replaced by the
variable-length ones as well (``seg3d.models.layers.point_transformer_layer``, ``seg3d.models.layers.cosine_msa``);
with ``layers=False`` the reference's padded PyTorch layer code runs unchanged on top of the new spconv / scatter /
``get_inner_win_inds`` ops.
"""
import importlib
import sys
import types


def _torch_scatter_module():
    from .ops.pooling import scatter_max, scatter_mean
    mod = types.ModuleType('torch_scatter')

    def scatter(src, index, dim=0, out=None, dim_size=None, reduce='sum'):
        """torch_scatter.scatter for the two reductions the reference uses, along dim 0 of a 2-D tensor."""
        if dim != 0 or out is not None or src.dim() != 2:
            raise NotImplementedError('openseg3d_b200 torch_scatter shim: scatter(src [N, C], index [N], dim=0) only')
        if reduce == 'max':
            return scatter_max(src, index, dim_size).to(src.dtype)
        if reduce == 'mean':
            return scatter_mean(src, index, dim_size).to(src.dtype)
        raise NotImplementedError(f'openseg3d_b200 torch_scatter shim: reduce={reduce!r} (the reference uses max / mean)')

    mod.scatter = scatter
    mod.__doc__ = 'openseg3d_b200 shim of torch_scatter (scatter max / mean over libos3d)'
    return mod


def _easydict_module():
    mod = types.ModuleType('easydict')

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    mod.EasyDict = EasyDict
    return mod


def install(layers=True):
    """Register the aliases.  Idempotent.  Call before the first ``import seg3d.models`` / ``import seg3d.ops``."""
    from . import ops, spconv
    sys.modules['spconv'] = sys.modules['spconv.pytorch'] = spconv
    spconv.pytorch = spconv                  # `import spconv.pytorch as spconv` resolves the attribute too
    if 'torch_scatter' not in sys.modules or getattr(sys.modules['torch_scatter'], '__doc__', '') and \
            'openseg3d_b200' in (sys.modules['torch_scatter'].__doc__ or ''):
        sys.modules['torch_scatter'] = _torch_scatter_module()
    try:
        importlib.import_module('easydict')
    except ImportError:
        sys.modules['easydict'] = _easydict_module()
    # seg3d.ops: the package __init__ imports compiled extensions; register ours under its name (and the submodule
    # names the reference imports from) before that __init__ can run
    sys.modules['seg3d.ops'] = ops
    if layers:
        import seg3d.models.layers.point_transformer_layer as ptl          # the reference module object
        import seg3d.models.layers.cosine_msa as cmsa
        from .models import layers as ours
        for name in ('SparseWindowPartitionLayer', 'WindowAttention', 'MLP', 'EncoderLayer', 'SWFormerBlock'):
            setattr(ptl, name, getattr(ours, name))
        cmsa.CosineMultiheadAttention = ours.CosineMultiheadAttention
        import seg3d.models.layers as pkg
        pkg.SparseWindowPartitionLayer, pkg.WindowAttention = ours.SparseWindowPartitionLayer, ours.WindowAttention
        for modname in ('seg3d.models.backbones.pointtransformer',):
            m = sys.modules.get(modname)
            if m is not None:                    # already imported: rebind the names it copied
                m.SparseWindowPartitionLayer, m.SWFormerBlock = ours.SparseWindowPartitionLayer, ours.SWFormerBlock
