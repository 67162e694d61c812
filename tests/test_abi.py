"""CPU-side checks of the drop-in boundary: libos3d.so loads and exports every symbol include/os3d.h declares
(no compute is run here), and the product package never reaches into oracle/."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'os3d.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(os3d_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from openseg3d_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'build libos3d.so first: python openseg3d_b200/csrc/build.py'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/os3d.h but not exported'
    # and the ctypes table binds exactly the declared entry points
    assert sorted(_lib.exported_symbols()) == names


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'openseg3d_b200')
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|libos3d_oracle|oracle[./]oracle', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f'{f} reaches into oracle/: the product path must not depend on it'


def test_no_cpu_fallback():
    import pytest
    import torch
    from openseg3d_b200.ops import voxel_to_point
    with pytest.raises(RuntimeError):
        voxel_to_point(torch.zeros(4, 8), torch.zeros(3, dtype=torch.long))


def test_dense_kernel_planners_on_cpu():
    """The host-side planners that decide which dense tcgen05 kernel takes a layer (shared-memory / tensor-memory
    budgets) are plain C and run without a GPU: the widths of the reference model must land where DESIGN.md says."""
    from openseg3d_b200 import _lib
    from openseg3d_b200.ops.mlp_chain import MlpChain, SwformerMlp
    L = _lib.lib()
    # persistent Linear: N <= 256 with the weights resident; the level-4 width does not fit
    assert L.os3d_linear_tc_fits(48, 48) and L.os3d_linear_tc_fits(192, 192) and L.os3d_linear_tc_fits(96, 256)
    assert not L.os3d_linear_tc_fits(384, 384) and not L.os3d_linear_tc_fits(96, 40) and not L.os3d_linear_tc_fits(0, 64)
    # SWFormer MLP with streamed weights: levels 1-3 (C = 48, 96, 192, hidden 2C); level 4 is out
    assert SwformerMlp.fits(48, 96) and SwformerMlp.fits(96, 192) and SwformerMlp.fits(192, 384)
    assert not SwformerMlp.fits(384, 768) and not SwformerMlp.fits(50, 100) and not SwformerMlp.fits(96, 200)
    # resident-weight chains: Segformer's point MLPs (segformer.py:21-32,58-76) fit, with and without the fp32 front layer
    assert MlpChain.fits([(128, 64), (256, 128), (64, 256)], front=True)
    assert MlpChain.fits([(256, 96), (128, 256), (64, 128)])
    assert MlpChain.fits([(64, 64), (22, 64)])
    assert not MlpChain.fits([(64, 64)])                              # one layer: os3d_linear_tc_bf16
    assert not MlpChain.fits([(512, 64), (64, 512)])                  # wider than a tensor-memory accumulator
    assert not MlpChain.fits([(384, 192), (192, 384)])                # weights beyond shared memory
    assert not MlpChain.fits([(128, 64), (64, 96)])                   # widths do not chain


def test_tile_order_simulation_uses_the_kernels_bit_order():
    """tools/sim_tile_order.py (the CPU simulation behind the conv tile order) and order_keys_kernel rank the 27 offsets
    identically, and the ranking is a permutation that keeps the symmetric pairs (k, 26 - k) adjacent in significance."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = open(os.path.join(root, 'openseg3d_b200', 'csrc', 'spconv_tc.cu')).read()
    sim = open(os.path.join(root, 'tools', 'sim_tile_order.py')).read()
    k_cu = [int(x) for x in re.search(r'kOrderBit\[OS3D_KVOL\] = \{([^}]*)\}', cu).group(1).split(',')]
    k_sim = [int(x) for x in re.search(r'ORDER=\[([^\]]*)\]', sim).group(1).split(',')]
    assert k_cu == k_sim and sorted(k_cu) == list(range(27)) and k_cu[0] == 13
    pos = {k: i for i, k in enumerate(k_cu)}
    assert all(abs(pos[k] - pos[26 - k]) <= 7 for k in range(27))          # a pair sits in the same frequency class
