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
