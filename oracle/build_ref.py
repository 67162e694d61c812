"""Test infrastructure only: compile the part of the REFERENCE that builds from its own sources -- the CPU functions of
seg3d/ops/voxel_pooling/src/voxel_pooling.cpp -- where they lie under /root/reference, into oracle/_ref/ (git-ignored).
Nothing is copied into the repo.  Used to pin oracle.voxel_avg_pooling and to generate tests/golden/voxel_avg_pooling.npz
(tests/golden/make_golden_pooling.py).  Everything else on the path is Python (imported directly for the goldens) or
lives in absent third-party wheels (spconv, torch_scatter): DESIGN.md section 4.

    python oracle/build_ref.py        ->  oracle/_ref/voxel_pooling_ext*.so   (skipped when /root/reference is absent)
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/seg3d/ops/voxel_pooling/src/voxel_pooling.cpp'
OUT_DIR = os.path.join(HERE, '_ref')
NAME = 'voxel_pooling_ext'


def out_path():
    return os.path.join(OUT_DIR, NAME + sysconfig.get_config_var('EXT_SUFFIX'))


def build(force=False):
    """Returns the path of the built extension, or None when the reference sources are not there (GPU box)."""
    out = out_path()
    if not os.path.exists(REF_SRC):
        return out if os.path.exists(out) else None
    stubs = os.path.join(HERE, 'ref_stubs.cpp')
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(REF_SRC), os.path.getmtime(stubs)):
        return out
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f'-I{p}' for p in ce.include_paths()] + [f'-I{sysconfig.get_paths()["include"]}',
                                                     f'-I{os.path.dirname(REF_SRC)}']
    libdir = ce.library_paths()[0]
    cmd = ['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-fopenmp', f'-DTORCH_EXTENSION_NAME={NAME}', '-DTORCH_API_INCLUDE_EXTENSION_H',
           '-D_GLIBCXX_USE_CXX11_ABI=' + str(int(__import__('torch')._C._GLIBCXX_USE_CXX11_ABI)),
           *inc, REF_SRC, stubs, '-o', out, f'-L{libdir}', '-ltorch', '-ltorch_cpu', '-lc10', '-ltorch_python',
           f'-Wl,-rpath,{libdir}']
    subprocess.check_call(cmd)
    return out


def load():
    """Import the built reference extension (None if it cannot exist here)."""
    path = build()
    if path is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
