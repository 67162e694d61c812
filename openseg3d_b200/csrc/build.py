"""Compile libos3d.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by developers."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'libos3d.so')
SOURCES = ['voxelize.cu', 'scatter.cu', 'rulebook.cu', 'spconv_f32.cu', 'spconv_tc.cu', 'window.cu', 'attention.cu', 'attention_tc.cu', 'attention_v2.cu', 'norm.cu', 'knn.cu', 'linear_tc.cu', 'mlp_tc.cu', 'mlp2_tc.cu', 'qkv_tc.cu', 'labels.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
              '--expt-relaxed-constexpr', '-Xptxas', '-v']


def build(verbose=False, force=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, 'common.cuh'), os.path.join(HERE, '..', '..', 'include', 'os3d.h')]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith('.cuh')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, '_obj'), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, '_obj', os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if not force and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d)
                                                  for d in [s] + [d for d in deps if d.endswith('h')]):
            continue
        procs.append((s, subprocess.Popen([nvcc] + NVCC_FLAGS + ['-c', s, '-o', o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([nvcc, '-shared', '-o', OUT] + objs + ['-lcudart'])
    return OUT


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
