"""Run the bf16 tensor-core sparse conv on real kernel maps (8 synthetic frames) for profiling:
    python tools/run_spconv.py LEVEL CIN COUT [reps] [flags]      LEVEL in 1..4 (submanifold map of that level);
    flags: 1 = ReLU, 3 = ReLU + pair-summed residual of 2*COUT channels (the UpBlock bottleneck)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200 import spconv, synthetic  # noqa: E402
from openseg3d_b200.core import voxelize_batch  # noqa: E402
from openseg3d_b200.spconv.modules import sparse_conv_forward, _PackedWeights  # noqa: E402


def main():
    level, cin, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    flags = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    frames = 8
    pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
    coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
    x = spconv.SparseConvTensor(torch.zeros(coors.shape[0], 1, device='cuda'), coors, [64, 1440, 1440], frames)
    for _ in range(level - 1):
        rb = spconv.build_strided_rulebook(x)
        x = spconv.SparseConvTensor(torch.zeros(rb.out_indices.shape[0], 1, device='cuda'), rb.out_indices, rb.out_shape, frames)
    rb = spconv.build_subm_rulebook(x)
    m = x.indices.shape[0]
    feats = torch.randn(m, cin, device='cuda').bfloat16()
    w = torch.randn(cout, 3, 3, 3, cin, device='cuda') * 0.02
    cache = _PackedWeights()
    scale, shift = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
    pairs = int((rb.nbr >= 0).sum().item())
    res = torch.randn(m, 2 * cout, device='cuda').bfloat16() if flags & 2 else None
    for _ in range(2):
        sparse_conv_forward(feats, rb.nbr, w, None, cache, scale, shift, res, flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sparse_conv_forward(feats, rb.nbr, w, None, cache, scale, shift, res, flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * pairs * cin * cout
    print(f'level {level} rows {m} pairs {pairs} ({pairs / m:.1f}/row) {cin}->{cout}: {ms:.3f} ms, '
          f'{fl / ms / 1e9:.1f} TFLOP/s algorithmic, {2.0 * m * 27 * cin * cout / ms / 1e9:.1f} TFLOP/s dense-equivalent')


if __name__ == '__main__':
    main()
