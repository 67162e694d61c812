"""Run the bf16 tensor-core window attention on real window partitions (8 synthetic frames) for profiling:
    python tools/run_attention.py LEVEL [reps]      LEVEL in 1..4 (C = 48 * 2^(LEVEL-1), 8 heads)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200 import spconv, synthetic  # noqa: E402
from openseg3d_b200.core import voxelize_batch  # noqa: E402
from openseg3d_b200.models import SparseWindowPartitionLayer, WindowAttention  # noqa: E402
from openseg3d_b200.models.segmentors import default_batching_info  # noqa: E402


def main():
    level = int(sys.argv[1])
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    frames = 8
    pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
    coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
    x = spconv.SparseConvTensor(torch.zeros(coors.shape[0], 1, device='cuda'), coors, [64, 1440, 1440], frames)
    for _ in range(level - 1):
        rb = spconv.build_strided_rulebook(x)
        x = spconv.SparseConvTensor(torch.zeros(rb.out_indices.shape[0], 1, device='cuda'), rb.out_indices, rb.out_shape, frames)
    c = 48 * 2 ** (level - 1)
    m = x.indices.shape[0]
    sx, sy, sz = 1440 // 2 ** (level - 1), 1440 // 2 ** (level - 1), 64 // 2 ** (level - 1)
    layer = SparseWindowPartitionLayer(default_batching_info()[level - 1], (10, 10, 8), (sx, sy, sz))
    feats = torch.randn(m, c, device='cuda').bfloat16()
    info = layer(spconv.SparseConvTensor(feats, x.indices, [sz, sy, sx], frames))
    attn = WindowAttention(c, 8, 0.0).cuda().eval()
    seg = info['flat2win_inds_shift0']['segments']
    n_win = int(seg.level_info[13])
    lens = seg.seg_len[:n_win].float()
    flops = 4.0 * float((lens * lens).sum()) * c
    with torch.no_grad():
        for _ in range(2):
            attn(feats, info['pos_dict_shift0'], info['flat2win_inds_shift0'])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            attn(feats, info['pos_dict_shift0'], info['flat2win_inds_shift0'])
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'level {level} tokens {m} windows {n_win} (mean {float(lens.mean()):.1f}, max {int(lens.max())}) C={c}: '
          f'{ms:.3f} ms per WindowAttention call (projections included), useful QK^T+PV {flops / 1e9:.1f} GFLOP')


if __name__ == '__main__':
    main()
