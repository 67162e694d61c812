"""Kernel-only timing of the two tensor-core window-attention kernels on the real window partitions of 8 synthetic frames,
per pyramid level, and the difference between their outputs:
    python tools/attn_bench.py [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200 import _lib, spconv, synthetic  # noqa: E402
from openseg3d_b200.core import voxelize_batch  # noqa: E402
from openseg3d_b200.models import SparseWindowPartitionLayer, WindowAttention  # noqa: E402
from openseg3d_b200.models import layers as lay  # noqa: E402
from openseg3d_b200.models.segmentors import default_batching_info  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    frames = 8
    pts, _ = synthetic.make_batch(list(range(frames)), 1, False)
    coors, _ = voxelize_batch(torch.from_numpy(pts).cuda(), [0.1, 0.1, 0.1], [-72, -72, -2, 72, 72, 4.4])
    x = spconv.SparseConvTensor(torch.zeros(coors.shape[0], 1, device='cuda'), coors, [64, 1440, 1440], frames)
    for level in range(1, 5):
        if level > 1:
            rb = spconv.build_strided_rulebook(x)
            x = spconv.SparseConvTensor(torch.zeros(rb.out_indices.shape[0], 1, device='cuda'), rb.out_indices, rb.out_shape, frames)
        c = 48 * 2 ** (level - 1)
        m = x.indices.shape[0]
        sx, sy, sz = 1440 // 2 ** (level - 1), 1440 // 2 ** (level - 1), 64 // 2 ** (level - 1)
        layer = SparseWindowPartitionLayer(default_batching_info()[level - 1], (10, 10, 8), (sx, sy, sz))
        torch.manual_seed(level)
        feats = torch.randn(m, c, device='cuda').bfloat16()
        info = layer(spconv.SparseConvTensor(feats, x.indices, [sz, sy, sx], frames))
        attn = WindowAttention(c, 8, 0.0).cuda().eval()
        for shift in (0, 1):
            seg = info[f'flat2win_inds_shift{shift}']['segments']
            n_win = int(seg.level_info[13])
            lens = seg.seg_len[:n_win].float()
            flops = 4.0 * float((lens * lens).sum()) * c
            outs, line = {}, []
            for impl in ('v1', 'v2'):
                lay._ATTN_IMPL = impl
                with torch.no_grad():
                    for _ in range(2):
                        o = attn(feats, info[f'pos_dict_shift{shift}'], info[f'flat2win_inds_shift{shift}'])
                    prof = []
                    _lib.PROFILE = prof
                    for _ in range(reps):
                        o = attn(feats, info[f'pos_dict_shift{shift}'], info[f'flat2win_inds_shift{shift}'])
                    torch.cuda.synchronize()
                    _lib.PROFILE = None
                outs[impl] = o.float()
                by = {}
                for name, e0, e1, w in prof:
                    by[name] = by.get(name, 0.0) + e0.elapsed_time(e1) / reps
                t_k = sum(v for n, v in by.items() if n.startswith('os3d_window_attention_bf16_' + ('tc' if impl == 'v1' else 'v2')))
                line.append(f'{impl}: attn {t_k:.3f} ms ({flops / 1e9 / max(t_k, 1e-9):.0f} TFLOP/s useful)'
                            + f' proj {by.get("os3d_wide_linear_bf16", 0.0):.3f}')
            diff = (outs['v1'] - outs['v2']).abs().max().item() / outs['v1'].abs().max().item()
            print(f'L{level} shift{shift} tokens {m} windows {n_win} (mean {float(lens.mean()):.1f}, max {int(lens.max())}) C={c} '
                  f'useful {flops / 1e9:.1f} GFLOP | ' + ' | '.join(line) + f' | max rel diff v1-v2 {diff:.2e}', flush=True)


if __name__ == '__main__':
    main()
