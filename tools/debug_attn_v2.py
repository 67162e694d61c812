"""Multi-window probe of os3d_window_attention_bf16_v2 against a torch restatement (debug tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from openseg3d_b200 import _lib  # noqa: E402


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    for heads, dp in ((8, 16), (8, 32), (8, 48)):
        for kind in ('small', 'mixed', 'big'):
            lens = {'small': rng.integers(1, 40, 200), 'mixed': rng.integers(1, 300, 40), 'big': rng.integers(200, 600, 8)}[kind]
            n = int(lens.sum())
            hd = heads * dp
            starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
            pos_seg = np.stack([np.repeat(starts, lens), np.repeat(lens, lens)], axis=1).astype(np.int32)
            order = rng.permutation(n).astype(np.int32)
            q = F.normalize(torch.randn(n, heads, dp, device='cuda'), dim=-1).reshape(n, hd).bfloat16().contiguous()
            k = F.normalize(torch.randn(n, heads, dp, device='cuda'), dim=-1).reshape(n, hd).bfloat16().contiguous()
            v = torch.randn(n, hd, device='cuda').bfloat16().contiguous()
            li = torch.zeros(16, dtype=torch.int32, device='cuda')
            li[14] = n
            li[13] = len(lens)
            out = torch.zeros((n, hd), dtype=torch.bfloat16, device='cuda')
            tau = torch.tensor([0.5], device='cuda')
            _lib.call('os3d_window_attention_bf16_v2', q, k, v, hd, hd, n, heads, dp, torch.from_numpy(order).cuda(),
                      torch.from_numpy(pos_seg).cuda(), li, tau, 0.01, out, hd)
            torch.cuda.synchronize()
            ref = torch.zeros(n, hd, device='cuda')
            o_t = torch.from_numpy(order).cuda().long()
            for s0, ln in zip(starts, lens):
                rows = o_t[s0:s0 + ln]
                qf, kf, vf = [t[rows].float().reshape(ln, heads, dp).transpose(0, 1) for t in (q, k, v)]
                ref[rows] = ((qf @ kf.transpose(1, 2) / 0.5).softmax(-1) @ vf).transpose(0, 1).reshape(ln, hd)
            err = (out.float() - ref).abs().reshape(n, heads, dp)
            per_head = [round(float(err[:, h].max()), 3) for h in range(heads)]
            # position (in grouped order) of the worst rows
            inv = torch.empty(n, dtype=torch.long, device='cuda')
            inv[o_t] = torch.arange(n, device='cuda')
            bad = (err.amax(dim=(1, 2)) > 0.02).nonzero().flatten()
            bad_pos = sorted(inv[bad].tolist())[:12]
            print(f'H{heads} dp{dp} {kind} n{n}: max err {float(err.max()):.3f} per head {per_head} bad rows {bad.numel()} first bad positions {bad_pos}', flush=True)


if __name__ == '__main__':
    main()
