"""Probe os3d_window_attention_bf16_v2 with crafted single-window inputs against a torch restatement (debug tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from openseg3d_b200 import _lib  # noqa: E402


def run(n, heads, dp, q, k, v, tau=0.5, entry='os3d_window_attention_bf16_v2'):
    dev = 'cuda'
    hd = heads * dp
    order = torch.arange(n, dtype=torch.int32, device=dev)
    pos_seg = torch.tensor([[0, n]] * n, dtype=torch.int32, device=dev).contiguous()
    li = torch.zeros(16, dtype=torch.int32, device=dev)
    li[14] = n
    li[13] = 1
    out = torch.zeros((n, hd), dtype=torch.bfloat16, device=dev)
    t = torch.tensor([tau], dtype=torch.float32, device=dev)
    _lib.call(entry, q, k, v, hd, hd, n, heads, dp, order, pos_seg, li, t, 0.01, out, hd)
    torch.cuda.synchronize()
    return out.float()


def ref(n, heads, dp, q, k, v, tau=0.5):
    qf, kf, vf = [t.float().reshape(n, heads, dp).transpose(0, 1) for t in (q, k, v)]
    s = qf @ kf.transpose(1, 2) / tau
    return (s.softmax(-1) @ vf).transpose(0, 1).reshape(n, heads * dp)


def main():
    torch.manual_seed(0)
    for heads, dp in ((8, 16), (8, 32), (8, 48)):
        for n in (64, 100, 300):
            for case in ('qk_lo', 'qk_hi', 'full'):
                hd = heads * dp
                q = torch.randn(n, heads, dp, device='cuda')
                k = torch.randn(n, heads, dp, device='cuda')
                v = torch.randn(n, heads, dp, device='cuda')
                if case == 'qk_lo':
                    q[:, :, 8:] = 0
                    k[:, :, 8:] = 0
                elif case == 'qk_hi':
                    q[:, :, :8] = 0
                    k[:, :, :8] = 0
                q = F.normalize(q, dim=-1).reshape(n, hd).bfloat16().contiguous()
                k = F.normalize(k, dim=-1).reshape(n, hd).bfloat16().contiguous()
                v = v.reshape(n, hd).bfloat16().contiguous()
                o = run(n, heads, dp, q, k, v)
                r = ref(n, heads, dp, q, k, v)
                err = (o - r).abs().reshape(n, heads, dp)
                per_chunk = [round(float(err[:, :, c * 8:(c + 1) * 8].max()), 3) for c in range(dp // 8)]
                per_head = [round(float(err[:, h].max()), 3) for h in range(heads)]
                print(f'H{heads} dp{dp} n{n} {case}: max err {float(err.max()):.3f} per out-chunk {per_chunk} per head {per_head}', flush=True)


if __name__ == '__main__':
    main()
