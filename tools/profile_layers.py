"""Per-call CUDA-event timing of one bench step (libos3d entry points), with shapes and achieved rates.
    python tools/profile_layers.py [--dtype bf16] [--frames 8]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from openseg3d_b200 import _lib, synthetic  # noqa: E402
from openseg3d_b200.models import build_segformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--frames', type=int, default=8)
    args = ap.parse_args()
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    model = build_segformer('waymo_one_sweep', compute_dtype=dtype).cuda().eval()
    pts, _ = synthetic.make_batch(list(range(args.frames)), 1, False)
    dev = torch.from_numpy(pts).cuda()
    for _ in range(3):
        with torch.no_grad():
            model({'points': dev, 'batch_size': args.frames})
    records = []
    orig = _lib.call

    def wrapped(name, *a, work=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        ints = [x for x in a if isinstance(x, int) and not isinstance(x, bool) and x < (1 << 40)]
        records.append((name, e0, e1, ints, work))

    _lib.call = wrapped
    for mod in list(sys.modules.values()):
        if mod is not None and getattr(mod, '_lib', None) is _lib:
            pass
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    with torch.no_grad():
        model({'points': dev, 'batch_size': args.frames})
    t1.record()
    torch.cuda.synchronize()
    _lib.call = orig
    total = 0.0
    for name, e0, e1, ints, work in records:
        ms = e0.elapsed_time(e1)
        total += ms
        extra = ''
        if work is not None:
            w = float(work())
            extra = f'  {w / 1e9:9.1f} GFLOP  {w / (ms * 1e-3) / 1e12:7.1f} TFLOP/s'
        print(f'{name:32s} {ms:8.3f} ms  {str(ints[:6]):40s}{extra}')
    print(f'libos3d calls: {total:.2f} ms; whole step (with event overhead): {t0.elapsed_time(t1):.2f} ms')


if __name__ == '__main__':
    main()
