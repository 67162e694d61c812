"""Which python lines issue torch (aten) ops in one bf16 inference forward:  python tools/trace_ops.py [config]
A TorchDispatchMode counts every dispatched aten op that is not a pure view / allocation, grouped by the innermost
openseg3d_b200 source line on the python stack -- the list of "glue" kernels to remove."""
import collections
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.utils._python_dispatch import TorchDispatchMode  # noqa: E402

from openseg3d_b200 import synthetic  # noqa: E402
from openseg3d_b200.models import build_segformer  # noqa: E402

NO_KERNEL = ('view', 'reshape', 'slice', 'select', 'transpose', 'permute', 'detach', 'expand', 'as_strided', 'empty', 't.default',
             'unsqueeze', 'squeeze', 'alias', 'size', 'stride', 'is_', 'numel', 'dim', '_unsafe_view', 'unbind', 'split', 'chunk',
             'narrow', 'storage_offset', 'sym_', 'lift_fresh', 'resolve_', '_local_scalar_dense', 'item', 'prim', 'record_stream',
             'is_pinned', 'result_type', 'can_cast')


class Tracer(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.counts = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func).replace('aten.', '')
        if not any(k in name for k in NO_KERNEL):
            site = '?'
            for fr in reversed(traceback.extract_stack(limit=30)):
                if 'openseg3d_b200' in fr.filename and 'trace_ops' not in fr.filename:
                    site = f"{fr.filename.split('openseg3d_b200/')[-1]}:{fr.lineno} {fr.line.strip()[:70]}"
                    break
            self.counts[(name, site)] += 1
        return func(*args, **(kwargs or {}))


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else 'waymo_one_sweep'
    frames = 8
    model = build_segformer(config, compute_dtype=torch.bfloat16).cuda().eval()
    from openseg3d_b200.models.segmentors import DATASET_CONFIGS
    c = DATASET_CONFIGS[config]
    pts, _ = synthetic.make_batch(list(range(frames)), c['num_sweeps'], c['use_cylinder'])
    dev = torch.from_numpy(pts).cuda()
    with torch.no_grad():
        for _ in range(2):
            model({'points': dev, 'batch_size': frames})
        torch.cuda.synchronize()
        tr = Tracer()
        with tr:
            model({'points': dev, 'batch_size': frames})
        torch.cuda.synchronize()
    print('aten ops that may launch a kernel:', sum(tr.counts.values()))
    for (name, site), n in sorted(tr.counts.items(), key=lambda kv: -kv[1])[:70]:
        print(f'{n:5d}  {name:28s} {site}')


if __name__ == '__main__':
    main()
