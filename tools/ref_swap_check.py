"""Run the REFERENCE's own model files on openseg3d_b200 through the import swap of INTEGRATION.md §A
(openseg3d_b200.compat.install) and compare with openseg3d_b200.models.Segformer over the same state_dict.

    python tools/ref_swap_check.py --layers {0,1} --device {cpu,cuda}

Needs a copy of the reference's `seg3d` package: baseline/_ref/seg3d (git-ignored; __graft_entry__.build() copies it from
/root/reference when that exists) or /root/reference itself.  --device cpu only builds both models and checks that
the state_dict keys and shapes are identical (strict load); --device cuda also runs a small synthetic batch through
both and prints the max-norm relative differences.  Prints one JSON line.
  --layers 0: the reference's padded window-partition / attention Python code runs unchanged on the new spconv / scatter /
              get_inner_win_inds ops;  --layers 1: those classes are swapped for the variable-length ones too."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_root():
    for cand in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(cand, 'seg3d', 'models')):
            return cand
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--layers', type=int, default=1)
    ap.add_argument('--device', default='cuda')
    args = ap.parse_args()
    ref_root = reference_root()
    if ref_root is None:
        print(json.dumps({'skipped': 'no copy of the reference seg3d package (baseline/_ref or /root/reference)'}))
        return
    sys.path.insert(0, ref_root)
    import torch
    import openseg3d_b200.compat as compat
    compat.install(layers=bool(args.layers))
    from seg3d.models.segmentors.segformer import Segformer as RefSegformer          # the reference's own class
    import seg3d.models.backbones.pointtransformer as ref_pt
    from openseg3d_b200 import synthetic
    from openseg3d_b200.core.voxel import voxelize_batch
    from openseg3d_b200.models import Segformer, layers as ours_layers
    from openseg3d_b200.models.segmentors import dataset_spec, default_batching_info

    assert os.path.abspath(ref_pt.__file__).startswith(os.path.abspath(ref_root)), ref_pt.__file__
    depths = [2, 2, 2, 2]
    ds = dataset_spec('waymo_one_sweep')
    torch.manual_seed(0)
    ours = Segformer(ds, default_batching_info(), [10, 10, 8], depths, 0.3).eval()
    torch.manual_seed(1)
    ref = RefSegformer(ds, default_batching_info(), [10, 10, 8], depths, 0.3).eval()
    out = {'layers': args.layers, 'device': args.device, 'reference_root': ref_root,
           'reference_layer_class': type(ref.point_transformer.swformer_block1[1]).__module__}
    swapped = type(ref.point_transformer.swformer_block1[1]) is ours_layers.SWFormerBlock
    assert swapped == bool(args.layers)
    sd = ours.state_dict()
    ref_sd = ref.state_dict()
    out['keys_equal'] = sorted(sd) == sorted(ref_sd)
    out['shapes_equal'] = out['keys_equal'] and all(tuple(sd[k].shape) == tuple(ref_sd[k].shape) for k in sd)
    ref.load_state_dict(sd, strict=True)                      # a checkpoint of one loads into the other unchanged
    out['n_keys'] = len(sd)
    if args.device == 'cuda':
        ours, ref = ours.cuda(), ref.cuda()
        with torch.no_grad():
            g = torch.Generator().manual_seed(1)
            for m in (ours, ref):
                g.manual_seed(1)
                for n, b in m.named_buffers():
                    if n.endswith('running_mean'):
                        b.copy_(0.1 * torch.randn(b.shape, generator=g))
                    elif n.endswith('running_var'):
                        b.copy_(torch.empty(b.shape).uniform_(0.5, 1.5, generator=g))
        pts, offs = synthetic.make_batch([0, 1], 1, False, 16, 300)
        dev = torch.from_numpy(pts).cuda()
        coords, pvid = voxelize_batch(dev, ds.voxel_size, ds.point_cloud_range, has_batch=True)

        def batch():          # what collate_batch + load_data_to_gpu hand the model (waymo_dataset.py:339-376, data_utils.py:6-15)
            return {'points': dev.clone(), 'voxel_coords': coords.float(), 'point_voxel_ids': pvid.clone(), 'batch_size': 2,
                    'point_id_offset': torch.from_numpy(offs).float().cuda()}
        with torch.no_grad():
            a = ours(batch())
            b = ref(batch())
        torch.cuda.synchronize()
        for k in ('point_out', 'voxel_out', 'aux_voxel_out'):
            out[f'rel_{k}'] = float((a[k].float() - b[k].float()).abs().max() / b[k].float().abs().max())
        out['coords_equal'] = bool((a['aux_voxel_coords'] == b['aux_voxel_coords']).all())
        out['argmax_agreement'] = float((a['point_out'].argmax(1) == b['point_out'].argmax(1)).float().mean())
    print(json.dumps(out))


if __name__ == '__main__':
    main()
