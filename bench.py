#!/usr/bin/env python
"""bench.py -- points/sec of the OpenSeg3D voxel-backbone hot path (Segformer forward) on B200.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (the CPU oracle on the host cores, same metric and config)
  python bench.py --mode train ...                          (BASELINE configs[4]: bf16 train step, DDP over NCCL)

Workload (BASELINE.json configs[1]): configs/waymo_one_sweep.yaml inference, a batch of 8 synthetic Waymo-shape
frames per GPU (~181k points each), random-init Segformer (PointTransformer backbone), bf16 backbone.
A step = voxelize -> point MLP -> scatter-max -> sparse UNet with window attention -> gather -> point head.
Frames are sharded over ranks (weak scaling, no data-path collective: inference has none).

JSON line keys follow the driver contract; `value` has the points already in HBM, `e2e` starts from pinned host
memory and ends with the predicted labels back on the host.  The same line carries, under `other_configs`, the fp32
(exact-parity) mode of the headline workload and the cylinder / multi-sweep workloads (BASELINE configs[2], [3]).

The reference arm (`--impl reference`) times the CPU restatement of the reference path (oracle/, fp32, every host
thread) on ONE full frame of the same workload per step -- BASELINE configs[0] -- and reports it under the GPU arm's
`config`, as the contract asks; the sample is described in `cpu_baseline.sample` and `sample`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 8
CONFIG = 'waymo_one_sweep'
POINTS_PER_FRAME_NOTE = '64 beams x 2650 columns, seeds rank*frames .. +frames-1'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tensor=p.get('bf16_tflops_sustained', p['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, tensor=1400.0, src='fallback')


def workload_config(config, frames, n_points, world):
    """The `config` object of the JSON line -- built by ONE function so that the GPU arm and the reference arm report the
    same workload (the reference arm's bounded per-step sample is stated beside it, not inside it)."""
    return {'workload': f'configs/{config}.yaml inference, batch of {frames} synthetic frames per GPU '
                        f'({n_points} points/GPU), voxelize + sparse UNet + window attention, random-init',
            'frames_per_gpu': frames, 'points_per_gpu': n_points, 'parallelism': f'frames sharded x{world}',
            'l2': 'per-step working set (activations > 2 GB) >> 126 MB L2; no explicit flush'}


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  Read through NVML in-process
    (nvidia_ml_py): a `nvidia-smi` subprocess per sample initialises the driver's management library every time and
    was seen to stall kernel launches of the timed steps for milliseconds (one step in ten at 46-50 ms instead of
    42.5); the subprocess query stays as the fallback when NVML cannot be loaded."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(int(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        act = lambda mask: 'Active' if bits & mask else 'Not Active'      # noqa: E731
        return [str(sm), str(self.max_sm), act(0x8), act(0x40), act(0x20), act(0x4)]    # hw, hw thermal, sw thermal, sw power cap

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(',')]
                    if len(f) == 6:
                        self.rows.append(f)
            except Exception:
                self.nvml = None              # NVML failed mid-run: fall back to the subprocess query
            time.sleep(0.05 if self.nvml is not None else 0.1)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons,
                'samples': len(self.rows), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's forward (the CPU restatement of the reference path) on ONE FULL frame of the workload
# ------------------------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_forward_sample(threads, config=CONFIG):
    """One full frame (frame 0 of the batch: BASELINE configs[0]) through oracle.segformer_forward, fp32, `threads` host
    threads.  Returns (points_per_sec, n_points, seconds)."""
    import torch
    from openseg3d_b200 import synthetic
    from openseg3d_b200.models import build_segformer
    from openseg3d_b200.models.segmentors import default_batching_info, DATASET_CONFIGS
    from oracle import oracle
    torch.set_num_threads(threads)
    c = DATASET_CONFIGS[config]
    if config not in _CPU_STATE:
        sd = build_segformer(config).eval().state_dict()
        pts, _ = synthetic.make_batch([0], c['num_sweeps'], c['use_cylinder'])
        _CPU_STATE[config] = (sd, pts)
    sd, pts = _CPU_STATE[config]
    t0 = time.perf_counter()
    with torch.no_grad():
        oracle.segformer_forward(sd, pts, c['voxel_size'], c['point_cloud_range'], default_batching_info(), [10, 10, 8],
                                 [3, 4, 8, 3], multi_sweeps=c['use_multi_sweeps'])
    dt = time.perf_counter() - t0
    return pts.shape[0] / dt, pts.shape[0], dt


def cpu_sample_text(n, dt=None):
    s = (f'1 full frame of the batch per step (frame seed 0: {n} points, BASELINE configs[0]), fp32, oracle port of the '
         'reference path (numba-equivalent C voxelizer + torch-CPU restatements)')
    return s + (f', {dt:.1f} s' if dt is not None else '')


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path is not runnable (spconv / torch_scatter
    absent, GPU-only extensions), so this times the oracle port on the host cores -- DESIGN.md "Measurement"."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from oracle import oracle
    from openseg3d_b200 import synthetic
    from openseg3d_b200.models.segmentors import DATASET_CONFIGS
    from openseg3d_b200.utils.sharding import frame_seeds
    oracle.lib()
    vals = []
    for i in range(args.warmup + args.steps):
        pps, n, dt = cpu_forward_sample(cores, args.config)
        if i >= args.warmup:
            vals.append((pps, dt))
    pps = sum(v[0] for v in vals) / len(vals)
    ms = 1e3 * sum(v[1] for v in vals) / len(vals)
    sample = cpu_sample_text(n)
    dcfg = DATASET_CONFIGS[args.config]
    n_batch = synthetic.make_batch(frame_seeds(0, args.frames), dcfg['num_sweeps'], dcfg['use_cylinder'])[0].shape[0]
    line = {'impl': 'reference', 'metric': 'points/sec, Waymo 1-sweep seg forward', 'value': pps, 'unit': 'points/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.config, args.frames, n_batch, max(args.gpus, 1)),
            'sample': sample, 'host_threads': cores,
            'cpu_baseline': {'value': pps, 'unit': 'points/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': pps, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Arm(object):
    """One (config, dtype) workload on this rank: model, pinned host points, resident device points."""

    def __init__(self, config, dtype_name, frames, rank):
        import torch
        from openseg3d_b200 import synthetic
        from openseg3d_b200.models import build_segformer
        from openseg3d_b200.models.segmentors import DATASET_CONFIGS
        from openseg3d_b200.utils.sharding import frame_seeds
        self.torch = torch
        self.config, self.dtype_name, self.frames = config, dtype_name, frames
        self.dtype = torch.bfloat16 if dtype_name == 'bf16' else torch.float32
        self.model = build_segformer(config, compute_dtype=self.dtype).cuda().eval()
        dcfg = DATASET_CONFIGS[config]
        pts_np, _ = synthetic.make_batch(frame_seeds(rank, frames), dcfg['num_sweeps'], dcfg['use_cylinder'])
        self.n_points = pts_np.shape[0]
        self.host = torch.from_numpy(pts_np).pin_memory()
        self.dev_pts = self.host.cuda(non_blocking=True)
        self.labels_host = None

    def step_resident(self):
        with self.torch.no_grad():
            return self.model({'points': self.dev_pts, 'batch_size': self.frames})['point_out']

    def step_e2e(self):
        torch = self.torch
        from openseg3d_b200.ops import predict_labels
        with torch.no_grad():
            d = self.host.cuda(non_blocking=True)
            out = self.model({'points': d, 'batch_size': self.frames})['point_out']
            self.labels_host.copy_(predict_labels(out), non_blocking=True)   # tools/test.py:58 argmax + .cpu()
        torch.cuda.current_stream().synchronize()

    def warm(self, n):
        for _ in range(n):
            n_labelled = self.step_resident().shape[0]      # multi-sweep: logits only for the current sweep's points
        self.n_labelled = n_labelled
        self.labels_host = self.torch.empty(n_labelled, dtype=self.torch.uint8).pin_memory()


def kernel_profile(arm):
    """Per-entry-point CUDA-event times of one more step (events on the launching stream) -> {name: [ms, flops, n, bytes]}
    plus the list of (name, ms, work) records."""
    import torch
    from openseg3d_b200 import _lib
    prof = []
    _lib.PROFILE = prof
    arm.step_resident()
    torch.cuda.synchronize()
    _lib.PROFILE = None
    by, recs = {}, []
    for name, e0, e1, work in prof:
        ms = e0.elapsed_time(e1)
        d = by.setdefault(name, [0.0, 0.0, 0, 0.0])
        d[0] += ms
        d[1] += work
        d[2] += 1
        d[3] += getattr(work, 'bytes', 0.0)
        recs.append((name, ms, work))
    return by, recs


def conv_layer_table(recs, pk):
    """Per-layer table of the sparse convolutions of one step: launches grouped by (Cin, Cout, output rows)."""
    rows = {}
    for name, ms, work in recs:
        d = getattr(work, 'detail', None)
        if not d or 'cin' not in d:
            continue
        key = (d['cin'], d['cout'], d['m_out'])
        r = rows.setdefault(key, dict(cin=d['cin'], cout=d['cout'], rows=d['m_out'], launches=0, ms=0.0, gflop=0.0,
                                      algorithmic_mb=0.0, pairs_per_row=d['pairs'] / max(d['m_out'], 1)))
        r['launches'] += 1
        r['ms'] += ms
        r['gflop'] += float(work) / 1e9
        r['algorithmic_mb'] += getattr(work, 'bytes', 0.0) / 1e6
    out = []
    for key in sorted(rows):
        r = rows[key]
        tf = r['gflop'] / r['ms'] if r['ms'] else 0.0            # GFLOP / ms = TFLOP/s
        gbs = r['algorithmic_mb'] / r['ms'] if r['ms'] else 0.0  # MB / ms = GB/s
        r.update(ms=round(r['ms'], 3), gflop=round(r['gflop'], 1), algorithmic_mb=round(r['algorithmic_mb'], 1),
                 pairs_per_row=round(r['pairs_per_row'], 2), tflops=round(tf, 1), frac_of_tensor_peak=round(tf / pk['tensor'], 3),
                 algorithmic_gbs=round(gbs, 1), frac_of_hbm_peak=round(gbs / pk['hbm'], 3))
        out.append(r)
    return out


def build_roofline(arm, by, recs, step_ms, pk):
    import torch
    dtype = arm.dtype
    # The dominant kernel family is spconv_tc_kernel: the sparse convolutions (os3d_spconv_fwd_bf16) and, in its dense
    # mode, the level-4 LayerNorm-fused Linear layers (os3d_linear_bf16).  achieved = algorithmic FLOPs of all its
    # launches in a step (2 * pairs * Cin * Cout per conv, recomputed from the live kernel maps; 2 * M * K * N per
    # Linear) / their CUDA-event time.  traffic = DRAM bytes per launch from the committed ncu launch list.
    if dtype == torch.bfloat16:
        entries = ['os3d_spconv_fwd_bf16_ld', 'os3d_spconv_fwd_bf16', 'os3d_linear_bf16']
        kname = 'spconv_tc_kernel (tcgen05 gather-GEMM: sparse conv + LayerNorm-fused Linear)'
    else:
        entries = ['os3d_spconv_fwd_f32']
        kname = 'spconv_f32_kernel (FP32 pipes)'
    zero = [0.0, 0.0, 0, 0.0]
    conv_ms = sum(by.get(e, zero)[0] for e in entries)
    conv_flops = sum(by.get(e, zero)[1] for e in entries)
    conv_n = sum(by.get(e, zero)[2] for e in entries)
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms else 0.0
    traffic, traffic_src, traffic_all = None, None, {}
    prof_dir = os.path.join(ROOT, 'profiles')
    cands = sorted(f for f in os.listdir(prof_dir) if f.endswith('_traffic.json')) if os.path.isdir(prof_dir) else []
    if cands and dtype == torch.bfloat16:
        traffic_all = json.load(open(os.path.join(prof_dir, cands[-1])))
        traffic, traffic_src = traffic_all.get('dram_bytes_per_launch'), 'profiles/' + cands[-1]
    roofline = {'kernel': kname, 'bound': 'tensor', 'achieved': achieved, 'peak': pk['tensor'],
                'unit': 'TFLOP/s', 'frac': achieved / pk['tensor'], 'traffic': traffic, 'traffic_unit': 'bytes/launch',
                'traffic_source': traffic_src, 'peak_source': pk['src'] + ' (bf16_tflops_sustained: kernel timed inside a long step)',
                'launches_per_step': conv_n, 'ms_per_step': conv_ms, 'share_of_step': conv_ms / step_ms,
                'algorithmic_gflop_per_step': conv_flops / 1e9,
                'per_entry_ms': {e: round(by[e][0], 3) for e in entries if e in by},
                'conv_layers': conv_layer_table(recs, pk),
                'other_kernels_ms_per_step': {k: round(v[0], 3) for k, v in sorted(by.items()) if k not in entries}}

    # HBM-bound kernels of stages 1-2 (+ the position-table gather-add): achieved = algorithmic bytes / CUDA-event time
    hbm_names = ['os3d_voxelize', 'os3d_scatter_max_f32', 'os3d_scatter_max_bf16', 'os3d_scatter_max_sorted_bf16', 'os3d_scatter_mean_small_bf16', 'os3d_scatter_mean_f32', 'os3d_gather_rows', 'os3d_subm_table',
                 'os3d_strided_tables', 'os3d_kernel_map_tiles', 'os3d_kernel_map_order', 'os3d_window_partition',
                 'os3d_add_table_rows', 'os3d_gelu_bf16']
    roofline['hbm_kernels'] = {
        k: {'ms': round(by[k][0], 3), 'algorithmic_mb': round(by[k][1] / 1e6, 1),
            'achieved_gbs': round(by[k][1] / (by[k][0] * 1e-3) / 1e9, 1) if by[k][0] else 0.0,
            'frac_of_hbm_peak': round(by[k][1] / (by[k][0] * 1e-3) / 1e9 / pk['hbm'], 3) if by[k][0] else 0.0}
        for k in hbm_names if k in by and by[k][1] > 0}
    roofline['hbm_peak_gbs'] = pk['hbm']
    # window attention: useful FLOPs = 4 * sum_windows n^2 * C (QK^T + PV over real tokens only; the reference pads to max_tokens)
    for entry in ('os3d_window_attention_bf16_tc', 'os3d_window_attention_bf16_tc_prenorm', 'os3d_window_attention_bf16_v2'):
        at = by.get(entry)
        if at and at[0] and at[1]:
            roofline['window_attention_tc_kernel'] = {
                'entry': entry, 'ms': round(at[0], 3), 'launches_per_step': at[2], 'useful_gflop': round(at[1] / 1e9, 1),
                'achieved_tflops': round(at[1] / (at[0] * 1e-3) / 1e12, 1),
                'frac_of_tensor_peak': round(at[1] / (at[0] * 1e-3) / 1e12 / pk['tensor'], 4),
                'algorithmic_mb': round(at[3] / 1e6, 1),
                'frac_of_hbm_peak': round(at[3] / (at[0] * 1e-3) / 1e9 / pk['hbm'], 3)}
            tk = traffic_all.get('window_attention_tc_kernel')
            if isinstance(tk, dict) and tk.get('dram_bytes_per_step'):
                roofline['window_attention_tc_kernel']['dram_mb_per_step_ncu'] = round(tk['dram_bytes_per_step'] / 1e6, 1)
    # the persistent dense kernels of the SWFormer layers and the point MLPs: HBM is their roofline
    # (bytes = inputs + outputs (+ residual) (+ weights once))
    for entry, key in (('os3d_linear_tc_bf16', 'linear_tc_kernel'), ('os3d_swformer_mlp_bf16', 'swformer_mlp_tc_kernel'),
                       ('os3d_mlp_chain_bf16', 'mlp_chain_tc_kernel'), ('os3d_qkv_proj_bf16', 'qkv_proj_tc_kernel')):
        lt = by.get(entry)
        if lt and lt[0]:
            roofline[key] = {
                'ms': round(lt[0], 3), 'launches_per_step': lt[2], 'algorithmic_mb': round(lt[3] / 1e6, 1),
                'achieved_gbs': round(lt[3] / (lt[0] * 1e-3) / 1e9, 1),
                'frac_of_hbm_peak': round(lt[3] / (lt[0] * 1e-3) / 1e9 / pk['hbm'], 3),
                'achieved_tflops': round(lt[1] / (lt[0] * 1e-3) / 1e12, 1)}
            tk = traffic_all.get(key)                 # measured DRAM bytes (ncu launch list) next to the algorithmic ones
            if isinstance(tk, dict) and tk.get('dram_bytes_per_step'):
                roofline[key]['dram_mb_per_step_ncu'] = round(tk['dram_bytes_per_step'] / 1e6, 1)
    return roofline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--frames', type=int, default=FRAMES_PER_GPU)
    ap.add_argument('--config', default=CONFIG, choices=['waymo_one_sweep', 'waymo_one_sweep_cylinder', 'waymo_multi_sweeps'],
                    help='default: the headline workload (BASELINE configs[1]); the other two are BASELINE configs[2], [3]')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true', help='skip the fp32 / cylinder / multi-sweep sub-records')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    if args.mode == 'train':
        from tools import train_step
        return train_step.bench_main(args)

    import torch
    import torch.distributed as dist
    from openseg3d_b200 import _lib
    from openseg3d_b200.utils.sharding import job_totals

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    _lib.lib()                                    # fail loudly if the CUDA library is missing

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return job_totals(0, e0.elapsed_time(e1), 'cuda')[1]          # MAX over ranks

    arm = Arm(args.config, args.dtype, args.frames, rank)
    arm.warm(args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launches()
    total_ms = timed(arm.step_resident, args.steps)
    launches = _lib.launches() - l0
    sampler.stop_flag = True
    arm.step_e2e()
    e2e_ms = timed(arm.step_e2e, args.steps)

    total_points = job_totals(arm.n_points, 0.0, 'cuda')[0]               # SUM over ranks
    value = total_points * args.steps / (total_ms * 1e-3)
    e2e_value = total_points * args.steps / (e2e_ms * 1e-3)

    pk = peaks()
    by, recs = kernel_profile(arm)
    roofline = build_roofline(arm, by, recs, total_ms / args.steps, pk)

    line = {'metric': 'points/sec, Waymo 1-sweep seg forward', 'value': value, 'unit': 'points/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': workload_config(args.config, args.frames, arm.n_points, world),
            'e2e': {'value': e2e_value, 'unit': 'points/s', 'h2d_bytes_per_step': arm.host.numel() * 4 * world,
                    'd2h_bytes_per_step': arm.n_labelled * world, 'ms_per_step': e2e_ms / args.steps},
            'gpu_launches': launches, 'roofline': roofline}

    # ---- the other measured modes of SURVEY.md §8(d): fp32 (the reference's arithmetic, the exact-parity mode) on the
    # headline workload, and the cylinder / 3-sweep workloads (BASELINE configs[2], [3]) -- each W warm-up + K timed steps,
    # same timing rules, max over ranks; sub-records of the one line the driver parses
    if not args.no_other_configs:
        del arm
        torch.cuda.empty_cache()
        others = {}
        sub = [(args.config, 'f32' if args.dtype == 'bf16' else 'bf16')]
        sub += [(c, args.dtype) for c in ('waymo_one_sweep_cylinder', 'waymo_multi_sweeps') if c != args.config]
        for cfg_name, dt in sub:
            a = Arm(cfg_name, dt, args.frames, rank)
            a.warm(args.warmup)
            k = max(3, args.steps // 2) if dt == 'f32' else args.steps
            ms = timed(a.step_resident, k)
            pts = job_totals(a.n_points, 0.0, 'cuda')[0]
            b2, r2 = kernel_profile(a)
            rl = build_roofline(a, b2, r2, ms / k, pk)
            others[f'{cfg_name}/{dt}'] = {
                'value': pts * k / (ms * 1e-3), 'unit': 'points/s', 'ms_per_step': ms / k, 'steps': k, 'dtype': dt,
                'points_per_gpu': a.n_points,
                'conv': {'kernel': rl['kernel'], 'ms_per_step': round(rl['ms_per_step'], 3),
                         'achieved_tflops': round(rl['achieved'], 1), 'frac_of_tensor_peak': round(rl['frac'], 3),
                         'peak': 'bf16 tensor sustained (no separate measured FP32-pipe peak)' if dt == 'f32' else 'bf16 tensor sustained'},
                'kernels_ms_per_step': {k2: round(v[0], 3) for k2, v in sorted(b2.items()) if v[0] >= 0.05}}
            del a
            torch.cuda.empty_cache()
        line['other_configs'] = others

    if rank == 0:
        sampler.join(timeout=2)
        line['clocks'] = sampler.summary()
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            pps, n, dt = cpu_forward_sample(cores, args.config)
            line['cpu_baseline'] = {'value': pps, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
                                    'sample': cpu_sample_text(n, dt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
