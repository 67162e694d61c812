"""Seeded synthetic Waymo-shape LiDAR frames (SURVEY.md §8(d) "Synthetic frame").

There is no dataset in the sandbox, so tests, the oracle and bench.py all draw frames from
this generator.  One frame = one 64-beam x 2650-column sweep over a ground plane with
piecewise-constant walls, 7 % jittered second returns; columns follow what
``WaymoDataset.load_points`` produces (reference seg3d/datasets/waymo_dataset.py:143-153):
``x, y, z, range(=0) | dt, tanh(intensity), elongation`` as float32.
"""
import numpy as np

N_BEAMS = 64
N_COLS = 2650
SENSOR_Z = 2.1


def make_sweep(seed, n_beams=N_BEAMS, n_cols=N_COLS):
    """One sweep -> float32 [N, 6]."""
    rng = np.random.default_rng(seed)
    incl = np.deg2rad(np.linspace(-17.6, 2.4, n_beams))
    azim = np.linspace(-np.pi, np.pi, n_cols, endpoint=False)
    n_blocks = (n_cols + 49) // 50
    wall = np.repeat(rng.uniform(8.0, 70.0, n_blocks), 50)[:n_cols]
    inc, az = np.meshgrid(incl, azim, indexing='ij')
    with np.errstate(divide='ignore'):
        ground = np.where(inc < 0, SENSOR_Z / np.sin(-inc), np.inf)
    r = np.minimum(ground, wall[None, :]) + rng.normal(0.0, 0.02, inc.shape)
    r = r.reshape(-1)
    inc = inc.reshape(-1)
    az = az.reshape(-1)
    keep = (r > 1.0) & (r < 75.0)
    r, inc, az = r[keep], inc[keep], az[keep]
    xyz = np.stack([r * np.cos(inc) * np.cos(az), r * np.cos(inc) * np.sin(az), r * np.sin(inc) + SENSOR_Z], axis=1)
    # 7 % jittered second returns
    n2 = int(round(0.07 * xyz.shape[0]))
    pick = rng.choice(xyz.shape[0], n2, replace=False)
    xyz2 = xyz[pick] + rng.normal(0.0, 0.3, (n2, 3))
    xyz = np.concatenate([xyz, xyz2], axis=0)
    n = xyz.shape[0]
    intensity = np.tanh(rng.exponential(0.3, n))
    elong = rng.uniform(0.0, 1.5, n)
    pts = np.zeros((n, 6), dtype=np.float32)
    pts[:, :3] = xyz.astype(np.float32)
    pts[:, 4] = intensity.astype(np.float32)
    pts[:, 5] = elong.astype(np.float32)
    return pts


def make_frame(seed, num_sweeps=1, n_beams=N_BEAMS, n_cols=N_COLS):
    """A frame of ``num_sweeps`` sweeps.  Sweep k>0 is seed+k, ego-shifted by (0.8, 0.05, 0)*k and
    carries dt = 0.1*k in column 3 (current sweep has dt == 0, reference segformer.py:97-99).
    Returns (points [N, 6] f32, cur_point_count)."""
    cur = make_sweep(seed, n_beams, n_cols)
    if num_sweeps == 1:
        return cur, cur.shape[0]
    sweeps = [cur]
    for k in range(1, num_sweeps):
        s = make_sweep(seed + k, n_beams, n_cols)
        s[:, 0] -= np.float32(0.8 * k)
        s[:, 1] -= np.float32(0.05 * k)
        s[:, 3] = np.float32(0.1 * k)
        sweeps.append(s)
    return np.concatenate(sweeps, axis=0), cur.shape[0]


def cart2polar_rows(points):
    """Cylinder-mode rows ``[rho, phi, z, x, y, rest]`` in numpy float32, exactly as the reference
    computes them on the host (seg3d/utils/pointops_utils.py:8-11, waymo_dataset.py:270-273)."""
    rho = np.sqrt(points[:, 0] ** 2 + points[:, 1] ** 2)
    phi = np.arctan2(points[:, 1], points[:, 0])
    polar = np.stack((rho, phi, points[:, 2]), axis=1)
    return np.concatenate((polar, points[:, :2], points[:, 3:]), axis=1).astype(np.float32)


def make_batch(seeds, num_sweeps=1, cylinder=False, n_beams=N_BEAMS, n_cols=N_COLS):
    """Collated raw batch: ``points [sum N, 1+D]`` with the batch index in column 0 (what
    collate_batch builds, waymo_dataset.py:347-352) and ``point_id_offset`` (cumulative current-sweep
    counts, :367-372)."""
    rows, offs, count = [], [], 0
    for b, seed in enumerate(seeds):
        pts, ncur = make_frame(seed, num_sweeps, n_beams, n_cols)
        if cylinder:
            pts = cart2polar_rows(pts)
        rows.append(np.concatenate([np.full((pts.shape[0], 1), b, np.float32), pts], axis=1))
        count += ncur
        offs.append(count)
    return np.concatenate(rows, axis=0), np.asarray(offs, dtype=np.int64)
