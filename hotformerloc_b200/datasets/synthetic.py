"""Synthetic inputs of the named workloads (BASELINE.json configs):
seeded, CPU-deterministic generators shared by bench.py, the tools, the tests and the oracle's golden
scripts.  No dataset or checkpoint can be fetched here, so every measured workload is synthetic."""
from __future__ import annotations

import os
import pickle
from typing import Dict, List

import numpy as np
import torch


def lidar_cloud(n: int, g: torch.Generator, aerial: bool = False) -> np.ndarray:
    """'lidar-ish' submap in [-1, 1]^3: uniform ground footprint, a dense ground sheet
    (60 %, aerial: 30 %) and structure above it (canopy-heavy when ``aerial``)."""
    xy = (torch.rand(n, 2, generator=g) * 2 - 1) * 0.95
    m = torch.rand(n, generator=g) < (0.3 if aerial else 0.6)
    if aerial:
        z = torch.where(m, 0.02 * torch.randn(n, generator=g) - 0.3,
                        torch.rand(n, generator=g) * 0.9 - 0.2)
    else:
        z = torch.where(m, 0.02 * torch.randn(n, generator=g) - 0.3,
                        torch.rand(n, generator=g) * 0.8 - 0.3)
    return torch.cat([xy, z[:, None]], 1).clamp(-1, 1).numpy()


def write_pcd(path: str, xyz: np.ndarray) -> None:
    """binary PCD v0.7 with x y z float fields (the CS-Wild-Places / Wild-Places on-disk format,
    datasets/CSWildPlaces/CSWildPlaces_raw.py:14-23 reads it through open3d)."""
    hdr = ('# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n'
           f'COUNT 1 1 1\nWIDTH {len(xyz)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(xyz)}\nDATA binary\n')
    with open(path, 'wb') as f:
        f.write(hdr.encode())
        f.write(np.ascontiguousarray(xyz, dtype=np.float32).tobytes())


def iter_trajectory_clouds(runs: int, per_run: int, points: int, seed: int = 11, jitter: float = 0.01,
                           yaw: float = 0.002, drop: float = 0.02):
    """``runs`` traversals of one trajectory of ``per_run`` places, in metres (~60 m submaps).  A place is
    a tilted ground plane + 12 box-shaped structures; a traversal re-observes it with point jitter (metres),
    a small yaw (radians) and a fraction of the points dropped -- near-duplicate positives, so recall on
    random-init descriptors is a meaningful check.  Defaults (1 cm, 2 mrad, 2 %) were chosen with the CPU
    oracle: a random-init network is not invariant to larger re-observation noise (5 cm / 30 mrad / 10 %
    gives recall@1 of 5-9 %, i.e. rank order dominated by noise; these give ~100 % with a top-1 margin
    two orders of magnitude above the bf16 descriptor error).  Yields (run, place, float64 (n, 3)) in the order
    the generator draws them (places first, then run by run), one cloud at a time."""
    rng = np.random.default_rng(seed)

    def place():
        k = 12
        ctr, half = rng.uniform(-0.8, 0.8, (k, 2)), rng.uniform(0.03, 0.15, (k, 1))
        top, tilt = rng.uniform(0.1, 0.6, k), rng.normal(0, 0.1, 2)
        n_obj = points // 2
        xy_g = rng.uniform(-0.95, 0.95, (points - n_obj, 2))
        z_g = -0.3 + xy_g @ tilt + rng.normal(0, 0.01, len(xy_g))
        which = rng.integers(0, k, n_obj)
        xy_o = ctr[which] + rng.uniform(-1, 1, (n_obj, 2)) * half[which]
        z_o = -0.3 + xy_o @ tilt + rng.uniform(0, 1, n_obj) * top[which]
        pts = np.concatenate([np.concatenate([xy_g, xy_o]), np.concatenate([z_g, z_o])[:, None]], 1)
        return np.clip(pts, -1, 1) * 30.0
    places = [place() for _ in range(per_run)]
    for r in range(runs):
        for i, base in enumerate(places):
            ang = rng.normal(0, yaw)
            c, sn = np.cos(ang), np.sin(ang)
            keep = rng.random(len(base)) > drop
            pts = base[keep] + rng.normal(0, jitter, (int(keep.sum()), 3))
            yield r, i, pts @ np.array([[c, -sn, 0], [sn, c, 0], [0, 0, 1]]).T


def trajectory_clouds(runs: int, per_run: int, points: int, seed: int = 11, **kw) -> List[List[np.ndarray]]:
    """All clouds of iter_trajectory_clouds() at once: clouds[run][place] float64 (n, 3)."""
    out = [[None] * per_run for _ in range(runs)]
    for r, i, pts in iter_trajectory_clouds(runs, per_run, points, seed, **kw):
        out[r][i] = pts
    return out


def eval_sets(runs: int, per_run: int, subdir: str = 'Venman', ext: str = 'pcd') -> List[Dict]:
    """Evaluation dicts in the reference's pickle format (datasets/WildPlaces/generate_test_sets.py:46-78):
    list over runs of {idx: {'query': relpath, 'northing', 'easting', <db run>: [true neighbour ids]}}."""
    sets = []
    for r in range(runs):
        s = {}
        for i in range(per_run):
            rel = os.path.join(subdir, f'run{r}', 'Clouds', f'{i:06d}.{ext}')
            s[i] = {'query': rel, 'northing': float(3.0 * i), 'easting': float(0.5 * r)}
            for m in range(runs):                       # the true neighbour: the same place in the other runs
                s[i][m] = [i] if m != r else []
        sets.append(s)
    return sets


def make_eval_dataset(root: str, runs: int, per_run: int, points: int, seed: int = 11,
                      subdir: str = 'Venman') -> List[Dict]:
    """Write trajectory_clouds() as binary .pcd files under ``root`` and return the evaluation dicts."""
    sets = eval_sets(runs, per_run, subdir)
    for r in range(runs):
        os.makedirs(os.path.join(root, subdir, f'run{r}', 'Clouds'), exist_ok=True)
    for r, i, pts in iter_trajectory_clouds(runs, per_run, points, seed):      # written as drawn: one cloud in memory
        write_pcd(os.path.join(root, sets[r][i]['query']), pts)
    return sets


def write_eval_pickles(root: str, sets: List[Dict], names) -> None:
    for name in names:
        pickle.dump(sets, open(os.path.join(root, name), 'wb'))
