"""ORACLE tooling (container-only): recall golden for BASELINE.json configs[3] on a reduced set.

    python -m oracle.make_golden_config4 [--runs 4 --per-run 512 --points 30000 --threads 6]

Wild-Places cfg (cylindrical, K = 48, no ADaPE, val_batch_size 128, skip_same_run), synthetic dataset of
`runs` traversals x `per-run` places in the reference's on-disk format
(hotformerloc_b200.datasets.synthetic.make_eval_dataset).  Every stage except the model forward is the
REFERENCE'S OWN code imported from /root/reference over the stand-ins (oracle/ocnn_standin.py):
  file loader (datasets/CSWildPlaces/CSWildPlaces_raw.py:14-23, open3d replaced by a 10-line PCD parser)
  -> Normalize (datasets/augmentation.py:185-235) -> range masks -> CylindricalCoordinates
  (datasets/coordinate_utils.py:68-116) -> Octree.build_octree / merge_octrees in chunks of val_batch_size
  (eval/pnv_evaluate.py:129-187) -> get_recall (eval/pnv_evaluate.py:228-315).
The forward is oracle.model_ref.forward (pinned to the reference's models/*.py at 2-5e-7 by
oracle/make_golden.py; the reference modules themselves materialise the (N_win,K,K,3,H) RPE gather and
need tens of GB per 128-submap batch on CPU).  Weights: name-seeded synthetic_state_dict(mode='init').

Writes tests/golden/config4_recall.npz: recall@N / recall@1% / MRR of every (database, query) pair and
their averages, plus the oracle descriptors of the first 32 places of every run for a cosine check.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

from . import model_ref as M
from . import ocnn_standin as S
from . import octree_ref as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
KEEP = 32          # descriptors kept per run


def _read_pcd_xyz(path):
    with open(path, 'rb') as f:
        n = 0
        while True:
            line = f.readline().decode()
            if line.startswith('POINTS'):
                n = int(line.split()[1])
            if line.startswith('DATA'):
                assert line.split()[1] == 'binary'
                break
        return np.frombuffer(f.read(12 * n), dtype=np.float32).reshape(n, 3).astype(np.float64)


def install_open3d_reader():
    o3d = sys.modules.get('open3d') or types.ModuleType('open3d')
    io = types.ModuleType('open3d.io')
    io.read_point_cloud = lambda p: types.SimpleNamespace(points=_read_pcd_xyz(p))
    o3d.io = io
    sys.modules['open3d'], sys.modules['open3d.io'] = o3d, io


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--runs', type=int, default=4)
    ap.add_argument('--per-run', type=int, default=512)
    ap.add_argument('--points', type=int, default=30000)
    ap.add_argument('--threads', type=int, default=6)
    ap.add_argument('--root', default='/tmp/hfl_config4_golden')
    ap.add_argument('--name', default='config4_recall')
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    S.install()
    install_open3d_reader()
    from hotformerloc_b200.datasets.synthetic import make_eval_dataset
    sets = make_eval_dataset(args.root, args.runs, args.per_run, args.points, seed=11)
    # the reference's own config files, dataset_folder redirected
    cfg = open(f'{S.REFERENCE_ROOT}/config/config_wild-places.txt').read()
    cfg = '\n'.join(('dataset_folder = ' + args.root) if l.startswith('dataset_folder') else l
                    for l in cfg.splitlines())
    cfg_path = os.path.join(args.root, 'config_wild-places.txt')
    open(cfg_path, 'w').write(cfg)
    model_cfg = f'{S.REFERENCE_ROOT}/models/hotformerloc_wild-places_cfg.txt'
    from misc.utils import TrainingParams
    import importlib
    ev = importlib.import_module('eval.pnv_evaluate')
    params = TrainingParams(cfg_path, model_cfg)
    assert params.val_batch_size == 128 and params.model_params.coordinates == 'cylindrical'
    shapes = json.load(open(os.path.join(OUT, 'state_shapes_wild-places.json')))
    sd = M.synthetic_state_dict(shapes, mode='init')
    hp = M.HParams.from_cfg(model_cfg)

    # eval/pnv_evaluate.py:129-187 with the model call replaced by the oracle forward
    loader = ev.CSWildPlacesPointCloudLoader()
    norm = ev.Normalize(scale_factor=params.scale_factor, unit_sphere_norm=params.unit_sphere_norm)
    conv = ev.CylindricalCoordinates(use_octree=True)

    def latent(data_set):
        out, cur = [], []
        keys = list(data_set)
        for i, k in enumerate(keys):
            data = torch.tensor(loader(os.path.join(params.dataset_folder, data_set[k]['query'])))
            data = norm(data)
            data = data[torch.all(abs(data) <= 1.0, dim=1)]
            data = data[torch.all(torch.linalg.norm(data[:, :2], dim=1)[:, None] <= 1.0, dim=1)]
            cur.append(conv(data).numpy())
            if len(cur) >= params.val_batch_size or i == len(keys) - 1:
                out.append(M.forward(sd, R.build_batch(cur, params.octree_depth), hp).numpy())
                cur = []
                print(f'  {i + 1}/{len(keys)} submaps, {time.time() - t0:.0f} s', flush=True)
        return np.concatenate(out)
    t0 = time.time()
    emb = [latent(s) for s in sets]
    pack = {'runs': np.array(args.runs), 'per_run': np.array(args.per_run), 'points': np.array(args.points)}
    recs, oprs, mrrs = [], [], []
    for m in range(args.runs):
        for n in range(args.runs):
            if m == n and params.skip_same_run:
                continue
            rec, opr, mrr = ev.get_recall(m, n, emb, emb, sets, sets)
            pack[f'recall_{m}_{n}'], pack[f'opr_{m}_{n}'], pack[f'mrr_{m}_{n}'] = map(np.asarray, (rec, opr, mrr))
            recs.append(rec), oprs.append(opr), mrrs.append(mrr)
    pack['ave_recall'] = np.mean(recs, axis=0)
    pack['ave_one_percent_recall'] = np.mean(oprs)
    pack['ave_mrr'] = np.mean(mrrs)
    for r in range(args.runs):
        pack[f'desc_run{r}'] = emb[r][:KEEP].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, args.name + '.npz'), **pack)
    print(f'recall@1 {pack["ave_recall"][0]:.3f}  recall@1% {pack["ave_one_percent_recall"]:.3f}  '
          f'MRR {pack["ave_mrr"]:.3f}  ({time.time() - t0:.0f} s)')


if __name__ == '__main__':
    main()
