"""ORACLE tooling (container-only): freeze golden vectors into tests/golden/.

Run in the authoring container, where ``/root/reference`` exists:

    python -m oracle.make_golden

What it writes (all small, all committed):

* ``octree_fixtures.npz``  -- the reference's own known-answer vectors for the
  integer octree path, re-packed from
  ``/root/reference/libs/dwconv/test/data/octree/test_00{1..5}.npz`` and
  ``.../batch_45.npz`` (inputs: points; answers: key, child, nnum, nnum_nempty,
  merged 27-neighbour table).  Data only, no reference source.
* ``state_shapes_<cfg>.json`` -- parameter names/shapes of the reference model
  built by the reference's own ``model_factory`` (the state_dict layout the
  drop-in must keep loadable).
* ``descriptors.npz`` -- descriptors computed by the reference's own,
  unmodified ``models/*.py`` (run on CPU over ``oracle/ocnn_standin.py``) for
  seeded synthetic submaps and name-seeded weights
  (``oracle.model_ref.synthetic_state_dict``); and, next to each, the value of
  ``oracle.model_ref.forward`` at generation time (must agree to 1e-5).
* ``octree_t.npz`` -- window / relay-token bookkeeping produced by the
  reference's ``models/octree.py:OctreeT.build_t`` for a CS-Wild-Places batch
  in which some submaps own zero relay tokens.
* ``cylindrical.npz`` -- reference ``CylindricalCoordinates`` outputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import model_ref as M
from . import ocnn_standin as S
from . import octree_ref as R

REF = S.REFERENCE_ROOT
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# name -> (model cfg, octree depth, [(n_points, aerial)], cloud seed, weight mode)
CASES = {
    'oxford_b1_init':   ('oxford', 9, [(4096, False)], 1, 'init'),
    'oxford_b1_stress': ('oxford', 9, [(4096, False)], 1, 'stress'),
    'oxford_b4_init':   ('oxford', 9, [(4096, False)] * 4, 2, 'init'),
    'oxford_b4_stress': ('oxford', 9, [(4096, False)] * 4, 2, 'stress'),
    'cswp_b4_init':     ('cs-wild-places', 7, [(30000, False), (60000, True)] * 2, 3, 'init'),
    'cswp_b6_stress':   ('cs-wild-places', 7, [(30000, False)] * 6, 4, 'stress'),
    'wp_b3_init':       ('wild-places', 7, [(30000, False)] * 3, 5, 'init'),
    'wp_b3_stress':     ('wild-places', 7, [(8000, False), (30000, False), (500, False)], 6, 'stress'),
}


def case_clouds(name):
    cfg, depth, spec, seed, mode = CASES[name]
    g = torch.Generator().manual_seed(seed)
    clouds = [M.lidar_cloud(n, g, aerial=a) for n, a in spec]
    if cfg == 'wild-places':
        clouds = [cylindrical_ref(c) for c in clouds]
    return clouds


def cylindrical_ref(cloud: np.ndarray) -> np.ndarray:
    """eval/pnv_evaluate.py:166-171 using the reference's own converter."""
    S.install()
    from datasets.coordinate_utils import CylindricalCoordinates
    data = torch.from_numpy(cloud)
    data = data[torch.linalg.norm(data[:, :2], dim=1) <= 1.0]
    return CylindricalCoordinates(use_octree=True)(data.clone()).numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    S.install()
    # ---- 1. integer octree fixtures -----------------------------------------
    base = f'{REF}/libs/dwconv/test/data/'
    pack = {}
    for i in range(1, 6):
        d = np.load(base + 'octree/test_%03d.npz' % i)
        for k in ('points', 'key', 'child', 'nnum', 'nnum_nempty'):
            pack[f't{i}_{k}'] = d[k]
        pack[f't{i}_depth'] = d['depth']
        pack[f't{i}_full_depth'] = d['full_depth']
    b = np.load(base + 'batch_45.npz')
    for k in ('key', 'child', 'nnum', 'nnum_nempty', 'neigh', 'depth', 'full_depth'):
        pack[f'b45_{k}'] = b[k]
    np.savez_compressed(os.path.join(OUT, 'octree_fixtures.npz'), **pack)

    # ---- 2. state_dict layouts ------------------------------------------------
    models = {}
    for cfg in ('oxford', 'cs-wild-places', 'wild-places', 'cs-campus3d'):
        m = S.reference_model(f'{REF}/models/hotformerloc_{cfg}_cfg.txt')
        models[cfg] = m
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        with open(os.path.join(OUT, f'state_shapes_{cfg}.json'), 'w') as f:
            json.dump(shapes, f, indent=0)

    # ---- 3. descriptors from the reference's own model code ---------------------
    desc = {}
    for name, (cfg, depth, spec, seed, mode) in CASES.items():
        m = models[cfg]
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = M.synthetic_state_dict(shapes, mode=mode)
        m.load_state_dict(sd)
        clouds = case_clouds(name)
        with torch.inference_mode():
            y = m(S.make_batch(clouds, depth))['global'].numpy()
        hp = M.HParams.from_cfg(f'{REF}/models/hotformerloc_{cfg}_cfg.txt')
        o = R.build_batch(clouds, depth)
        g = M.forward(sd, o, hp).numpy()
        err = np.abs(y - g).max()
        print(f'{name}: reference-vs-oracle max-abs {err:.2e}', flush=True)
        assert err < 1e-5, name
        desc[name + '_reference'] = y
        desc[name + '_oracle'] = g
        desc[name + '_nnum_nempty'] = o.nnum_nempty
    np.savez_compressed(os.path.join(OUT, 'descriptors.npz'), **desc)

    # ---- 4. OctreeT bookkeeping (zero-RT submaps) -------------------------------
    from models.octree import OctreeT
    tpack = {}
    for name in ('cswp_b6_stress', 'oxford_b4_init'):
        cfg, depth, spec, seed, mode = CASES[name]
        hp = M.HParams.from_cfg(f'{REF}/models/hotformerloc_{cfg}_cfg.txt')
        octree = S.make_batch(case_clouds(name), depth)['octree']
        d0 = depth - hp.stem_down
        t = OctreeT(octree, hp.patch_size, hp.dilation, True, max_depth=d0,
                    start_depth=d0 - 3, rt_size=1, ADaPE_mode=hp.ADaPE_mode,
                    num_pyramid_levels=3, num_octf_levels=1)
        t.build_t()
        for d in range(d0 - 3, d0 + 1):
            tpack[f'{name}_nnum_a_{d}'] = np.array(int(t.nnum_a[d]))
            tpack[f'{name}_batch_idx_{d}'] = t.batch_idx[d].numpy().astype(np.int16)
            if d < d0:
                tpack[f'{name}_num_windows_{d}'] = t.batch_num_windows[d].numpy()
                tpack[f'{name}_rt_batch_idx_{d}'] = t.rt_batch_idx[d].numpy().astype(np.int16)
                tpack[f'{name}_rt_init_mask_{d}'] = np.packbits(t.rt_init_mask[d].numpy())
                tpack[f'{name}_window_stats_{d}'] = t.window_stats[d].numpy()
        tpack[f'{name}_rt_combined'] = t.batch_num_relay_tokens_combined.numpy()
        tpack[f'{name}_rt_attn_allowed'] = np.packbits((t.rt_attn_mask == 0).numpy())
        tpack[f'{name}_rt_attn_shape'] = np.array(t.rt_attn_mask.shape)
    np.savez_compressed(os.path.join(OUT, 'octree_t.npz'), **tpack)

    # ---- 5. cylindrical conversion ----------------------------------------------
    g = torch.Generator().manual_seed(7)
    c = M.lidar_cloud(5000, g)
    np.savez_compressed(os.path.join(OUT, 'cylindrical.npz'), cloud=c, out=cylindrical_ref(c))
    print('golden vectors written to', OUT)



def recall_case(seed=11, n_runs=3, n_per_run=300, dim=256):
    """Synthetic evaluation sets in the reference's pickle format
    (datasets/WildPlaces/generate_test_sets.py:46-78): list over runs of
    {idx: {'query': path, 'northing', 'easting', <db_run>: [true neighbour ids]}}."""
    rng = np.random.default_rng(seed)
    sets, vecs = [], []
    base = rng.normal(size=(n_per_run, dim)).astype(np.float32)
    for r in range(n_runs):
        pos = np.stack([np.arange(n_per_run) * 10.0, np.zeros(n_per_run)], 1)
        v = base + 2.0 * rng.normal(size=base.shape).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        vecs.append(v.astype(np.float32))
        sets.append({i: {'query': f'run{r}/{i}.bin', 'northing': float(pos[i, 0]),
                         'easting': float(pos[i, 1])} for i in range(n_per_run)})
    for r in range(n_runs):
        for i in range(n_per_run):
            for m in range(n_runs):
                near = [j for j in range(n_per_run) if abs(j - i) <= 1] if (i % 7) else []
                sets[r][i][m] = near
    return sets, vecs


def make_recall_golden():
    """Reference get_recall (eval/pnv_evaluate.py:228-315, sklearn KDTree path) on
    synthetic sets -> tests/golden/recall.npz."""
    S.install()
    import importlib
    ev = importlib.import_module('eval.pnv_evaluate')
    sets, vecs = recall_case()
    out = {}
    for m in range(3):
        for n in range(3):
            if m == n:
                continue
            rec, opr, mrr = ev.get_recall(m, n, vecs, vecs, sets, sets)
            out[f'recall_{m}_{n}'] = np.asarray(rec)
            out[f'opr_{m}_{n}'] = np.asarray(opr)
            out[f'mrr_{m}_{n}'] = np.asarray(mrr)
    np.savez_compressed(os.path.join(OUT, 'recall.npz'), **out)
    print('recall golden written')


def make_normalize_golden():
    """Reference ``Normalize`` (datasets/augmentation.py:185-235) on a metric-scale cloud, every mode the
    eval path can select (config keys normalize_points / scale_factor / unit_sphere_norm) -> normalize.npz."""
    S.install()
    from datasets.augmentation import Normalize
    from hotformerloc_b200.datasets.synthetic import trajectory_clouds
    cloud = trajectory_clouds(1, 1, 6000, seed=3)[0][0].astype(np.float32)
    out = {'cloud': cloud}
    variants = {'bbox': {}, 'scale30': dict(scale_factor=30.0), 'sphere': dict(unit_sphere_norm=True),
                'sphere_scale40': dict(unit_sphere_norm=True, scale_factor=40.0),
                'range2': dict(norm_range=2.0), 'bbox_nocenter': dict(zero_mean=False)}
    for name, kw in variants.items():
        out[name] = Normalize(**kw)(torch.from_numpy(cloud).clone()).numpy()
    np.savez_compressed(os.path.join(OUT, 'normalize.npz'), **out)
    print('normalize golden written')


def make_gem_golden():
    """pooling=PyramidOctGeM through the reference's own model code (models/layers/pooling.py:58-103):
    state_dict layout + descriptors of a 3-submap Oxford batch -> state_shapes_oxford_gem.json,
    descriptors_gem.npz."""
    import tempfile
    S.install()
    cfg = open(f'{REF}/models/hotformerloc_oxford_cfg.txt').read()
    lines = [('pooling = PyramidOctGeM' if l.split('=')[0].strip() == 'pooling' else l) for l in cfg.splitlines()]
    path = os.path.join(tempfile.mkdtemp(prefix='hfl_gem_'), 'hotformerloc_oxford_gem_cfg.txt')
    open(path, 'w').write('\n'.join(lines) + '\n')
    m = S.reference_model(path)
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, 'state_shapes_oxford_gem.json'), 'w') as f:
        json.dump(shapes, f, indent=0)
    m.load_state_dict(M.synthetic_state_dict({k: tuple(v) for k, v in shapes.items()}, mode='stress'))
    g = torch.Generator().manual_seed(9)
    clouds = [M.lidar_cloud(4096, g) for _ in range(3)]
    with torch.inference_mode():
        y = m(S.make_batch(clouds, 9))['global'].numpy()
    np.savez_compressed(os.path.join(OUT, 'descriptors_gem.npz'), reference=y)
    print('GeM golden written')


if __name__ == '__main__':
    which = sys.argv[1:] or ['main', 'recall', 'normalize', 'gem']
    if 'main' in which:
        main()
    if 'recall' in which:
        make_recall_golden()
    if 'normalize' in which:
        make_normalize_golden()
    if 'gem' in which:
        make_gem_golden()
