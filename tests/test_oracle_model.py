"""Pins oracle/model_ref.py against descriptors produced by the reference's
own, unmodified models/*.py (tests/golden/descriptors.npz, made by
oracle/make_golden.py in the authoring container).  fp32, max-abs 1e-5."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_ref as M
from oracle import octree_ref as R
from oracle.make_golden import CASES

from hotformerloc_b200.config.presets import write_configs


def _clouds(name):
    cfg, depth, spec, seed, mode = CASES[name]
    g = torch.Generator().manual_seed(seed)
    clouds = [M.lidar_cloud(n, g, aerial=a) for n, a in spec]
    if cfg == 'wild-places':
        from hotformerloc_b200.datasets.coordinate_utils import cylindrical_for_octree
        clouds = [cylindrical_for_octree(c) for c in clouds]
    return clouds


# the fp32 CPU forward costs 1.5 - 3.5 minutes per case on 8 cores: two cases by default (the headline cfg and
# the cylindrical one), all four with HFL_SLOW_TESTS=1
_SLOW = os.environ.get('HFL_SLOW_TESTS', '0') == '1'


@pytest.mark.parametrize('name', ['oxford_b1_init', 'wp_b3_stress'] +
                         (['oxford_b1_stress', 'cswp_b6_stress'] if _SLOW else []))
def test_oracle_reproduces_reference_descriptors(golden_dir, name, tmp_path):
    cfg, depth, spec, seed, mode = CASES[name]
    gold = np.load(os.path.join(golden_dir, 'descriptors.npz'))
    shapes = json.load(open(os.path.join(golden_dir, f'state_shapes_{cfg}.json')))
    sd = M.synthetic_state_dict(shapes, mode=mode)
    hp = M.HParams.from_cfg(write_configs(str(tmp_path), cfg)['model_config'])
    o = R.build_batch(_clouds(name), depth)
    assert np.array_equal(o.nnum_nempty, gold[name + '_nnum_nempty'])
    g = M.forward(sd, o, hp).numpy()
    ref = gold[name + '_reference']
    assert np.abs(g - ref).max() < 1e-5
    cos = (g * ref).sum(1) / np.linalg.norm(g, axis=1) / np.linalg.norm(ref, axis=1)
    assert cos.min() > 0.99999


def test_staged_reference_reproduces_golden(golden_dir):
    """oracle/_ref (the reference's own models/*.py staged by oracle/build_ref.py -- the `--impl reference` arm and
    the cpu_baseline of bench.py) gives the frozen reference descriptors again: the staged copy is the code the
    goldens came from.  Skipped where it has not been staged (it is a build output, not part of the history)."""
    from oracle import build_ref, ocnn_standin as S
    if not build_ref.available():
        pytest.skip('oracle/_ref not staged (python -m oracle.build_ref needs /root/reference)')
    name = 'oxford_b1_init'
    cfg, depth, spec, seed, mode = CASES[name]
    S.install(build_ref.DST)
    m = S.reference_model(os.path.join(build_ref.DST, 'models', f'hotformerloc_{cfg}_cfg.txt'))
    shapes = json.load(open(os.path.join(golden_dir, f'state_shapes_{cfg}.json')))
    m.load_state_dict(M.synthetic_state_dict(shapes, mode=mode))
    with torch.inference_mode():
        y = m(S.make_batch(_clouds(name), depth))['global'].numpy()
    ref = np.load(os.path.join(golden_dir, 'descriptors.npz'))[name + '_reference']
    assert np.abs(y - ref).max() < 1e-6
