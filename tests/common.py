"""Shared helpers for the parity tests (inputs / weights are regenerated from
seeds on both sides; nothing here reads /root/reference)."""
import json
import os

import numpy as np
import torch

from oracle import model_ref as M
from oracle.make_golden import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def case_clouds(name):
    cfg, depth, spec, seed, mode = CASES[name]
    g = torch.Generator().manual_seed(seed)
    clouds = [M.lidar_cloud(n, g, aerial=a) for n, a in spec]
    if cfg == 'wild-places':
        from hotformerloc_b200.datasets.coordinate_utils import cylindrical_for_octree
        clouds = [cylindrical_for_octree(c) for c in clouds]
    return clouds


def case_state_dict(name):
    cfg, depth, spec, seed, mode = CASES[name]
    shapes = json.load(open(os.path.join(GOLDEN, f'state_shapes_{cfg}.json')))
    return M.synthetic_state_dict(shapes, mode=mode)


def native_model(cfg, sd, tmp_dir):
    from hotformerloc_b200.config.presets import write_configs
    from hotformerloc_b200.misc.utils import ModelParams
    from hotformerloc_b200.models.model_factory import model_factory
    paths = write_configs(str(tmp_dir), cfg)
    model = model_factory(ModelParams(paths['model_config']))
    model.load_state_dict(sd)
    return model.cuda().eval(), paths


def cosine(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
