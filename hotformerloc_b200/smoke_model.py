"""Used by __graft_entry__.smoke(): one small forward of the hot path on cuda:0,
checked against the CPU oracle."""
import json
import os
import tempfile

import numpy as np
import torch


def run(clouds, octree, ref_octree):
    from oracle import model_ref as M
    from .config.presets import write_configs
    from .misc.utils import ModelParams
    from .models.model_factory import model_factory
    from . import native
    d = tempfile.mkdtemp(prefix='hfl_smoke_')
    paths = write_configs(d, 'oxford', dataset_folder=d)
    torch.manual_seed(0)
    model = model_factory(ModelParams(paths['model_config'])).cuda().eval()
    l0 = native.launch_count()
    y = model({'octree': octree})['global'].float().cpu().numpy()
    launches = native.launch_count() - l0
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = M.forward(sd, ref_octree, M.HParams.from_cfg(paths['model_config'])).numpy()
    cos = (y * g).sum(1) / np.linalg.norm(y, axis=1) / np.linalg.norm(g, axis=1)
    print(f'smoke: descriptors {y.shape}, {launches} kernel launches, min cosine vs oracle '
          f'{cos.min():.6f}, max-abs {np.abs(y - g).max():.2e}')
    assert cos.min() >= 0.999
