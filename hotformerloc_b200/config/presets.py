"""The reference's shipped configurations (``config/config_*.txt`` and
``models/hotformerloc_*_cfg.txt``) as data, plus a writer that materialises
them as INI files with the reference's exact ``[MODEL] / [DEFAULT] / [TRAIN]``
schema (parsed by hotformerloc_b200.misc.utils, mirroring misc/utils.py:15-255).
User-supplied INI files of the reference work unchanged."""
from __future__ import annotations

import os
from typing import Dict

_MODEL_COMMON = dict(
    channels='128,256', num_blocks='4,10', num_heads='8,16', num_pyramid_levels=3, ct_size=1,
    ct_propagation=False, input_features='P', downsample_input_embeddings=True,
    num_input_downsamples=2, disable_RPE=False, grad_checkpoint=True, conv_norm='layernorm',
    feature_size=256, output_dim=256, pooling='PyramidAttnPoolMixer', normalize_embeddings=True)

MODEL_PRESETS: Dict[str, dict] = {
    'oxford': dict(_MODEL_COMMON, model='HOTFormerLoc-Oxford', ADaPE_mode='cov', patch_size=48,
                   k_pooled_tokens='74,36,18', coordinates='cartesian'),
    'cs-wild-places': dict(_MODEL_COMMON, model='HOTFormerLoc-CSWildPlaces', ADaPE_mode='cov',
                           patch_size=64, k_pooled_tokens='74,36,18', coordinates='cartesian'),
    'cs-campus3d': dict(_MODEL_COMMON, model='HOTFormerLoc-CSCampus3D', ADaPE_mode='cov',
                        patch_size=64, k_pooled_tokens='148,72,36', coordinates='cartesian'),
    'wild-places': dict(_MODEL_COMMON, model='HOTFormerLoc-WildPlaces', patch_size=48,
                        k_pooled_tokens='148,72,36', coordinates='cylindrical'),
}

_TRAIN_COMMON = dict(num_workers=2, batch_size=2048, batch_split_size=128, save_freq=10,
                     eval_freq=5, wandb=True, lr='5e-4', epochs=150, scheduler_milestones=100,
                     warmup_epochs=5, aug_mode=1, set_aug_mode=1, weight_decay='1e-4',
                     loss='TruncatedSmoothAP', tau1=0.01, positives_per_query=4,
                     skip_same_run=True, validation=True)

TRAIN_PRESETS: Dict[str, dict] = {
    'oxford': dict(_TRAIN_COMMON, val_batch_size=256, normalize_points=False, octree_depth=9,
                   dataset_name='Oxford', train_file='training_queries_baseline2.pickle',
                   val_file='test_queries_baseline2.pickle'),
    'cs-wild-places': dict(_TRAIN_COMMON, val_batch_size=128, normalize_points=True,
                           octree_depth=7, dataset_name='CSWildPlaces',
                           train_file='training_queries_CSWildPlaces_baseline_v2.pickle',
                           val_file='test_queries_CSWildPlaces_v2.pickle'),
    'cs-campus3d': dict(_TRAIN_COMMON, val_batch_size=256, normalize_points=False, octree_depth=7,
                        dataset_name='CSCampus3D', skip_same_run=False, validation=False,
                        train_file='training_queries_umd_4096_v2.pickle'),
    'wild-places': dict(_TRAIN_COMMON, val_batch_size=128, normalize_points=True, octree_depth=7,
                        dataset_name='WildPlaces', validation=False,
                        train_file='training_wild-places.pickle'),
}


def _ini(section: str, d: dict) -> str:
    return f'[{section}]\n' + ''.join(f'{k}={v}\n' for k, v in d.items())


def write_configs(out_dir: str, name: str, dataset_folder: str = '.') -> Dict[str, str]:
    """Write ``config_<name>.txt`` and ``hotformerloc_<name>_cfg.txt`` into out_dir."""
    os.makedirs(out_dir, exist_ok=True)
    mp = os.path.join(out_dir, f'hotformerloc_{name}_cfg.txt')
    cp = os.path.join(out_dir, f'config_{name}.txt')
    with open(mp, 'w') as f:
        f.write(_ini('MODEL', MODEL_PRESETS[name]))
    with open(cp, 'w') as f:
        f.write(_ini('DEFAULT', {'dataset_folder': dataset_folder}))
        f.write(_ini('TRAIN', TRAIN_PRESETS[name]))
    return {'model_config': mp, 'config': cp}
