"""Host-side configuration objects with the attribute names of the reference's
``misc/utils.py`` (``ModelParams`` :15-116, ``TrainingParams`` :118-255,
``set_seed`` :281, ``rescale_octree_points`` :293) so that reference INI files
and call sites work unchanged.  Parsing only -- no training logic."""
from __future__ import annotations

import configparser
import os
import random
import time

import numpy as np
import torch


def _ints(s):
    return tuple(int(e) for e in s.split(','))


class ModelParams:
    def __init__(self, model_params_path):
        cp = configparser.ConfigParser()
        if not cp.read(model_params_path):
            raise FileNotFoundError(model_params_path)
        p = cp['MODEL']
        self.model_params_path = model_params_path
        self.model = p.get('model')
        self.output_dim = p.getint('output_dim', 256)
        self.coordinates = p.get('coordinates', 'polar')
        assert self.coordinates in ['polar', 'cartesian', 'cylindrical'], \
            f'Unsupported coordinates: {self.coordinates}'
        if self.coordinates == 'cartesian':
            self.quantizer = None
        elif self.coordinates == 'cylindrical':
            from ..datasets.coordinate_utils import CylindricalCoordinates
            self.quantizer = CylindricalCoordinates(use_octree=True)
        else:
            raise NotImplementedError(f'Unsupported coordinates: {self.coordinates}')
        self.normalize_embeddings = p.getboolean('normalize_embeddings', False)
        self.feature_size = p.getint('feature_size', 256)
        self.pooling = p.get('pooling', 'OctGeM')
        self.num_top_down = p.getint('num_top_down', 1)
        self.channels = _ints(p['channels']) if 'channels' in p else (96, 192, 384, 384)
        self.num_blocks = _ints(p['num_blocks']) if 'num_blocks' in p else (2, 2, 6, 2)
        self.num_heads = _ints(p['num_heads']) if 'num_heads' in p else None
        self.patch_size = p.getint('patch_size', 32)
        self.dilation = p.getint('dilation', 4)
        self.ct_size = p.getint('ct_size', 1)
        self.ct_propagation = p.getboolean('ct_propagation', False)
        self.ct_propagation_scale = p.getfloat('ct_propagation_scale', None)
        self.ADaPE_mode = p.get('ADaPE_mode', None)
        if self.ADaPE_mode == 'None':
            self.ADaPE_mode = None
        self.drop_path = p.getfloat('drop_path', 0.5)
        self.input_features = p.get('input_features', 'P')
        self.downsample_input_embeddings = p.getboolean('downsample_input_embeddings', True)
        self.num_input_downsamples = p.getint('num_input_downsamples', 2)
        self.disable_RPE = p.getboolean('disable_RPE', False)
        self.conv_norm = p.get('conv_norm', 'batchnorm')
        assert self.conv_norm in ['batchnorm', 'layernorm', 'powernorm']
        self.layer_scale = p.getfloat('layer_scale', None)
        self.grad_checkpoint = p.getboolean('grad_checkpoint', True)
        if 'qkv_init' in p:
            self.qkv_init = list(p['qkv_init'].split(','))
            if len(self.qkv_init) > 1:
                self.qkv_init[1] = None if self.qkv_init[1] == 'None' else float(self.qkv_init[1])
        else:
            self.qkv_init = ['trunc_normal', 0.02]
        self.xcpe = p.getboolean('xCPE', False)
        if 'hotformerloc' in self.model.lower():
            self.num_pyramid_levels = p.getint('num_pyramid_levels', 3)
            self.num_octf_levels = p.getint('num_octf_levels', 1)
            k = p.get('k_pooled_tokens', '64')
            self.k_pooled_tokens = int(k) if k.isdigit() else _ints(k)
            self.disable_rt = p.getboolean('disable_rt', False)
        else:
            self.ct_layers = tuple(e == 'True' for e in p['ct_layers'].split(',')) \
                if 'ct_layers' in p else tuple([False] * len(self.channels))

    def print(self):
        print('Model parameters:')
        for k, v in vars(self).items():
            print(f'{k}: {v}')
        print('')


class TrainingParams:
    """Evaluation reads: dataset_folder, val_batch_size, normalize_points,
    scale_factor, unit_sphere_norm, octree_depth, dataset_name, skip_same_run,
    load_octree, debug, model_params (eval/pnv_evaluate.py:129-187)."""

    def __init__(self, params_path: str, model_params_path: str, debug: bool = False,
                 verbose: bool = False):
        assert os.path.exists(params_path), f'Cannot find configuration file: {params_path}'
        assert os.path.exists(model_params_path), \
            f'Cannot find model-specific configuration file: {model_params_path}'
        self.params_path, self.model_params_path = params_path, model_params_path
        self.debug, self.verbose = debug, verbose
        cp = configparser.ConfigParser()
        cp.read(params_path)
        self.dataset_folder = cp['DEFAULT'].get('dataset_folder')
        p = cp['TRAIN']
        self.save_freq = p.getint('save_freq', 0)
        self.eval_freq = p.getint('eval_freq', 0)
        self.num_workers = p.getint('num_workers', 0)
        self.wandb = p.getboolean('wandb', True)
        self.batch_size = p.getint('batch_size', 64)
        self.batch_split_size = p.getint('batch_split_size', None)
        self.batch_expansion_th = p.getfloat('batch_expansion_th', None)
        if self.batch_expansion_th is not None:
            assert 0. < self.batch_expansion_th < 1.
            self.batch_size_limit = p.getint('batch_size_limit', 256)
            self.batch_expansion_rate = p.getfloat('batch_expansion_rate', 1.5)
        else:
            self.batch_size_limit = self.batch_size
            self.batch_expansion_rate = None
        self.val_batch_size = p.getint('val_batch_size', self.batch_size_limit)
        self.lr = p.getfloat('lr', 1e-3)
        self.epochs = p.getint('epochs', 20)
        self.warmup_epochs = p.getint('warmup_epochs', None)
        self.optimizer = p.get('optimizer', 'Adam')
        self.scheduler = p.get('scheduler', 'MultiStepLR')
        self.gamma = p.getfloat('gamma', 0.1)
        self.scheduler_milestones = [int(e) for e in p.get('scheduler_milestones',
                                                           str(self.epochs + 1)).split(',')]
        self.weight_decay = p.getfloat('weight_decay', None)
        self.loss = (p.get('loss') or '').lower()
        self.similarity = p.get('similarity', 'euclidean')
        self.aug_mode = p.getint('aug_mode', 1)
        self.set_aug_mode = p.getint('set_aug_mode', 1)
        self.random_rot_theta = p.getfloat('random_rot_theta', 5.0)
        self.normalize_points = p.getboolean('normalize_points', False)
        self.scale_factor = p.getfloat('scale_factor', None)
        self.unit_sphere_norm = p.getboolean('unit_sphere_norm', False)
        self.zero_mean = p.getboolean('zero_mean', True)
        self.octree_depth = p.getint('octree_depth', 11)
        self.full_depth = p.getint('full_depth', 2)
        self.train_file = p.get('train_file')
        self.val_file = p.get('val_file', None)
        self.validation = p.getboolean('validation', True)
        self.test_file = p.get('test_file', None)
        self.dataset_name = p.get('dataset_name', None)
        self.skip_same_run = p.getboolean('skip_same_run', True)
        self.mesa = p.getfloat('mesa', 0.0)
        self.mesa_start_ratio = p.getfloat('mesa_start_ratio', 0.25)
        self.model_params = ModelParams(self.model_params_path)
        self.load_octree = any(m in self.model_params.model.lower()
                               for m in ('octformer', 'hotformer'))
        self.hyperparam_search = p.getboolean('hyperparam_search', False)
        self._check_params()

    def _check_params(self):
        assert os.path.exists(self.dataset_folder), f'Cannot access dataset: {self.dataset_folder}'

    def print(self):
        print('Parameters:')
        for k, v in vars(self).items():
            if k != 'model_params':
                print(f'{k}: {v}')
        self.model_params.print()
        print('')


def get_datetime():
    return time.strftime('%Y%m%d_%H%M')


def set_seed(seed: int = 42):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    print('Determinism: Enabled')


def rescale_octree_points(points: torch.Tensor, depth: int) -> torch.Tensor:
    """[0, 2^d] octree units -> [-1, 1]."""
    return points * (2 ** (1 - depth)) - 1.0
