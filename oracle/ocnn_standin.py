"""ORACLE (test infrastructure, not product code) -- container-only helper.

Registers ``sys.modules`` stand-ins for the third-party packages the reference
imports but that are absent/un-installable offline (``ocnn==2.2.2``,
``dwconv`` (CUDA-only, libs/dwconv), ``matplotlib``, ``open3d``) plus explicit
namespace packages for the reference's ``models/misc/datasets/eval/libs``
directories, so that the reference's *own, unmodified* Python
(``/root/reference/models/*.py`` ...) can be imported and run on CPU.  This is
the recipe of SURVEY.md Appendix C.  It is used ONLY in this container by
``oracle/make_golden.py`` to (a) validate ``oracle/model_ref.py`` against the
real reference code and (b) freeze golden vectors into ``tests/golden``.
Nothing on the GPU box imports it (``/root/reference`` does not exist there).

The octree arithmetic behind the stand-in is ``oracle/octree_ref.py`` (pinned
bit-exact against the reference's fixtures).  Float ops follow SURVEY.md
Appendix A "Float ops".
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from . import octree_ref as R

REFERENCE_ROOT = '/root/reference'


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dtype) if dtype is not None else t


class Points:
    def __init__(self, points, normals=None, features=None, labels=None,
                 batch_id=None, batch_size=1):
        self.points = points
        self.batch_size = batch_size


class Octree:
    def __init__(self, depth, full_depth=2, batch_size=1, device='cpu', **kw):
        self.depth = int(depth)
        self.full_depth = int(full_depth)
        self.batch_size = batch_size
        self.device = torch.device(device)
        n = self.depth + 1
        self.keys = [None] * n
        self.children = [None] * n
        self.neighs = [None] * n
        self.points = [None] * n
        self.features = [None] * n
        self.normals = [None] * n
        self.nnum = torch.zeros(n, dtype=torch.long)
        self.nnum_nempty = torch.zeros(n, dtype=torch.long)
        self.batch_nnum = None
        self.batch_nnum_nempty = None
        self._ref = None

    def _load(self, ref: R.RefOctree):
        self._ref = ref
        self.batch_size = ref.batch_size
        for d in range(self.depth + 1):
            self.keys[d] = _t(ref.keys[d])
            self.children[d] = _t(ref.children[d])
            if ref.neighs[d] is not None:
                self.neighs[d] = _t(ref.neighs[d])
        self.points[self.depth] = _t(ref.points[self.depth])
        self.nnum = _t(ref.nnum)
        self.nnum_nempty = _t(ref.nnum_nempty)
        self.batch_nnum = _t(ref.batch_nnum)
        self.batch_nnum_nempty = _t(ref.batch_nnum_nempty)
        return self

    def build_octree(self, point_cloud: Points):
        ref = R.RefOctree(self.depth, self.full_depth)
        idx = ref.build_octree(point_cloud.points.detach().cpu().numpy())
        self._load(ref)
        return _t(idx)

    def construct_all_neigh(self):
        self._ref.construct_all_neigh()
        for d in range(1, self.depth + 1):
            self.neighs[d] = _t(self._ref.neighs[d])

    def nempty_mask(self, depth):
        return self.children[depth] >= 0

    def key(self, depth, nempty=False):
        key = self.keys[depth]
        return key[self.nempty_mask(depth)] if nempty else key

    def xyzb(self, depth, nempty=False):
        return key2xyz(self.key(depth, nempty), depth)

    def batch_id(self, depth, nempty=False):
        return self.key(depth, nempty) >> 48

    def get_neigh(self, depth, kernel='333', stride=1, nempty=False):
        if isinstance(kernel, (list, tuple)):
            kernel = ''.join(str(k) for k in kernel)
        return _t(self._ref.get_neigh(depth, kernel, stride, nempty))

    def to(self, device, non_blocking=False):
        return self

    def cpu(self):
        return self


def merge_octrees(octrees):
    ref = R.merge_octrees([o._ref for o in octrees])
    out = Octree(ref.depth, ref.full_depth, batch_size=len(octrees))
    return out._load(ref)


def xyz2key(x, y, z, b=None, depth=16):
    bb = None if b is None else (b.numpy() if isinstance(b, torch.Tensor) else b)
    return _t(R.xyz2key(x.numpy(), y.numpy(), z.numpy(), bb, depth))


def key2xyz(key, depth=16):
    x, y, z, b = R.key2xyz(key.numpy(), depth)
    return _t(x), _t(y), _t(z), _t(b)


def _kernel_str(kernel_size):
    ks = list(kernel_size)
    if len(ks) == 1:
        ks = ks * 3
    return ''.join(str(k) for k in ks)


def _gather(data, neigh):
    """(rows, kdim, C) buffer, zeros where neigh < 0 (SURVEY App. A Float ops)."""
    buf = data.new_zeros(neigh.shape[0], neigh.shape[1], data.shape[1])
    valid = neigh >= 0
    buf[valid] = data[neigh[valid]]
    return buf


class OctreeConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=[3], stride=1,
                 nempty=False, direct_method=False, use_bias=False, max_buffer=int(2e8)):
        super().__init__()
        self.kernel = _kernel_str(kernel_size)
        self.kdim = len(R.LUT_KERNEL[self.kernel])
        self.stride, self.nempty, self.use_bias = stride, nempty, use_bias
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weights = torch.nn.Parameter(torch.empty(self.kdim, in_channels, out_channels))
        torch.nn.init.xavier_uniform_(self.weights)
        if use_bias:
            self.bias = torch.nn.Parameter(torch.zeros(out_channels))

    def forward(self, data, octree, depth):
        neigh = octree.get_neigh(depth, self.kernel, self.stride, self.nempty)
        out = _gather(data, neigh).flatten(1) @ self.weights.flatten(0, 1)
        if self.use_bias:
            out = out + self.bias
        return out


class OctreeDeconv(OctreeConv):
    def forward(self, data, octree, depth):  # not on the hot path
        raise NotImplementedError


class OctreeDWConv(torch.nn.Module):
    def __init__(self, in_channels, kernel_size=[3], stride=1, nempty=False, use_bias=False):
        super().__init__()
        self.kernel = _kernel_str(kernel_size)
        self.kdim = len(R.LUT_KERNEL[self.kernel])
        self.stride, self.nempty, self.use_bias = stride, nempty, use_bias
        self.weights = torch.nn.Parameter(torch.empty(self.kdim, 1, in_channels))
        torch.nn.init.xavier_uniform_(self.weights)
        if use_bias:
            self.bias = torch.nn.Parameter(torch.zeros(in_channels))

    def forward(self, data, octree, depth):
        neigh = octree.get_neigh(depth, self.kernel, self.stride, self.nempty)
        out = torch.einsum('ikc,kc->ic', _gather(data, neigh), self.weights[:, 0, :])
        if self.use_bias:
            out = out + self.bias
        return out


class DWConvCuda(OctreeDWConv):
    """stand-in for libs/dwconv/dwconv/nn.py:49 (same math as dwconv.cu:30-41)."""
    def __init__(self, channels, kernel_size=[3], nempty=False, use_bias=False):
        super().__init__(channels, kernel_size, 1, nempty, use_bias)


class OctreeGlobalPool(torch.nn.Module):
    def __init__(self, nempty=False):
        super().__init__()
        self.nempty = nempty

    def forward(self, data, octree, depth):
        bid = octree.batch_id(depth, self.nempty)
        B = octree.batch_size
        out = data.new_zeros(B, data.shape[1]).index_add_(0, bid, data)
        cnt = torch.bincount(bid, minlength=B).clamp(min=1).to(data.dtype)
        return out / cnt[:, None]


class InputFeature(torch.nn.Module):
    def __init__(self, feature='P', nempty=True):
        super().__init__()
        assert feature == 'P'

    def forward(self, octree):
        D = octree.depth
        return octree.points[D] * (2 ** (1 - D)) - 1.0


def install(reference_root: str = REFERENCE_ROOT):
    """Register the stand-ins; idempotent."""
    if 'ocnn' in sys.modules and getattr(sys.modules['ocnn'], '_hfl_standin', False):
        return
    ocnn = types.ModuleType('ocnn')
    ocnn._hfl_standin = True
    m_oct = types.ModuleType('ocnn.octree')
    for n, v in dict(Octree=Octree, Points=Points, merge_octrees=merge_octrees,
                     key2xyz=key2xyz, xyz2key=xyz2key).items():
        setattr(m_oct, n, v)
    m_nn = types.ModuleType('ocnn.nn')
    for n, v in dict(OctreeConv=OctreeConv, OctreeDeconv=OctreeDeconv,
                     OctreeDWConv=OctreeDWConv, OctreeGlobalPool=OctreeGlobalPool).items():
        setattr(m_nn, n, v)
    m_mod = types.ModuleType('ocnn.modules')
    m_mod.InputFeature = InputFeature
    ocnn.octree, ocnn.nn, ocnn.modules = m_oct, m_nn, m_mod
    sys.modules.update({'ocnn': ocnn, 'ocnn.octree': m_oct, 'ocnn.nn': m_nn,
                        'ocnn.modules': m_mod})
    dw = types.ModuleType('dwconv')
    dw.OctreeDWConv = DWConvCuda
    sys.modules['dwconv'] = dw
    for name in ('matplotlib', 'matplotlib.pyplot', 'open3d', 'tqdm'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if hasattr(sys.modules.get('matplotlib'), '__dict__') and 'matplotlib.pyplot' in sys.modules:
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    for n in ('datasets', 'models', 'misc', 'eval', 'libs'):
        m = types.ModuleType(n)
        m.__path__ = [f'{reference_root}/{n}']
        sys.modules[n] = m
    for n in ('models.layers', 'models.losses', 'datasets.pointnetvlad', 'datasets.CSWildPlaces'):
        m = types.ModuleType(n)
        m.__path__ = [f'{reference_root}/{n.replace(".", "/")}']
        sys.modules[n] = m


def reference_model(model_cfg_path: str):
    """model_factory(ModelParams(cfg)) using the reference's own code."""
    install()
    from misc.utils import ModelParams
    from models.model_factory import model_factory
    return model_factory(ModelParams(model_cfg_path)).eval()


def make_batch(clouds, depth, full_depth=2):
    """collate_batch of eval/pnv_evaluate.py:122-126 on the stand-in."""
    install()
    octs = []
    for c in clouds:
        o = Octree(depth, full_depth)
        o.build_octree(Points(torch.as_tensor(c)))
        octs.append(o)
    m = merge_octrees(octs)
    m.construct_all_neigh()
    return {'octree': m}
