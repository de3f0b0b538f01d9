"""``to_device`` / ``release_cuda`` with the reference's semantics
(misc/torch_utils.py:12-56) for the native octree object."""
import numpy as np
import torch

from ..octree import Octree


def release_cuda(x, to_numpy=False):
    if isinstance(x, (list, tuple)):
        return type(x)(release_cuda(i, to_numpy) for i in x)
    if isinstance(x, dict):
        return {k: release_cuda(v, to_numpy) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        if x.numel() == 1:
            return x.item()
        x = x.detach().cpu()
        return x.numpy() if to_numpy else x
    return x


def to_device(x, device, non_blocking=False, construct_octree_neigh=False):
    if isinstance(x, (list, tuple)):
        return type(x)(to_device(i, device, non_blocking, construct_octree_neigh) for i in x)
    if isinstance(x, dict):
        return {k: to_device(v, device, non_blocking, construct_octree_neigh) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        return x.to(device=device, non_blocking=non_blocking)
    if isinstance(x, Octree):
        x = x.to(device)
        if construct_octree_neigh:
            x.construct_all_neigh()
    return x
