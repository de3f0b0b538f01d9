"""Pre-octree point transforms used on the evaluation path (host side; SURVEY
section 8 row a0).  The exactness contract of the octree starts at the fp32
(P,3) tensor these functions return, so they reproduce the reference's
arithmetic operation by operation:
``CylindricalCoordinates`` datasets/coordinate_utils.py:68-116 (atan2/sqrt in
fp32 torch, rescale through fp64 ``np.interp``, clamp) and ``Normalize``
datasets/augmentation.py:185-235."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


class CylindricalCoordinates:
    def __init__(self, use_octree: bool = True):
        self.use_octree = use_octree

    def __call__(self, pc: torch.Tensor) -> torch.Tensor:
        assert pc.ndim == 2 and pc.shape[1] == 3
        assert torch.all(abs(pc) <= 1.0)
        phi = torch.atan2(pc[:, 1], pc[:, 0])
        rho = torch.sqrt(pc[:, 0] ** 2 + pc[:, 1] ** 2)
        out = torch.stack([rho, phi, pc[:, 2]], dim=1)
        if self.use_octree:
            # fp64 linear rescale ([0,1]->[-1,1], [-pi,pi]->[-1,1]) written back into fp32
            out[:, 0] = torch.tensor(np.interp(out[:, 0].numpy(), [0, 1], [-1, 1]))
            out[:, 1] = torch.tensor(np.interp(out[:, 1].numpy(), [-np.pi, np.pi], [-1, 1]))
            out = torch.clamp(out, -1.0, 1.0)
        return out


def cylindrical_for_octree(cloud: np.ndarray) -> np.ndarray:
    """eval/pnv_evaluate.py:166-171: radial mask, then cylindrical + rescale."""
    data = torch.from_numpy(np.ascontiguousarray(cloud, dtype=np.float32))
    data = data[torch.linalg.norm(data[:, :2], dim=1) <= 1.0]
    return CylindricalCoordinates(True)(data).numpy()


class Normalize:
    """Bounding-box / unit-sphere / fixed-scale normalisation."""

    def __init__(self, norm_range: Optional[float] = None, scale_factor: Optional[float] = None,
                 unit_sphere_norm: bool = False, zero_mean: bool = True):
        assert norm_range is None or scale_factor is None
        self.norm_range = 1.0 if scale_factor is None else None
        if norm_range is not None:
            assert norm_range > 0
            self.norm_range = norm_range
        self.scale_factor = scale_factor
        self.unit_sphere_norm, self.zero_mean = unit_sphere_norm, zero_mean

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        if not self.unit_sphere_norm:
            lo, hi = coords.min(dim=0).values, coords.max(dim=0).values
            if self.zero_mean:
                coords = coords - (lo + hi) * 0.5
            if self.scale_factor is not None:
                return coords / self.scale_factor
            return coords * (2.0 * self.norm_range / ((hi - lo).max() + 1.0e-6))
        if self.zero_mean:
            coords = coords - torch.mean(coords, axis=0)
        if self.scale_factor is not None:
            return coords / self.scale_factor
        return coords / (torch.max(torch.linalg.norm(coords, dim=1)) / self.norm_range)
