"""CPU oracle for the step immediately upstream of the path.  TEST
INFRASTRUCTURE ONLY -- same rules as oracle/photometric.py (PARITY UNPINNED: the upstream repository
has no code; these are the Monodepth2 formulas recalled per assumption A0).

  * `transformation_from_parameters(axisangle, translation, invert)`: axis-angle (Rodrigues with the
    1e-7 guard on the angle) + translation -> 4x4 `T`; `invert=True` gives `R^T | -R^T t` (used for the
    frame t-1, whose network output is the motion source->target).
  * `disp_to_depth(disp, min_depth, max_depth)`: sigmoid output -> depth,
    `depth = 1 / (1/max_depth + (1/min_depth - 1/max_depth) * disp)`.
"""
from __future__ import annotations

from typing import Sequence

import torch


def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """`vec [...,3]` -> `[...,3,3]`."""
    angle = vec.norm(dim=-1, keepdim=True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle)[..., 0], torch.sin(angle)[..., 0]
    C = 1 - ca
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rows = [x * xC + ca, xyC - zs, zxC + ys, xyC + zs, y * yC + ca, yzC - xs, zxC - ys, yzC + xs, z * zC + ca]
    return torch.stack(rows, dim=-1).reshape(*vec.shape[:-1], 3, 3)


def transformation_from_parameters(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False) -> torch.Tensor:
    """`axisangle, translation [...,3]` -> `[...,4,4]`."""
    R = rot_from_axisangle(axisangle)
    t = translation
    if invert:
        R = R.transpose(-1, -2)
        t = -(R @ t.unsqueeze(-1)).squeeze(-1)
    T = torch.zeros(*axisangle.shape[:-1], 4, 4, dtype=axisangle.dtype)
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T


def poses_from_parameters(axisangle: torch.Tensor, translation: torch.Tensor, invert: Sequence[bool]) -> torch.Tensor:
    """`[B,N,3]` each, `invert` one flag per source frame -> `[B,N,4,4]`."""
    return torch.stack([transformation_from_parameters(axisangle[:, n], translation[:, n], bool(invert[n]))
                        for n in range(axisangle.shape[1])], dim=1)


def disp_to_depth(disp: torch.Tensor, min_depth: float = 0.1, max_depth: float = 100.0) -> torch.Tensor:
    min_disp, max_disp = 1.0 / max_depth, 1.0 / min_depth
    return 1.0 / (min_disp + (max_disp - min_disp) * disp)
