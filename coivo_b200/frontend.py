"""Front-end of the path: accept the raw outputs of the depth and pose
networks.  `poses_from_parameters` and `disp_to_depth` mirror Monodepth2's
`transformation_from_parameters` / `disp_to_depth` (oracle/frontend.py has the same signatures);
`photometric_loss_raw` chains them with the loss so a trainer passes sigmoid disparities and
axis-angle/translation pairs directly.  CUDA-only, through the C ABI (include/colvo.h).
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch

from . import _lib
from .loss import photometric_loss


def _cuda_f32(*ts):
    for t in ts:
        if not isinstance(t, torch.Tensor):
            raise TypeError("inputs must be torch.Tensor")
        if t.dtype != torch.float32:
            raise TypeError("inputs must be float32")
        if t.device.type != "cuda":
            raise ValueError("the front-end is CUDA-only (no CPU fallback)")
        if not t.is_contiguous():
            raise ValueError("inputs must be contiguous")


class _PoseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, mask):
        lib = _lib.load()
        B, N = axisangle.shape[:2]
        T = torch.empty(B, N, 4, 4, dtype=torch.float32, device=axisangle.device)
        with torch.cuda.device(axisangle.device):
            rc = lib.colvo_pose_from_axisangle(B, N, mask, axisangle.data_ptr(), translation.data_ptr(), T.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "colvo_pose_from_axisangle")
        ctx.save_for_backward(axisangle, translation)
        ctx.mask = mask
        return T

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gT):
        axisangle, translation = ctx.saved_tensors
        lib = _lib.load()
        B, N = axisangle.shape[:2]
        gT = gT.contiguous()
        ga, gt = torch.empty_like(axisangle), torch.empty_like(translation)
        with torch.cuda.device(axisangle.device):
            rc = lib.colvo_pose_from_axisangle_backward(B, N, ctx.mask, axisangle.data_ptr(), translation.data_ptr(),
                                                        gT.data_ptr(), ga.data_ptr(), gt.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "colvo_pose_from_axisangle_backward")
        return ga, gt, None


def poses_from_parameters(axisangle: torch.Tensor, translation: torch.Tensor, invert: Sequence[bool]) -> torch.Tensor:
    """`axisangle, translation [B,N,3]`, one `invert` flag per source frame -> `T [B,N,4,4]`."""
    _cuda_f32(axisangle, translation)
    if axisangle.dim() != 3 or axisangle.shape[-1] != 3 or axisangle.shape != translation.shape:
        raise ValueError("axisangle and translation must both be [B,N,3]")
    if len(invert) != axisangle.shape[1]:
        raise ValueError("one invert flag per source frame")
    mask = sum(1 << n for n, f in enumerate(invert) if f)
    return _PoseFn.apply(axisangle, translation, mask)


class _DispFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, min_depth, max_depth, *disp):
        lib = _lib.load()
        S = len(disp)
        depth = [torch.empty_like(d) for d in disp]
        counts = (ctypes.c_int64 * S)(*[d.numel() for d in disp])
        with torch.cuda.device(disp[0].device):
            rc = lib.colvo_disp_to_depth(S, counts, _lib.ptr_array([d.data_ptr() for d in disp]),
                                         _lib.ptr_array([d.data_ptr() for d in depth]), min_depth, max_depth,
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "colvo_disp_to_depth")
        ctx.save_for_backward(*depth)
        ctx.range = (min_depth, max_depth)
        return tuple(depth)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gdepth):
        depth = ctx.saved_tensors
        lib = _lib.load()
        S = len(depth)
        gdepth = [g.contiguous() for g in gdepth]
        gdisp = [torch.empty_like(d) for d in depth]
        counts = (ctypes.c_int64 * S)(*[d.numel() for d in depth])
        with torch.cuda.device(depth[0].device):
            rc = lib.colvo_disp_to_depth_backward(S, counts, _lib.ptr_array([d.data_ptr() for d in depth]),
                                                  _lib.ptr_array([g.data_ptr() for g in gdepth]),
                                                  _lib.ptr_array([g.data_ptr() for g in gdisp]), ctx.range[0], ctx.range[1],
                                                  torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "colvo_disp_to_depth_backward")
        return (None, None, *gdisp)


def disp_to_depth(disp: Sequence[torch.Tensor], min_depth: float = 0.1, max_depth: float = 100.0):
    """Sigmoid disparities (1..4 tensors) -> depths, `1 / (1/max_depth + (1/min_depth - 1/max_depth) * disp)`."""
    if not isinstance(disp, (list, tuple)) or not 1 <= len(disp) <= _lib.MAX_SCALES:
        raise ValueError(f"disp must be a sequence of 1..{_lib.MAX_SCALES} tensors")
    _cuda_f32(*disp)
    if not (min_depth > 0 and max_depth > min_depth):
        raise ValueError("need 0 < min_depth < max_depth")
    return list(_DispFn.apply(float(min_depth), float(max_depth), *disp))


def photometric_loss_raw(disp: Sequence[torch.Tensor], axisangle: torch.Tensor, translation: torch.Tensor, K, tgt, srcs, *,
                         invert: Sequence[bool] = (True, False), min_depth: float = 0.1, max_depth: float = 100.0, **kw):
    """The loss on raw network outputs: S sigmoid disparity maps and one (axis-angle, translation) pair per
    source frame.  `invert[n]` marks the sources whose pose is predicted source->target (Monodepth2 does so
    for the frame t-1)."""
    depth = disp_to_depth(disp, min_depth, max_depth)
    pose = poses_from_parameters(axisangle, translation, invert)
    return photometric_loss(depth, pose, K, tgt, srcs, **kw)
