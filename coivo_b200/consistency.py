"""Inference-time warp + LCC consistency sweep over a frame sequence (BASELINE config 5).

ColVO reconstructs the colon "by stitching together the dense depth maps of each frame using
the colonoscopic trajectory" (/root/reference/README.md:29); this sweep is the check that gates
that stitching: for every consecutive pair it re-projects frame t+1 into frame t with the
predicted depth and pose, recalibrates the brightness (LCC, README.md:7) and reports the
residual.  One batched launch over all F-1 pairs; forward only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def consistency(depth_seq: torch.Tensor, pose_seq: torch.Tensor, K: torch.Tensor, frames: torch.Tensor, *,
                lcc: bool = True) -> torch.Tensor:
    """depth_seq `[F,1,H,W]`, pose_seq `[F-1,4,4]` (T_{t->t+1}), K `[3,3]` or `[F-1,3,3]`,
    frames `[F,3,H,W]` -> `[F-1,4]` = (mean pe over valid pixels, a, b, valid fraction).
    Same signature as `oracle.photometric.consistency`.  CUDA-only."""
    if frames.dim() != 4 or frames.shape[1] != 3:
        raise ValueError("frames must be [F,3,H,W]")
    F, _, H, W = frames.shape
    if F < 2:
        raise ValueError("need at least two frames")
    if tuple(depth_seq.shape) != (F, 1, H, W):
        raise ValueError("depth_seq must be [F,1,H,W]")
    if tuple(pose_seq.shape) != (F - 1, 4, 4):
        raise ValueError("pose_seq must be [F-1,4,4]")
    if tuple(K.shape) not in ((3, 3), (F - 1, 3, 3)):
        raise ValueError("K must be [3,3] or [F-1,3,3]")
    for t in (depth_seq, pose_seq, K, frames):
        if t.device.type != "cuda":
            raise ValueError("consistency is CUDA-only (no CPU fallback)")
        if t.dtype != torch.float32:
            raise TypeError("inputs must be float32")
        if not t.is_contiguous():
            raise ValueError("inputs must be contiguous")
    lib = _lib.load()
    dev = frames.device
    n = ctypes.c_size_t()
    _lib.check(lib.colvo_consistency_workspace_bytes(F, H, W, ctypes.byref(n)), "colvo_consistency_workspace_bytes")
    with torch.cuda.device(dev):
        ws = torch.empty(max(n.value, 1), dtype=torch.uint8, device=dev)
        out = torch.empty(F - 1, 4, dtype=torch.float32, device=dev)
        rc = lib.colvo_consistency(F, H, W, _lib.F_LCC if lcc else 0, frames.data_ptr(), depth_seq.data_ptr(),
                                   pose_seq.data_ptr(), K.data_ptr(), 1 if K.dim() == 3 else 0, out.data_ptr(),
                                   ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "colvo_consistency")
    return out
