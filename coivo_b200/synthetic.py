"""Deterministic synthetic colonoscopy-shaped inputs.

The upstream data (VCD / CSD sequences, `/root/reference/README.md:13`) is an external
link and there is no network, so every test and benchmark uses this generator.  All
tensors are produced on the CPU from `torch.Generator().manual_seed(seed)` so that the
CPU oracle and the CUDA path see identical bits.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch


def pyramid_shapes(H: int, W: int, S: int):
    return [(H >> k, W >> k) for k in range(S)]


def _so3_exp(omega: torch.Tensor) -> torch.Tensor:
    """Rodrigues formula in fp64: `[...,3]` -> `[...,3,3]`."""
    theta = omega.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = omega / theta
    kx, ky, kz = k.unbind(-1)
    zero = torch.zeros_like(kx)
    Kx = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=-1).reshape(*omega.shape[:-1], 3, 3)
    th = theta.unsqueeze(-1)
    eye = torch.eye(3, dtype=omega.dtype).expand_as(Kx)
    return eye + torch.sin(th) * Kx + (1 - torch.cos(th)) * (Kx @ Kx)


def make_intrinsics(B: int, H: int, W: int) -> torch.Tensor:
    K = torch.zeros(B, 3, 3, dtype=torch.float32)
    K[:, 0, 0] = 0.58 * W
    K[:, 1, 1] = 0.58 * W
    K[:, 0, 2] = W / 2
    K[:, 1, 2] = H / 2
    K[:, 2, 2] = 1.0
    return K


def make_poses(shape, g: torch.Generator, trans_sigma=0.02, rot_sigma=0.01) -> torch.Tensor:
    """Small SE(3) motions `[*shape,4,4]`: |omega| ~ rot_sigma, t ~ N(0, trans_sigma^2)."""
    om = torch.randn(*shape, 3, generator=g, dtype=torch.float64) * (rot_sigma / math.sqrt(3))
    t = torch.randn(*shape, 3, generator=g, dtype=torch.float64) * trans_sigma
    T = torch.zeros(*shape, 4, 4, dtype=torch.float64)
    T[..., :3, :3] = _so3_exp(om)
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T.to(torch.float32)


def make_target(B: int, H: int, W: int, g: torch.Generator) -> torch.Tensor:
    """Low-texture colour field + noise: 0.5 + 0.3 sin(2 pi (3x + c y)) + N(0, 0.05^2)."""
    y = torch.linspace(0, 1, H).reshape(1, 1, H, 1)
    x = torch.linspace(0, 1, W).reshape(1, 1, 1, W)
    c = torch.tensor([1.0, 2.0, 3.0]).reshape(1, 3, 1, 1)
    phase = torch.rand(B, 1, 1, 1, generator=g)
    base = 0.5 + 0.3 * torch.sin(2 * math.pi * (3 * x + c * y + phase))
    noise = 0.05 * torch.randn(B, 3, H, W, generator=g)
    return (base + noise).clamp_(0, 1).contiguous()


def make_triplets(B: int, H: int, W: int, N: int = 2, S: int = 4, seed: int = 0) -> Dict[str, object]:
    """One batch of frame triplets: `tgt [B,3,H,W]`, `srcs [B,N,3,H,W]`, `depth` list of
    `[B,1,h_k,w_k]`, `K [B,3,3]`, `pose [B,N,4,4]`.  Sources are the target rolled by
    -/+2 px, brightness-perturbed (0.9 I + 0.03 / 1.1 I - 0.02: exercises LCC) plus noise."""
    g = torch.Generator().manual_seed(seed)
    tgt = make_target(B, H, W, g)
    gains = [(0.9, 0.03), (1.1, -0.02), (0.95, 0.01), (1.05, -0.01)]
    srcs = []
    for n in range(N):
        shift = 2 * (-1) ** (n + 1) * (1 + n // 2)          # -2, +2, -4, +4
        a, b = gains[n % len(gains)]
        s = a * torch.roll(tgt, shifts=shift, dims=3) + b
        s = s + 0.02 * torch.randn(B, 3, H, W, generator=g)
        srcs.append(s.clamp_(0, 1))
    srcs = torch.stack(srcs, dim=1).contiguous()
    depth: List[torch.Tensor] = [
        (1.0 + 0.5 * torch.rand(B, 1, h, w, generator=g)).contiguous() for (h, w) in pyramid_shapes(H, W, S)
    ]
    return {
        "tgt": tgt,
        "srcs": srcs,
        "depth": depth,
        "K": make_intrinsics(B, H, W),
        "pose": make_poses((B, N), g),
    }


def make_sequence(F: int, H: int, W: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """A smooth random-walk sequence for the consistency sweep (BASELINE config 5):
    `frames [F,3,H,W]`, `depth [F,1,H,W]`, `pose [F-1,4,4]` (T_{t -> t+1}), `K [3,3]`."""
    g = torch.Generator().manual_seed(seed)
    base = make_target(1, H, W + 2 * F, g)[0]               # a long strip; the camera pans along it
    gain = 1.0 + 0.1 * torch.sin(torch.arange(F) * 0.37)
    bias = 0.02 * torch.cos(torch.arange(F) * 0.53)
    frames = torch.stack([base[:, :, 2 * t: 2 * t + W] for t in range(F)], dim=0)
    frames = frames * gain.reshape(F, 1, 1, 1) + bias.reshape(F, 1, 1, 1)
    frames = (frames + 0.01 * torch.randn(F, 3, H, W, generator=g)).clamp_(0, 1).contiguous()
    yy = torch.linspace(-1, 1, H).reshape(1, 1, H, 1)
    xx = torch.linspace(-1, 1, W).reshape(1, 1, 1, W)
    depth = (1.0 + 0.4 * (xx * xx + yy * yy) + 0.05 * torch.rand(F, 1, H, W, generator=g)).contiguous()
    return {
        "frames": frames,
        "depth": depth,
        "pose": make_poses((F - 1,), g),
        "K": make_intrinsics(1, H, W)[0],
    }
