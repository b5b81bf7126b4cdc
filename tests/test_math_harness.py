"""Checks the per-pixel formulas in coivo_b200/csrc/colvo_math.cuh (the header the CUDA kernels
include) against the oracle on the CPU, through the single-threaded g++ harness under
tests/cpu_harness/.  Test scaffolding only; the product path never touches this harness."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import photometric as O
from coivo_b200.synthetic import make_triplets

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_harness", "harness.cpp")
OUT = os.path.join(HERE, "cpu_harness", "_build", "libharness.so")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdrs = [os.path.join(HERE, "..", "coivo_b200", "csrc", h) for h in ("colvo_math.cuh", "colvo_f2.cuh", "colvo_pe.cuh")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max([os.path.getmtime(SRC)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-std=c++17", SRC, "-o", OUT])
    return ctypes.CDLL(OUT)


def _fp(t):
    return t.data_ptr()


def run_harness(lib, d, N, S, flags):
    B, _, H, W = d["tgt"].shape
    HW = H * W
    loss = torch.zeros(1)
    valid = torch.zeros(B, N, S, H, W, dtype=torch.uint8)
    sel = torch.zeros(B, S, H, W, dtype=torch.uint8)
    ab = torch.zeros(B, N, S, 2)
    gdepth = [torch.zeros_like(x) for x in d["depth"]]
    gT = torch.zeros(B, N, 4, 4)
    gsrc = torch.zeros_like(d["srcs"])
    PtrArr = ctypes.c_void_p * 4
    dp = PtrArr(*([_fp(x) for x in d["depth"]] + [None] * (4 - S)))
    gp = PtrArr(*([_fp(x) for x in gdepth] + [None] * (4 - S)))
    lib.harness_run.restype = ctypes.c_int
    rc = lib.harness_run(
        ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(S), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_uint(flags),
        ctypes.c_void_p(_fp(d["tgt"])), ctypes.c_void_p(_fp(d["srcs"])), dp, ctypes.c_void_p(_fp(d["K"])),
        ctypes.c_void_p(_fp(d["pose"])), ctypes.c_void_p(_fp(loss)), ctypes.c_void_p(_fp(valid)),
        ctypes.c_void_p(_fp(sel)), ctypes.c_void_p(_fp(ab)), gp, ctypes.c_void_p(_fp(gT)), ctypes.c_void_p(_fp(gsrc)))
    assert rc == 0
    return loss, valid, sel, ab, gdepth, gT, gsrc


def relinf(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("B,H,W,N,S,flags", [(2, 16, 24, 2, 3, 1), (1, 19, 27, 2, 4, 1), (1, 16, 24, 1, 2, 3), (1, 12, 20, 2, 2, 0)])
def test_harness_matches_oracle(harness, B, H, W, N, S, flags):
    d = make_triplets(B, H, W, N=N, S=S, seed=11)
    lcc, detach = bool(flags & 1), bool(flags & 2)
    loss, valid, sel, ab, gdepth, gT, gsrc = run_harness(harness, d, N, S, flags)
    depth = [x.clone().requires_grad_() for x in d["depth"]]
    pose = d["pose"].clone().requires_grad_()
    srcs = d["srcs"].clone().requires_grad_()
    with torch.no_grad():
        l0, v0, s0, ab0 = O.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], smooth_weight=0.0,
                                             lcc=lcc, lcc_detach=detach, return_masks=True)
        gap = O.candidate_gap(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], lcc=lcc)
    assert torch.equal(valid, v0), "valid mask must be bit-exact"
    assert torch.allclose(ab, ab0, rtol=1e-5, atol=1e-6)
    mism = sel != s0
    assert (gap[mism] < 1e-4).all(), "sel may differ only at near-ties"
    assert abs(loss.item() - l0.item()) <= 1e-5 * abs(l0.item())
    l1 = O.photometric_loss(depth, pose, d["K"], d["tgt"], srcs, smooth_weight=0.0, lcc=lcc, lcc_detach=detach,
                            sel_override=sel, ab_override=ab)
    l1.backward()
    for k in range(S):
        assert relinf(gdepth[k], depth[k].grad) < 1e-4, f"grad_depth[{k}]"
    assert relinf(gT[:, :, :3], pose.grad[:, :, :3]) < 1e-4
    assert gT[:, :, 3].abs().max() == 0
    assert relinf(gsrc, srcs.grad) < 1e-4


def test_packed_window_evaluation_matches_scalar_reference(harness):
    """colvo_pe.cuh (what k_photo_fwd runs, sources in the two lanes of a packed register) against the scalar
    per-channel formulas of colvo_math.cuh, which the test above ties to the oracle's autograd: value, dpe/da, dpe/db
    and the SSIM adjoint coefficients of both sources, including windows inside and outside the clamp."""
    g = torch.Generator().manual_seed(3)
    nw = 4096
    y = torch.rand(nw, 3, 9, generator=g)
    x = (y.unsqueeze(1) + 0.2 * torch.randn(nw, 2, 3, 9, generator=g)).contiguous()
    x[: nw // 8] = 1.0 - y[: nw // 8].unsqueeze(1)            # anti-correlated windows: SSIM < 0, some beyond the clamp
    x[nw // 8: nw // 4] = x[nw // 8: nw // 4].mean(dim=-1, keepdim=True)   # flat windows: variance ~ 0 against C2
    ab = torch.stack([0.8 + 0.4 * torch.rand(nw, 2, generator=g), 0.1 * torch.randn(nw, 2, generator=g)], dim=-1).contiguous()
    ref, new, val = torch.zeros(nw, 2, 12), torch.zeros(nw, 2, 12), torch.zeros(nw, 2)
    one = torch.zeros(nw, 12)
    harness.harness_pe_fused.restype = ctypes.c_int
    rc = harness.harness_pe_fused(ctypes.c_int(nw), ctypes.c_void_p(_fp(x)), ctypes.c_void_p(_fp(y)), ctypes.c_void_p(_fp(ab)),
                                  ctypes.c_float(0.85), ctypes.c_float(1e-4), ctypes.c_float(9e-4), ctypes.c_void_p(_fp(ref)),
                                  ctypes.c_void_p(_fp(new)), ctypes.c_void_p(_fp(one)), ctypes.c_void_p(_fp(val)))
    assert rc == 0
    # the reference returns pe summed over channels (= 3 pe), dpa / dpb summed over channels, unit-weight coefficients
    scale = ref.abs().amax(dim=(0, 1))
    err = (new - ref).abs().amax(dim=(0, 1))
    assert (err <= 5e-5 * scale + 1e-7).all(), (err / scale)
    assert torch.allclose(val, new[:, :, 0], rtol=1e-6, atol=1e-7)            # value-only == fused value
    assert (one - new[:, 0]).abs().max().item() <= 1e-5 * scale.max().item()  # N = 1 instantiation == lane 0 of N = 2
