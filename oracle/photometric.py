"""CPU oracle for the ColVO photometric-loss hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the upstream repository (HNUicda/CoIVO, mounted at /root/reference)
ships a README and three PNG figures and *no* source, tests, golden vectors or
weights (`/root/reference/README.md:1-31` is the whole text).  There is therefore no
reference output this file could be pinned to.  It is a plain-PyTorch CPU
restatement of the Monodepth2-style view-synthesis loss that BASELINE.json's
north_star names, with every assumption listed in `oracle/ASSUMPTIONS.md`
(A0..A16).  What pins it instead: closed-form known-answer tests, fp64
`gradcheck`, and cross-checks against torch library ops (`tests/test_oracle_*.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / reference
legs may import this module, and only as the checker or the reported CPU baseline.
Nothing under `coivo_b200/` imports it: the product path is CUDA-only.

What the README says about this path (the only upstream evidence):
  * "loss function constraints to couple depth and pose estimation modes ensures
    seamless alignment of geometric projections between consecutive frames"
    (`/root/reference/README.md:7`)  -> rows 0-4, 6-10 of SURVEY.md section 8(a).
  * "LCC accounts for brightness variations by recalibrating the luminosity values
    of adjacent frames" (`/root/reference/README.md:7`) -> row 5 (`lcc_fit`).

Arithmetic contract (what the CUDA kernels must reproduce):
  * The geometry chain (depth upsample, back-projection, SE(3) transform,
    projection, validity mask) is written as *single-rounded elementwise fp32 ops in
    a fixed order* (no bmm / addcmul / lerp), so a kernel using
    __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__frcp_rn in the same order reproduces
    `valid` bit for bit.
  * LCC statistics are accumulated in fp64 and the (a, b) pair is cast to fp32.
  * Everything else (bilinear blend, SSIM, L1, smoothness) is fp32 with no order
    guarantee; parity there is a tolerance (rel 1e-4), not bit equality.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F

# Constants of the path (A7, A8, A12).
ALPHA = 0.85
SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2
EPS_PROJ = 1e-7
EPS_LCC = 1e-6
EPS_MEAN_DISP = 1e-7
Z_MIN = 1e-3
SMOOTH_WEIGHT = 1e-3


def pyramid_shapes(H: int, W: int, S: int):
    """Scale k has h_k = floor(H / 2^k), w_k = floor(W / 2^k)  [A1, A3]."""
    return [(H >> k, W >> k) for k in range(S)]


# --------------------------------------------------------------------------------------
# Row 0: depth upsample (bilinear, align_corners=False semantics), pinned order.
# --------------------------------------------------------------------------------------
def _upsample_axis(n_out: int, n_in: int, dtype):
    """Source index / weight along one axis:  s = max((i + 0.5) * (n_in / n_out) - 0.5, 0).

    `n_in / n_out` is evaluated in double on the host and rounded once to `dtype`
    (the kernel receives the same rounded float as a launch parameter).
    """
    ratio = torch.tensor(float(n_in) / float(n_out), dtype=dtype)
    i = torch.arange(n_out, dtype=dtype)
    s = (i + 0.5) * ratio          # two roundings: add (exact), mul
    s = s - 0.5
    s = torch.clamp_min(s, 0.0)
    i0f = torch.floor(s)
    w1 = s - i0f
    i0 = i0f.to(torch.long).clamp_max(n_in - 1)
    i1 = (i0 + 1).clamp_max(n_in - 1)
    return i0, i1, w1


def upsample_depth(D: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """`D [B,1,h,w]` -> `[B,1,H,W]`.  Identity when (h, w) == (H, W)  [A3].

    Blend order (each op rounded once):  top = (1-wx)*d00 + wx*d01,
    bot = (1-wx)*d10 + wx*d11,  out = (1-wy)*top + wy*bot.
    """
    h, w = D.shape[-2:]
    if (h, w) == (H, W):
        return D
    y0, y1, wy = _upsample_axis(H, h, D.dtype)
    x0, x1, wx = _upsample_axis(W, w, D.dtype)
    wy = wy[:, None]
    wx = wx[None, :]
    d00 = D[:, :, y0[:, None], x0[None, :]]
    d01 = D[:, :, y0[:, None], x1[None, :]]
    d10 = D[:, :, y1[:, None], x0[None, :]]
    d11 = D[:, :, y1[:, None], x1[None, :]]
    omx = 1.0 - wx
    omy = 1.0 - wy
    top = omx * d00 + wx * d01
    bot = omx * d10 + wx * d11
    return omy * top + wy * bot


# --------------------------------------------------------------------------------------
# Rows 1-3: back-project, transform, project, validity.  Pinned order.
# --------------------------------------------------------------------------------------
def reproject(Dhat: torch.Tensor, K: torch.Tensor, T: torch.Tensor):
    """`Dhat [B,1,H,W]`, `K [B,3,3]`, `T [B,4,4]` (target camera -> source camera, A5).

    Returns `u, v [B,H,W]` (source pixel coordinates), `valid [B,H,W]` (bool) and `Zp`.

      rx = (u - cx) / fx ; ry = (v - cy) / fy
      X = rx*D ; Y = ry*D ; Z = D
      X'_i = ((R_i0*X + R_i1*Y) + R_i2*Z) + t_i
      x = fx*X' + cx*Z' ; y = fy*Y' + cy*Z'
      iz = 1 / (Z' + 1e-7) ; u' = x*iz ; v' = y*iz                        [A15]
      valid = (0 <= u' <= W-1) & (0 <= v' <= H-1) & (Z' > z_min)          [A7]
    """
    B, _, H, W = Dhat.shape
    dt = Dhat.dtype
    fx = K[:, 0, 0].reshape(B, 1, 1)
    fy = K[:, 1, 1].reshape(B, 1, 1)
    cx = K[:, 0, 2].reshape(B, 1, 1)
    cy = K[:, 1, 2].reshape(B, 1, 1)
    uu = torch.arange(W, dtype=dt).reshape(1, 1, W)
    vv = torch.arange(H, dtype=dt).reshape(1, H, 1)
    rx = (uu - cx) / fx                      # [B,1,W]
    ry = (vv - cy) / fy                      # [B,H,1]
    D = Dhat[:, 0]
    X = rx * D
    Y = ry * D
    Z = D

    def row(i):
        r0 = T[:, i, 0].reshape(B, 1, 1)
        r1 = T[:, i, 1].reshape(B, 1, 1)
        r2 = T[:, i, 2].reshape(B, 1, 1)
        t = T[:, i, 3].reshape(B, 1, 1)
        acc = r0 * X
        acc = acc + r1 * Y
        acc = acc + r2 * Z
        return acc + t

    Xp, Yp, Zp = row(0), row(1), row(2)
    x = fx * Xp + cx * Zp
    y = fy * Yp + cy * Zp
    iz = torch.ones_like(Zp) / (Zp + EPS_PROJ)
    u = x * iz
    v = y * iz
    valid = (u >= 0) & (u <= W - 1) & (v >= 0) & (v <= H - 1) & (Zp > Z_MIN)
    return u, v, valid, Zp


# --------------------------------------------------------------------------------------
# Row 4: bilinear sampling with border padding on pixel coordinates (A2).
# --------------------------------------------------------------------------------------
def _clamp_coord(c: torch.Tensor, hi: int) -> torch.Tensor:
    """Border clamp with grid_sample's coordinate gradient: zero at and outside the
    border, identity strictly inside.  NaN is mapped to 0."""
    c = torch.where(c == c, c, torch.zeros_like(c))
    inside = (c > 0) & (c < hi)
    cl = c.clamp(0, hi)
    return torch.where(inside, c, cl.detach())


def bilinear_sample(src: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """`src [B,C,H,W]`, `u, v [B,H,W]` -> `[B,C,H,W]`.
    Equals `F.grid_sample(mode="bilinear", padding_mode="border", align_corners=True)`
    on pixel coordinates (KAT-6)."""
    B, C, H, W = src.shape
    ut = _clamp_coord(u, W - 1)
    vt = _clamp_coord(v, H - 1)
    x0f = torch.floor(ut.detach())
    y0f = torch.floor(vt.detach())
    wx = (ut - x0f).unsqueeze(1)
    wy = (vt - y0f).unsqueeze(1)
    x0 = x0f.to(torch.long).clamp(0, W - 1)
    y0 = y0f.to(torch.long).clamp(0, H - 1)
    x1 = (x0 + 1).clamp_max(W - 1)
    y1 = (y0 + 1).clamp_max(H - 1)
    flat = src.reshape(B, C, H * W)

    def tap(yi, xi):
        idx = (yi * W + xi).reshape(B, 1, H * W).expand(B, C, H * W)
        return flat.gather(2, idx).reshape(B, C, H, W)

    i00, i01, i10, i11 = tap(y0, x0), tap(y0, x1), tap(y1, x0), tap(y1, x1)
    top = (1.0 - wx) * i00 + wx * i01
    bot = (1.0 - wx) * i10 + wx * i11
    return (1.0 - wy) * top + wy * bot


# --------------------------------------------------------------------------------------
# Row 5: LCC -- light consistent calibration (README.md:5,7), closed-form affine fit (A6).
# --------------------------------------------------------------------------------------
def lcc_fit(Iw: torch.Tensor, tgt: torch.Tensor, valid: torch.Tensor):
    """Least-squares `a*Iw + b ~ tgt` over valid pixels, jointly over the 3 channels.

    Sums are fp64.  Returns `(a, b)` as `[B]` tensors of `Iw.dtype`.  n == 0 -> (1, 0).
    The mask is a constant (not differentiated); `tgt` gets no gradient (A14).
    """
    dt = Iw.dtype
    m = valid.unsqueeze(1).to(torch.float64)
    x = Iw.to(torch.float64)
    y = tgt.detach().to(torch.float64)
    n = 3.0 * m.sum(dim=(1, 2, 3))
    nz = n.clamp_min(1.0)
    Sx = (x * m).sum(dim=(1, 2, 3))
    Sy = (y * m).sum(dim=(1, 2, 3))
    Sxx = (x * x * m).sum(dim=(1, 2, 3))
    Sxy = (x * y * m).sum(dim=(1, 2, 3))
    mx = Sx / nz
    my = Sy / nz
    var = Sxx / nz - mx * mx
    cov = Sxy / nz - mx * my
    a = cov / (var + EPS_LCC)
    b = my - a * mx
    has = n > 0
    a = torch.where(has, a, torch.ones_like(a))
    b = torch.where(has, b, torch.zeros_like(b))
    return a.to(dt), b.to(dt)


# --------------------------------------------------------------------------------------
# Rows 6-7: SSIM (3x3, reflect pad) + L1 photometric error.
# --------------------------------------------------------------------------------------
def box3_reflect(x: torch.Tensor) -> torch.Tensor:
    """3x3 box mean with reflect padding 1 (KAT-7 cross-checks a manual version)."""
    return F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), 3, 1)


def ssim_term(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """clamp((1 - SSIM(x, y)) / 2, 0, 1) per channel  [A8]."""
    mu_x = box3_reflect(x)
    mu_y = box3_reflect(y)
    sig_x = box3_reflect(x * x) - mu_x * mu_x
    sig_y = box3_reflect(y * y) - mu_y * mu_y
    sig_xy = box3_reflect(x * y) - mu_x * mu_y
    num = (2 * mu_x * mu_y + SSIM_C1) * (2 * sig_xy + SSIM_C2)
    den = (mu_x * mu_x + mu_y * mu_y + SSIM_C1) * (sig_x + sig_y + SSIM_C2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


def photometric_error(x: torch.Tensor, y: torch.Tensor, alpha: float = ALPHA) -> torch.Tensor:
    """`pe = alpha * mean_c(ssim_term) + (1 - alpha) * mean_c |x - y|` -> `[B,H,W]`."""
    l1 = (x - y).abs().mean(dim=1)
    ss = ssim_term(x, y).mean(dim=1)
    return alpha * ss + (1 - alpha) * l1


# --------------------------------------------------------------------------------------
# Row 9: edge-aware smoothness on mean-normalised inverse depth (A11).
# --------------------------------------------------------------------------------------
def smoothness(Dk: torch.Tensor, Ik: torch.Tensor) -> torch.Tensor:
    d = 1.0 / Dk
    mean = d.mean(dim=(2, 3), keepdim=True)
    dn = d / (mean + EPS_MEAN_DISP)
    gx = (dn[:, :, :, :-1] - dn[:, :, :, 1:]).abs()
    gy = (dn[:, :, :-1, :] - dn[:, :, 1:, :]).abs()
    ix = (Ik[:, :, :, :-1] - Ik[:, :, :, 1:]).abs().mean(dim=1, keepdim=True)
    iy = (Ik[:, :, :-1, :] - Ik[:, :, 1:, :]).abs().mean(dim=1, keepdim=True)
    tx = gx * torch.exp(-ix)
    ty = gy * torch.exp(-iy)
    zero = Dk.new_zeros(())
    lx = tx.mean() if tx.numel() else zero
    ly = ty.mean() if ty.numel() else zero
    return lx + ly


def target_pyramid(tgt: torch.Tensor, S: int):
    """`I_t^k = avg_pool2d(I_t, 2^k)` (floor-cropped)  [A11]."""
    return [tgt if k == 0 else F.avg_pool2d(tgt, 2 ** k) for k in range(S)]


# --------------------------------------------------------------------------------------
# geometric consistency (SC-Depth style), re-using the same warp.
# --------------------------------------------------------------------------------------
def geometric_consistency(Zp: torch.Tensor, src_depth_n: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                          valid: torch.Tensor) -> torch.Tensor:
    """The title of the upstream README promises "Geometric ... Consistency" (`README.md:1`) and DCDP
    couples depth and pose through "loss function constraints" (`README.md:7`); the formulation is the
    SC-Depth one [A16]: the depth of the re-projected point, `Z'`, must agree with the source frame's
    own depth map sampled where the point lands,

        D_s' = bilinear_sample(D_s, u', v')            (border padding, as row 4)
        diff = clamp(|Z' - D_s'| / (Z' + D_s'), 0, 1)
        L_geo = mean over b, pixels of (valid ? diff : 0)

    `Zp, u, v, valid [B,H,W]`, `src_depth_n [B,1,H,W]` -> scalar.  Differentiable in the target depth and
    the pose (through Z', u', v') and in the source depth map; the mask is a constant."""
    return geometric_diff(Zp, src_depth_n, u, v, valid).mean()


def geometric_diff(Zp, src_depth_n, u, v, valid):
    """Per-pixel `diff` of `geometric_consistency` (0 where invalid), `[B,H,W]`.  `where(valid, 1 - diff, 0)` is
    SC-Depth's soft occlusion / weight mask (SURVEY.md section 8(f)-2 "also yields a soft occlusion mask")."""
    Ds = bilinear_sample(src_depth_n, u, v)[:, 0]
    diff = torch.clamp((Zp - Ds).abs() / (Zp + Ds), 0, 1)
    return torch.where(valid, diff, torch.zeros_like(diff))


# --------------------------------------------------------------------------------------
# Rows 8, 10: min-reprojection / auto-mask and the total.
# --------------------------------------------------------------------------------------
def _min_first(cands: torch.Tensor):
    """Strict-`<` first-index minimum along dim 1 (A9: no tie-break noise)."""
    m = cands[:, 0]
    sel = torch.zeros_like(m, dtype=torch.long)
    for i in range(1, cands.shape[1]):
        lt = cands[:, i] < m
        m = torch.where(lt, cands[:, i], m)
        sel = torch.where(lt, torch.full_like(sel, i), sel)
    return m, sel


def _validate(depth, pose, K, tgt, srcs):
    if tgt.dim() != 4 or tgt.shape[1] != 3:
        raise ValueError("tgt must be [B,3,H,W]")
    B, _, H, W = tgt.shape
    if srcs.dim() != 5 or srcs.shape[0] != B or srcs.shape[2:] != (3, H, W):
        raise ValueError("srcs must be [B,N,3,H,W]")
    N = srcs.shape[1]
    if pose.shape != (B, N, 4, 4):
        raise ValueError("pose must be [B,N,4,4]")
    if K.shape != (B, 3, 3):
        raise ValueError("K must be [B,3,3]")
    S = len(depth)
    if not 1 <= S <= 4:
        raise ValueError("1 <= len(depth) <= 4")
    if H < 2 or W < 2:
        raise ValueError("H, W >= 2 (reflect padding)")
    for k, (d, (hk, wk)) in enumerate(zip(depth, pyramid_shapes(H, W, S))):
        if d.shape != (B, 1, hk, wk):
            raise ValueError(f"depth[{k}] must be [B,1,{hk},{wk}], got {tuple(d.shape)}")
    return B, N, S, H, W


def photometric_loss(
    depth: Sequence[torch.Tensor],
    pose: torch.Tensor,
    K: torch.Tensor,
    tgt: torch.Tensor,
    srcs: torch.Tensor,
    *,
    alpha: float = ALPHA,
    smooth_weight: float = SMOOTH_WEIGHT,
    lcc: bool = True,
    lcc_detach: bool = False,
    return_masks: bool = False,
    sel_override: Optional[torch.Tensor] = None,
    ab_override: Optional[torch.Tensor] = None,
    src_depth: Optional[torch.Tensor] = None,
    geo_weight: float = 0.0,
    return_occlusion: bool = False,
):
    """The oracle for SURVEY.md section 8(a) rows 0-10 (row 11 = autograd of this).

    depth: S tensors `[B,1,h_k,w_k]`;  pose `[B,N,4,4]`;  K `[B,3,3]`;  tgt `[B,3,H,W]`;
    srcs `[B,N,3,H,W]`.  Returns the scalar loss, or
    `(loss, valid u8 [B,N,S,H,W], sel u8 [B,S,H,W], ab [B,N,S,2])`.

    `sel_override` (`[B,S,H,W]` integer) replaces the arg-min decision and
    `ab_override` (`[B,N,S,2]`) replaces the *values* of (a, b) while keeping their
    dependence on the warped image (straight-through), so that gradients can be
    compared against a kernel that made a different near-tie / last-ulp choice
    (SURVEY.md section 7.4 H2).

    `src_depth [B,N,1,H,W]` with `geo_weight > 0` adds the geometric-consistency term of
    SURVEY.md section 8(f)-2 (see `geometric_consistency`), per scale, with the same 1/S;
    `return_occlusion=True` then appends `occ [B,N,S,H,W]`, the soft occlusion mask `valid ? 1 - diff : 0`
    (a detached by-product, see `geometric_diff`) to the returned tuple.
    """
    B, N, S, H, W = _validate(depth, pose, K, tgt, srcs)
    tgt_c = tgt.detach()
    geo_on = src_depth is not None and geo_weight != 0.0
    if geo_on and tuple(src_depth.shape) != (B, N, 1, H, W):
        raise ValueError("src_depth must be [B,N,1,H,W]")
    # Identity candidates: raw sources, no LCC, detached (A10).
    ident = [photometric_error(srcs[:, n].detach(), tgt_c, alpha) for n in range(N)]
    pyr = target_pyramid(tgt_c, S)
    total = tgt.new_zeros(())
    if return_occlusion and not geo_on:
        raise ValueError("return_occlusion needs src_depth and geo_weight != 0")
    valids, sels, abs_, occs = [], [], [], []
    for k in range(S):
        Dhat = upsample_depth(depth[k], H, W)
        cands = list(ident)
        v_k, ab_k = [], []
        l_geo = tgt.new_zeros(())
        for n in range(N):
            u, v, valid, Zp = reproject(Dhat, K, pose[:, n])
            Iw = bilinear_sample(srcs[:, n], u, v)
            if geo_on:
                diff = geometric_diff(Zp, src_depth[:, n], u, v, valid)
                l_geo = l_geo + diff.mean() / N
                occs.append(torch.where(valid, 1.0 - diff.detach(), torch.zeros_like(diff)))
            if lcc:
                a, b = lcc_fit(Iw.detach() if lcc_detach else Iw, tgt_c, valid)
                if ab_override is not None:
                    a = a + (ab_override[:, n, k, 0].to(a.dtype) - a).detach()
                    b = b + (ab_override[:, n, k, 1].to(b.dtype) - b).detach()
            else:
                a = Iw.new_ones(B)
                b = Iw.new_zeros(B)
            Ic = a.reshape(B, 1, 1, 1) * Iw + b.reshape(B, 1, 1, 1)
            cands.append(photometric_error(Ic, tgt_c, alpha))
            v_k.append(valid)
            ab_k.append(torch.stack([a, b], dim=-1))
        cands = torch.stack(cands, dim=1)                     # [B, 2N, H, W]
        if sel_override is not None:
            sel = sel_override[:, k].to(torch.long)
            m = cands.gather(1, sel.unsqueeze(1)).squeeze(1)
        else:
            m, sel = _min_first(cands)
        l_photo = m.mean()
        l_sm = smoothness(depth[k], pyr[k])
        total = total + l_photo + (smooth_weight / (2 ** k)) * l_sm + geo_weight * l_geo
        valids.append(torch.stack(v_k, dim=1))                # [B,N,H,W]
        sels.append(sel)
        abs_.append(torch.stack(ab_k, dim=1))                 # [B,N,2]
    loss = total / S
    if not (return_masks or return_occlusion):
        return loss
    valid_u8 = torch.stack(valids, dim=2).to(torch.uint8)     # [B,N,S,H,W]
    sel_u8 = torch.stack(sels, dim=1).to(torch.uint8)         # [B,S,H,W]
    ab = torch.stack(abs_, dim=2).detach()                    # [B,N,S,2]
    if return_occlusion:                                      # occs is ordered (k, n) -> [B,N,S,H,W]
        occ = torch.stack(occs, dim=1).reshape(B, S, N, H, W).transpose(1, 2).contiguous()
        return loss, valid_u8, sel_u8, ab, occ
    return loss, valid_u8, sel_u8, ab


def candidates(depth, pose, K, tgt, srcs, *, alpha=ALPHA, lcc=True, ab_override=None):
    """All 2N candidates of min_reprojection_automask per scale, `[B,S,2N,H,W]`, in the dtype of the inputs (pass
    float64 tensors for the fp64 adjudication of arg-min near-ties).  `ab_override`
    `[B,N,S,2]` replaces the calibration values."""
    B, N, S, H, W = _validate(depth, pose, K, tgt, srcs)
    with torch.no_grad():
        ident = [photometric_error(srcs[:, n], tgt, alpha) for n in range(N)]
        out = []
        for k in range(S):
            Dhat = upsample_depth(depth[k], H, W)
            cands = list(ident)
            for n in range(N):
                u, v, valid, _ = reproject(Dhat, K, pose[:, n])
                Iw = bilinear_sample(srcs[:, n], u, v)
                if lcc:
                    if ab_override is not None:
                        a, b = ab_override[:, n, k, 0].to(Iw.dtype), ab_override[:, n, k, 1].to(Iw.dtype)
                    else:
                        a, b = lcc_fit(Iw, tgt, valid)
                    Iw = a.reshape(B, 1, 1, 1) * Iw + b.reshape(B, 1, 1, 1)
                cands.append(photometric_error(Iw, tgt, alpha))
            out.append(torch.stack(cands, 1))
        return torch.stack(out, dim=1)


def candidate_gap(depth, pose, K, tgt, srcs, *, alpha=ALPHA, lcc=True):
    """Gap between the two smallest candidates per pixel `[B,S,H,W]` -- the near-tie
    detector of the `sel` parity protocol."""
    two = candidates(depth, pose, K, tgt, srcs, alpha=alpha, lcc=lcc).topk(2, dim=2, largest=False).values
    return two[:, :, 1] - two[:, :, 0]


def adjudicate_sel(depth, pose, K, tgt, srcs, sel, ab, *, alpha=ALPHA, lcc=True):
    """fp64 adjudication of an arg-min decision `sel [B,S,H,W]` made by an fp32 evaluation (the kernel's, or this
    oracle's own): evaluates every candidate in float64 with the given calibration `ab` and returns
    `(excess [B,S,H,W] float64, sel64)` where `excess = c64[sel] - min c64 >= 0` is how far the chosen candidate is
    from the true minimum.  A legitimate near-tie flip has an excess inside the fp32 evaluation noise of pe."""
    d64 = [x.detach().double() for x in depth]
    c = candidates(d64, pose.detach().double(), K.double(), tgt.double(), srcs.detach().double(), alpha=alpha, lcc=lcc,
                   ab_override=ab.double())
    m, sel64 = c.min(dim=2)
    chosen = c.gather(2, sel.to(torch.long).unsqueeze(2)).squeeze(2)
    return chosen - m, sel64.to(torch.uint8)


def l1_kink_count(depth, pose, K, tgt, srcs, sel, ab, *, tol: float = 1e-6) -> int:
    """Number of samples (b, k, pixel, channel) of the WINNING re-projection candidate whose
    calibrated residual `a*I_w + b - I_t` lies within `tol` of the kink of |.|.

    torch's `abs` has sub-gradient sign(0) = 0 there; a second fp32 evaluation of the same
    residual (the CUDA kernel's, with fused multiply-adds) can land on the other side of 0 or
    exactly on it, which changes dL/dI_w of that one sample by up to (1 - alpha)/3 * a / (S B HW)
    -- the L1 analogue of the arg-min near-tie of SURVEY.md section 7.4 H2.  Parity tests allow that
    many isolated outliers (bounded in size) and are strict when the count is 0.
    `sel` / `ab` are the decisions the gradients were computed with."""
    B, N, S, H, W = _validate(depth, pose, K, tgt, srcs)
    count = 0
    with torch.no_grad():
        for k in range(S):
            Dhat = upsample_depth(depth[k], H, W)
            for n in range(N):
                u, v, _, _ = reproject(Dhat, K, pose[:, n])
                Iw = bilinear_sample(srcs[:, n], u, v)
                a = ab[:, n, k, 0].to(Iw.dtype).reshape(B, 1, 1, 1)
                b = ab[:, n, k, 1].to(Iw.dtype).reshape(B, 1, 1, 1)
                near = ((a * Iw + b - tgt).abs() < tol) & (sel[:, k].to(torch.long) == N + n).unsqueeze(1)
                count += int(near.sum().item())
    return count


# --------------------------------------------------------------------------------------
# SURVEY.md section 8(f)-1 / BASELINE config 5: inference-time warp + LCC consistency.
# --------------------------------------------------------------------------------------
def consistency(depth_seq, pose_seq, K, frames, *, alpha: float = ALPHA, lcc: bool = True):
    """For every consecutive pair (t, t+1) of a sequence: warp frame t+1 into frame t
    with depth_seq[t] and pose_seq[t] (= T_{t -> t+1}), LCC-calibrate, and report
    `[mean pe over valid pixels, a, b, valid fraction]` -> `[F-1, 4]`.

    depth_seq `[F,1,H,W]` (only the first F-1 are used), pose_seq `[F-1,4,4]`,
    K `[3,3]` or `[F-1,3,3]`, frames `[F,3,H,W]`.  (README.md:29: depth maps are
    stitched along the trajectory; this is the check that gates that stitching.)
    """
    Fr, _, H, W = frames.shape
    P = Fr - 1
    if K.dim() == 2:
        K = K.unsqueeze(0).expand(P, 3, 3)
    with torch.no_grad():
        tgt = frames[:-1]
        src = frames[1:]
        u, v, valid, _ = reproject(depth_seq[:P], K, pose_seq)
        Iw = bilinear_sample(src, u, v)
        if lcc:
            a, b = lcc_fit(Iw, tgt, valid)
        else:
            a, b = Iw.new_ones(P), Iw.new_zeros(P)
        Ic = a.reshape(P, 1, 1, 1) * Iw + b.reshape(P, 1, 1, 1)
        pe = photometric_error(Ic, tgt, alpha)
        vf = valid.to(torch.float64)
        cnt = vf.sum(dim=(1, 2))
        pe_mean = (pe.to(torch.float64) * vf).sum(dim=(1, 2)) / cnt.clamp_min(1.0)
        frac = cnt / float(H * W)
        return torch.stack([pe_mean.to(frames.dtype), a, b, frac.to(frames.dtype)], dim=1)
