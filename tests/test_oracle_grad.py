"""fp64 gradcheck of the oracle at tiny sizes."""
import torch

from oracle import photometric as O
from coivo_b200.synthetic import make_triplets


def _inputs(B, H, W, N, S, seed):
    d = make_triplets(B, H, W, N=N, S=S, seed=seed)
    f = lambda t: t.to(torch.float64)
    return [f(x) for x in d["depth"]], f(d["pose"]), f(d["K"]), f(d["tgt"]), f(d["srcs"])


def test_gradcheck_full_loss_fp64():
    depth, pose, K, tgt, srcs = _inputs(1, 12, 16, 2, 2, seed=0)
    with torch.no_grad():
        _, _, sel, _ = O.photometric_loss(depth, pose, K, tgt, srcs, return_masks=True)
    # identity candidates are constants by design (A10): finite differences w.r.t. srcs would see
    # them, autograd must not -> route every pixel to a re-projection candidate for this check
    sel = torch.where(sel < 2, sel + 2, sel)
    depth = [x.requires_grad_() for x in depth]
    pose.requires_grad_()
    srcs.requires_grad_()

    def fn(d0, d1, p, s):
        # freeze the arg-min (piecewise-constant decision) so finite differences are well defined
        return O.photometric_loss([d0, d1], p, K, tgt, s, sel_override=sel, smooth_weight=0.1)

    assert torch.autograd.gradcheck(fn, (depth[0], depth[1], pose, srcs), eps=1e-6, atol=1e-6, rtol=1e-4, nondet_tol=0.0)


def test_gradcheck_lcc_detach_and_no_lcc():
    depth, pose, K, tgt, srcs = _inputs(1, 10, 12, 1, 1, seed=1)
    with torch.no_grad():
        _, _, sel, _ = O.photometric_loss(depth, pose, K, tgt, srcs, lcc=False, return_masks=True)
    depth[0].requires_grad_()
    pose.requires_grad_()
    fn = lambda d0, p: O.photometric_loss([d0], p, K, tgt, srcs, lcc=False, sel_override=sel)
    assert torch.autograd.gradcheck(fn, (depth[0], pose), eps=1e-6, atol=1e-6, rtol=1e-4)


def test_no_grad_to_K_or_tgt():
    depth, pose, K, tgt, srcs = _inputs(1, 10, 12, 2, 2, seed=2)
    K.requires_grad_()
    tgt.requires_grad_()
    depth[0].requires_grad_()
    loss = O.photometric_loss(depth, pose, K, tgt, srcs, smooth_weight=0.0)
    g = torch.autograd.grad(loss, [tgt], allow_unused=True)
    assert g[0] is None                     # A14: tgt is detached everywhere


def test_gradcheck_geometric_consistency_term():
    depth, pose, K, tgt, srcs = _inputs(1, 12, 16, 2, 2, seed=3)
    g = torch.Generator().manual_seed(0)
    sd = (1.0 + 0.5 * torch.rand(1, 2, 1, 12, 16, generator=g, dtype=torch.float64)).requires_grad_()
    with torch.no_grad():
        _, _, sel, _ = O.photometric_loss(depth, pose, K, tgt, srcs, return_masks=True)
    depth = [x.requires_grad_() for x in depth]
    pose.requires_grad_()
    fn = lambda d0, d1, p, s_: O.photometric_loss([d0, d1], p, K, tgt, srcs, sel_override=sel, src_depth=s_, geo_weight=0.5)
    assert torch.autograd.gradcheck(fn, (depth[0], depth[1], pose, sd), eps=1e-6, atol=1e-6, rtol=1e-4)


def test_geometric_consistency_is_zero_for_a_consistent_scene():
    # fronto-parallel plane seen by a camera translated along x: Z' = D everywhere, so a source depth map equal
    # to the same constant is perfectly consistent; a different constant gives the closed-form ratio
    B, H, W = 1, 12, 20
    from coivo_b200.synthetic import make_intrinsics
    K = make_intrinsics(B, H, W)
    T = torch.eye(4).reshape(1, 4, 4).clone()
    T[0, 0, 3] = 0.02
    D = torch.full((B, 1, H, W), 2.0)
    u, v, valid, Zp = O.reproject(D, K, T)
    assert O.geometric_consistency(Zp, torch.full((B, 1, H, W), 2.0), u, v, valid).item() < 1e-7
    got = O.geometric_consistency(Zp, torch.full((B, 1, H, W), 3.0), u, v, valid).item()
    assert abs(got - valid.float().mean().item() * (1.0 / 5.0)) < 1e-6
