"""Front-end: pose parameterisation and disparity -> depth.
CPU: the oracle's own checks.  GPU: the CUDA kernels against the oracle, forward and backward."""
import math

import pytest
import torch

from oracle import frontend as OF
from oracle import photometric as O


def test_oracle_rotation_is_orthonormal_and_matches_matrix_exp():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(5, 3, generator=g, dtype=torch.float64) * 0.3
    R = OF.rot_from_axisangle(v)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3, dtype=torch.float64).expand_as(R), atol=1e-6)
    zero = torch.zeros(5, dtype=torch.float64)
    Kx = torch.stack([zero, -v[:, 2], v[:, 1], v[:, 2], zero, -v[:, 0], -v[:, 1], v[:, 0], zero], -1).reshape(5, 3, 3)
    assert torch.allclose(R, torch.linalg.matrix_exp(Kx), atol=1e-6)
    assert torch.allclose(OF.rot_from_axisangle(torch.zeros(1, 3)), torch.eye(3).unsqueeze(0), atol=1e-7)


def test_oracle_invert_is_the_inverse_transform():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(4, 3, generator=g, dtype=torch.float64) * 0.2
    t = torch.randn(4, 3, generator=g, dtype=torch.float64)
    T = OF.transformation_from_parameters(a, t, False)
    Ti = OF.transformation_from_parameters(a, t, True)
    assert torch.allclose(T @ Ti, torch.eye(4, dtype=torch.float64).expand_as(T), atol=1e-6)
    assert torch.equal(T[:, 3], torch.tensor([0.0, 0, 0, 1], dtype=torch.float64).expand(4, 4))


def test_oracle_disp_to_depth_range_and_gradcheck():
    d = torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64)
    z = OF.disp_to_depth(d, 0.1, 100.0)
    assert abs(z[0].item() - 100.0) < 1e-9 and abs(z[2].item() - 0.1) < 1e-12
    x = torch.rand(7, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda u: OF.disp_to_depth(u, 0.1, 100.0), (x,))
    a = (torch.randn(2, 2, 3, dtype=torch.float64) * 0.1).requires_grad_()
    t = torch.randn(2, 2, 3, dtype=torch.float64).requires_grad_()
    assert torch.autograd.gradcheck(lambda p, q: OF.poses_from_parameters(p, q, [True, False])[..., :3, :], (a, t))


@pytest.mark.gpu
def test_cuda_frontend_matches_oracle():
    import coivo_b200
    dev = "cuda:0"
    g = torch.Generator().manual_seed(2)
    aa = (torch.randn(3, 2, 3, generator=g) * 0.05)
    aa[0, 0] = 0.0                                   # zero rotation: guarded division
    tr = torch.randn(3, 2, 3, generator=g) * 0.1
    w = torch.randn(3, 2, 4, 4, generator=g)
    for inv in ([True, False], [False, False], [True, True]):
        a_c, t_c = aa.clone().requires_grad_(), tr.clone().requires_grad_()
        T_ref = OF.poses_from_parameters(a_c, t_c, inv)
        (T_ref * w).sum().backward()
        a_g, t_g = aa.to(dev).requires_grad_(), tr.to(dev).requires_grad_()
        T = coivo_b200.poses_from_parameters(a_g, t_g, inv)
        (T * w.to(dev)).sum().backward()
        assert torch.allclose(T.cpu(), T_ref, atol=1e-6)
        assert torch.allclose(a_g.grad.cpu(), a_c.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(t_g.grad.cpu(), t_c.grad, rtol=1e-4, atol=1e-6)
    disp = [torch.rand(2, 1, 16 >> k, 24 >> k, generator=g) for k in range(3)]
    wd = [torch.randn_like(d) for d in disp]
    dc = [d.clone().requires_grad_() for d in disp]
    sum((OF.disp_to_depth(d, 0.1, 100.0) * q).sum() for d, q in zip(dc, wd)).backward()
    dg = [d.to(dev).requires_grad_() for d in disp]
    out = coivo_b200.disp_to_depth(dg, 0.1, 100.0)
    sum((o * q.to(dev)).sum() for o, q in zip(out, wd)).backward()
    for k in range(3):
        assert torch.allclose(out[k].cpu(), OF.disp_to_depth(disp[k], 0.1, 100.0), rtol=1e-6)
        assert torch.allclose(dg[k].grad.cpu(), dc[k].grad, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_cuda_loss_on_raw_network_outputs_matches_oracle_chain():
    import coivo_b200
    from coivo_b200.synthetic import make_triplets
    dev = "cuda:0"
    d = make_triplets(2, 48, 64, seed=31)
    g = torch.Generator().manual_seed(3)
    disp = [torch.rand(2, 1, 48 >> k, 64 >> k, generator=g) * 0.05 + 0.005 for k in range(4)]   # depth ~ 1.8 .. 16
    aa = torch.randn(2, 2, 3, generator=g) * 0.005
    tr = torch.randn(2, 2, 3, generator=g) * 0.02
    K, tgt, srcs = d["K"].to(dev), d["tgt"].to(dev), d["srcs"].to(dev)
    # CUDA chain, explicit so that the intermediate depth / pose are visible
    dg = [x.to(dev).requires_grad_() for x in disp]
    ag, tg = aa.to(dev).requires_grad_(), tr.to(dev).requires_grad_()
    depth_g = coivo_b200.disp_to_depth(dg, 0.1, 100.0)
    pose_g = coivo_b200.poses_from_parameters(ag, tg, (True, False))
    loss, valid, sel, ab = coivo_b200.photometric_loss(depth_g, pose_g, K, tgt, srcs, return_masks=True)
    loss.backward()
    # the one-call wrapper is the same chain
    with torch.no_grad():
        l_raw = coivo_b200.photometric_loss_raw([x.detach() for x in dg], ag.detach(), tg.detach(), K, tgt, srcs, invert=(True, False))
    assert l_raw.item() == loss.item()
    # Oracle: the loss on the SAME depth / pose tensors (the geometry chain is pinned bit for bit, a 1-ulp
    # difference in depth could move a sample across a texel boundary), then the front-end's vector-Jacobian
    # products through the oracle front-end.
    depth_leaf = [x.detach().cpu().requires_grad_() for x in depth_g]
    pose_leaf = pose_g.detach().cpu().requires_grad_()
    with torch.no_grad():
        _, v0, _, _ = O.photometric_loss(depth_leaf, pose_leaf, d["K"], d["tgt"], d["srcs"], return_masks=True)
    l_ref = O.photometric_loss(depth_leaf, pose_leaf, d["K"], d["tgt"], d["srcs"], sel_override=sel.cpu(), ab_override=ab.cpu())
    l_ref.backward()
    assert abs(loss.item() - l_ref.item()) <= 1e-4 * abs(l_ref.item())
    assert torch.equal(valid.cpu(), v0)
    dc = [x.clone().requires_grad_() for x in disp]
    ac, tc = aa.clone().requires_grad_(), tr.clone().requires_grad_()
    torch.autograd.backward([OF.disp_to_depth(x, 0.1, 100.0) for x in dc], [x.grad for x in depth_leaf])
    OF.poses_from_parameters(ac, tc, [True, False]).backward(pose_leaf.grad)
    rel = lambda a, b: (a.cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    for k in range(4):
        assert rel(dg[k].grad, dc[k].grad) < 2e-4
    assert rel(ag.grad, ac.grad) < 2e-4 and rel(tg.grad, tc.grad) < 2e-4
