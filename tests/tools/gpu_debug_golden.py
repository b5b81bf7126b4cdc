"""Diagnostic (B200 only; prints, does not assert): the CUDA path on one golden fixture -- where do the gradients
differ from the golden values and from the oracle run with the kernel's own sel / (a, b)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coivo_b200
import test_golden as G
from oracle import photometric as O

DEV = "cuda:0"
pat = sys.argv[1] if len(sys.argv) > 1 else "b2_24x32"
path = [p for p in G.FILES if pat in p][0]
name, t, S = G.load(path)
kw = G.KW.get(name, {})
depth = [t[f"depth{k}"].to(DEV).requires_grad_() for k in range(S)]
pose = t["pose"].to(DEV).requires_grad_()
srcs = t["srcs"].to(DEV).requires_grad_()
loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, t["K"].to(DEV), t["tgt"].to(DEV), srcs, return_masks=True, **kw)
loss.backward()
torch.cuda.synchronize()
print(name, "loss", loss.item(), t["loss"].item(), "sel mism", int((sel.cpu() != t["sel"]).sum()), "ab err", (ab.cpu() - t["ab"]).abs().max().item())
od = [t[f"depth{k}"].clone().requires_grad_() for k in range(S)]
op = t["pose"].clone().requires_grad_()
osr = t["srcs"].clone().requires_grad_()
O.photometric_loss(od, op, t["K"], t["tgt"], osr, sel_override=sel.cpu(), ab_override=ab.cpu(), **kw).backward()
for k in range(S):
    g = depth[k].grad.cpu()
    for what, r in (("golden", t[f"grad_depth{k}"]), ("oracle(sel,ab)", od[k].grad)):
        e = (g - r).abs()
        print(f"k={k} vs {what}: max|ref| {r.abs().max():.3e} max err {e.max():.3e} rel {e.max() / r.abs().max():.2e} #>1e-5: {(e > 1e-5 * r.abs().max()).sum().item()}")
    e = (g - od[k].grad).abs()
    top = e.flatten().topk(min(6, e.numel()))
    for v, i in zip(top.values, top.indices):
        idx = [int(x) for x in torch.unravel_index(i, g.shape)]
        print(f"     {idx} err {v:.3e} cuda {g.flatten()[i]:.6e} ref {od[k].grad.flatten()[i]:.6e}")
print("pose rel", G.relinf(pose.grad.cpu(), op.grad), "srcs rel", G.relinf(srcs.grad.cpu(), osr.grad))
e = (srcs.grad.cpu() - osr.grad).abs()
top = e.flatten().topk(6)
for v, i in zip(top.values, top.indices):
    print("   srcs", [int(x) for x in torch.unravel_index(i, e.shape)], f"err {v:.3e} ref {osr.grad.flatten()[i]:.4e}")
