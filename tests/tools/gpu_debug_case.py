"""Diagnostic: where does grad_depth differ from the oracle for one case?  (B200 only; prints, does not assert)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets
from oracle import photometric as O

DEV = "cuda:0"
B, H, W, N, S = [int(x) for x in sys.argv[1:6]] if len(sys.argv) > 5 else (1, 256, 320, 2, 1)
d = make_triplets(B, H, W, N=N, S=S, seed=B + H)
depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
pose = d["pose"].to(DEV).requires_grad_(); srcs = d["srcs"].to(DEV).requires_grad_()
loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_masks=True)
loss.backward(); torch.cuda.synchronize()
od = [x.clone().requires_grad_() for x in d["depth"]]; op = d["pose"].clone().requires_grad_(); osr = d["srcs"].clone().requires_grad_()
O.photometric_loss(od, op, d["K"], d["tgt"], osr, sel_override=sel.cpu(), ab_override=ab.cpu()).backward()
for k in range(S):
    g, r = depth[k].grad.cpu(), od[k].grad
    e = (g - r).abs()
    print(f"k={k} max|ref| {r.abs().max():.3e} max err {e.max():.3e} rel {e.max() / r.abs().max():.2e}; #err>1e-5*max: {(e > 1e-5 * r.abs().max()).sum().item()}")
    flat = e.flatten().topk(8)
    for v, i in zip(flat.values, flat.indices):
        i = i.item(); bb = i // (g.shape[2] * g.shape[3]); rem = i % (g.shape[2] * g.shape[3]); y = rem // g.shape[3]; x = rem % g.shape[3]
        print(f"   b={bb} y={y} x={x} err {v:.3e} cuda {g[bb,0,y,x]:.6e} ref {r[bb,0,y,x]:.6e} sel {sel[bb,k,max(y-1,0):y+2,max(x-1,0):x+2].flatten().tolist() if k == 0 else ''}")
gs, rs = srcs.grad.cpu(), osr.grad
e = (gs - rs).abs()
print(f"srcs: max|ref| {rs.abs().max():.3e} max err {e.max():.3e}")
flat = e.flatten().topk(5)
for v, i in zip(flat.values, flat.indices):
    idx = torch.unravel_index(i, gs.shape)
    print("   ", [int(t) for t in idx], f"err {v:.3e} cuda {gs[idx]:.6e} ref {rs[idx]:.6e}")
