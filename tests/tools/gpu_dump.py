"""Diagnostic (B200 only): run one synthetic case through the CUDA path and dump outputs, gradients and the raw
forward->backward `saved` buffer to a .pt file, so that two builds of the library (COLVO_LIB=...) can be compared
field by field with gpu_dump_cmp.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets

B, H, W, N, S = [int(x) for x in sys.argv[1:6]]
out = sys.argv[6]
kw = {}
for a in sys.argv[7:]:
    k, v = a.split("=")
    kw[k] = (v == "True") if v in ("True", "False") else float(v)
DEV = "cuda:0"
d = make_triplets(B, H, W, N=N, S=S, seed=B + H)
depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
pose = d["pose"].to(DEV).requires_grad_(); srcs = d["srcs"].to(DEV).requires_grad_()
loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_masks=True, **kw)
saved = loss.grad_fn.saved_tensors[5].clone()
loss.backward(); torch.cuda.synchronize()
HW = H * W
BNS, BS = B * N * S, B * S
nd = BNS * 8 + BS * 2
fl = saved.view(torch.float32)[2 * nd:]
nf = sum(B * (H >> k) * (W >> k) for k in range(S))
nf = (nf + 3) // 4 * 4
if nd & 1:
    nf += 2
coef = fl[nf:nf + BS * 12 * HW].view(B, S, 3, H, W, 4).cpu()
geo = fl[nf + BS * 12 * HW: nf + BS * 12 * HW + BNS * 4 * HW].cpu()
torch.save({"loss": loss.item(), "valid": valid.cpu(), "sel": sel.cpu(), "ab": ab.cpu(), "gd": [x.grad.cpu() for x in depth],
            "gT": pose.grad.cpu(), "gs": srcs.grad.cpu(), "frame": saved[:BNS * 8].view(BNS, 8).cpu(),
            "scale": saved[BNS * 8: nd].view(BS, 2).cpu(), "coef": coef, "geo": geo, "dims": (B, H, W, N, S)}, out)
print("dumped", out, loss.item())
