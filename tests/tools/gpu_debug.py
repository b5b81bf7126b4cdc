"""Diagnostic run on a B200: prints parity errors (does not assert) and a rough timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets
from oracle import photometric as O

DEV = "cuda:0"

def relinf(a, b):
    return (a.cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-30)

def one(B, H, W, N, S, seed=0, **kw):
    d = make_triplets(B, H, W, N=N, S=S, seed=seed)
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_(); srcs = d["srcs"].to(DEV).requires_grad_()
    loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_masks=True, **kw)
    loss.backward(); torch.cuda.synchronize()
    with torch.no_grad():
        l0, v0, s0, ab0 = O.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], return_masks=True, **kw)
    od = [x.clone().requires_grad_() for x in d["depth"]]; op = d["pose"].clone().requires_grad_(); osr = d["srcs"].clone().requires_grad_()
    O.photometric_loss(od, op, d["K"], d["tgt"], osr, sel_override=sel.cpu(), ab_override=ab.cpu(), **kw).backward()
    print(f"[{B}x{H}x{W} N={N} S={S} {kw}] loss {loss.item():.7f} vs {l0.item():.7f} rel {abs(loss.item()-l0.item())/abs(l0.item()):.2e} "
          f"valid_mism {(valid.cpu()!=v0).sum().item()} sel_mism {(sel.cpu()!=s0).sum().item()} ab_err {(ab.cpu()-ab0).abs().max().item():.2e}")
    print("   gdepth", ["%.2e" % relinf(depth[k].grad, od[k].grad) for k in range(S)], "gpose %.2e" % relinf(pose.grad, op.grad), "gsrcs %.2e" % relinf(srcs.grad, osr.grad), flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).L2_cache_size)
    one(1, 16, 24, 1, 1, smooth_weight=0.0)
    one(1, 16, 24, 2, 2, smooth_weight=0.0)
    one(2, 64, 96, 2, 4, smooth_weight=0.0)
    one(2, 64, 96, 2, 4)
    one(1, 37, 53, 2, 4)
    one(1, 256, 320, 2, 4)
    one(2, 48, 64, 2, 4, lcc=False)
    one(2, 48, 64, 2, 4, lcc_detach=True)
    # rough timing, config 2
    d = make_triplets(12, 256, 320, seed=0)
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_(); srcs = d["srcs"].to(DEV).requires_grad_()
    K = d["K"].to(DEV); tgt = d["tgt"].to(DEV)
    for it in range(3):
        coivo_b200.photometric_loss(depth, pose, K, tgt, srcs).backward()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(20):
        coivo_b200.photometric_loss(depth, pose, K, tgt, srcs).backward()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"config2 fwd+bwd {ms:.3f} ms/step -> {12/ms*1e3:.0f} triplets/s")
