"""Diagnostic (B200 only): per-pixel internals of k_photo_bwd from a COLVO_DEBUG_DUMP build (build/variants/lib_dbg.so)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import coivo_b200
from coivo_b200 import _lib
from coivo_b200.synthetic import make_triplets
B, H, W, N, S = [int(x) for x in sys.argv[1:6]]
DEV = "cuda:0"
d = make_triplets(B, H, W, N=N, S=S, seed=B + H)
depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
pose = d["pose"].to(DEV).requires_grad_(); srcs = d["srcs"].to(DEV).requires_grad_()
lib = _lib.load()
dbg = torch.zeros(B, H, W, 12, device=DEV)
lib.colvo_debug_set_buffer.restype = ctypes.c_int
print("set", lib.colvo_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr())))
loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_masks=True)
loss.backward(); torch.cuda.synchronize()
dbg = dbg.cpu()
torch.set_printoptions(linewidth=250, precision=4, sci_mode=False)
print("ab", ab.cpu()[:, 0, 0])
for b in range(B):
    print("b", b, "a", dbg[b, :, :, 0].unique()[:5], "b", dbg[b, :, :, 1].unique()[:5], "cc.wl1", dbg[b, :, :, 11].unique()[:5])
    m = dbg[b, :, :, 3] == N
    print("   own_sel==N count", int(m.sum()), "sel tensor count", int((sel[b, 0] == N).sum()), "mismatch", int((dbg[b, :, :, 3] != sel[b, 0].cpu().float()).sum()))
    a, bb = ab[b, 0, 0, 0].item(), ab[b, 0, 0, 1].item()
    diff_ref = a * dbg[b, :, :, 7] + bb - dbg[b, :, :, 8]
    print("   diff vs a*xq+b-y max abs", (dbg[b, :, :, 4] - diff_ref).abs().max().item(), " y vs tgt", (dbg[b, :, :, 8] - d["tgt"][b, 0]).abs().max().item())
    sg_ref = torch.sign(dbg[b, :, :, 4]) * dbg[b, :, :, 2]
    print("   sg vs sign(diff)*wl1", (dbg[b, :, :, 5] - sg_ref).abs().max().item(), "wl1 values", dbg[b, :, :, 2].unique()[:4])
