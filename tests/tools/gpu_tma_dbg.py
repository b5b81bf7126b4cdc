"""Diagnostic: forward + backward, blocking launches, over a list of shapes (B H W N S on the command line, or defaults)."""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets
shapes = [(3, 7, 63, 2, 3), (1, 7, 63, 2, 3), (3, 8, 63, 2, 3), (3, 7, 64, 2, 3), (3, 7, 63, 1, 3), (3, 7, 63, 2, 1), (1, 5, 63, 2, 1), (1, 6, 63, 2, 1), (1, 4, 8, 1, 1), (1, 2, 2, 1, 1)]
if len(sys.argv) > 5:
    shapes = [tuple(int(x) for x in sys.argv[1:6])]
dev = "cuda:0"
for (B, H, W, N, S) in shapes:
    d = make_triplets(B, H, W, N=N, S=S, seed=1)
    try:
        depth = [x.to(dev).requires_grad_() for x in d["depth"]]
        pose = d["pose"].to(dev).requires_grad_()
        srcs = d["srcs"].to(dev).requires_grad_()
        l = coivo_b200.photometric_loss(depth, pose, d["K"].to(dev), d["tgt"].to(dev), srcs)
        l.backward()
        torch.cuda.synchronize()
        print((B, H, W, N, S), "ok", l.item(), flush=True)
    except Exception as e:
        print((B, H, W, N, S), "FAILED", str(e)[:120], flush=True)
