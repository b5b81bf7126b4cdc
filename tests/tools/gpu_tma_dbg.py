"""Diagnostic: forward only, blocking launches, N = 1 / 2, a few sizes."""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets
from oracle import photometric as O
for (B, H, W, N, S) in [(1, 24, 64, 1, 1), (1, 24, 64, 2, 1), (1, 256, 320, 2, 1), (2, 64, 96, 2, 4)]:
    d = make_triplets(B, H, W, N=N, S=S, seed=1)
    dev = "cuda:0"
    try:
        with torch.no_grad():
            l = coivo_b200.photometric_loss([x.to(dev) for x in d["depth"]], d["pose"].to(dev), d["K"].to(dev), d["tgt"].to(dev), d["srcs"].to(dev))
            torch.cuda.synchronize()
            l0 = O.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"])
        print((B, H, W, N, S), "ok", l.item(), l0.item(), flush=True)
    except Exception as e:
        print((B, H, W, N, S), "FAILED", str(e)[:200], flush=True)
        break
