"""Small-shape exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets, make_sequence

DEV = "cuda:0"
for (B, H, W, N, S) in [(1, 37, 53, 2, 4), (2, 16, 40, 1, 2), (1, 9, 33, 2, 3)]:
    d = make_triplets(B, H, W, N=N, S=S, seed=1)
    depth = [x.to(DEV).requires_grad_() for x in d["depth"]]
    pose = d["pose"].to(DEV).requires_grad_()
    srcs = d["srcs"].to(DEV).requires_grad_()
    loss, valid, sel, ab = coivo_b200.photometric_loss(depth, pose, d["K"].to(DEV), d["tgt"].to(DEV), srcs, return_masks=True)
    loss.backward()
    torch.cuda.synchronize()
    print(B, H, W, N, S, float(loss))
s = make_sequence(4, 19, 35, seed=2)
print(coivo_b200.consistency(s["depth"].to(DEV), s["pose"].to(DEV), s["K"].to(DEV), s["frames"].to(DEV)).cpu()[0])
d = make_triplets(2, 16, 24, seed=3)
st = coivo_b200.HostStepper(2, 2, 4, 16, 24, device=DEV, chunks=2)
pin = lambda t: t.pin_memory()
st.step([pin(x) for x in d["depth"]], pin(d["pose"]), pin(d["K"]), pin(d["tgt"]), pin(d["srcs"]))
print(float(st.finish()))
print("sanitize_small ok")
