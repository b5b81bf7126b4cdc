export COLVO_DEBUG=1
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch, coivo_b200
from coivo_b200.synthetic import make_triplets
dev = "cuda:0"
def run(tag, B, H, W, N, S, packed, **kw):
    d = make_triplets(B, H, W, N=N, S=S, seed=1)
    tgt, srcs = d["tgt"], d["srcs"]
    if packed:
        tgt, srcs = coivo_b200.pack_images(tgt), coivo_b200.pack_images(srcs)
    try:
        depth = [x.to(dev).requires_grad_() for x in d["depth"]]
        pose = d["pose"].to(dev).requires_grad_()
        l = coivo_b200.photometric_loss(depth, pose, d["K"].to(dev), tgt.to(dev), srcs.to(dev), **kw)
        l.backward(); torch.cuda.synchronize()
        print(tag, (B, H, W, N, S), "ok", flush=True)
    except Exception as e:
        print(tag, (B, H, W, N, S), "FAILED", str(e)[:100], flush=True)
run("packed a=.5", 2, 48, 64, 2, 4, True, alpha=0.5)
run("packed default", 2, 48, 64, 2, 4, True)
run("planar nosrcgrad", 2, 48, 64, 2, 4, False)
run("planar nosrcgrad a=.5", 2, 48, 64, 2, 4, False, alpha=0.5)
PY
