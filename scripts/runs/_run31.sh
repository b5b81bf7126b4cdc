export COLVO_LIB=${COLVO_LIB_SEL:-}; [ -z "$COLVO_LIB" ] && unset COLVO_LIB
python - <<'PY'
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.getcwd())
import torch, coivo_b200
from coivo_b200.synthetic import make_triplets
dev = "cuda:0"
for (B, H, W, N, S) in [(3, 7, 63, 2, 3), (2, 7, 13, 2, 1), (2, 48, 64, 2, 4), (3, 7, 63, 2, 3)]:
    d = make_triplets(B, H, W, N=N, S=S, seed=1)
    tp, sp = coivo_b200.pack_images(d["tgt"]), coivo_b200.pack_images(d["srcs"])
    try:
        depth = [x.to(dev).requires_grad_() for x in d["depth"]]
        pose = d["pose"].to(dev).requires_grad_()
        l = coivo_b200.photometric_loss(depth, pose, d["K"].to(dev), tp.to(dev), sp.to(dev), alpha=0.5)
        l.backward(); torch.cuda.synchronize()
        print("packed", (B, H, W, N, S), "ok", l.item(), flush=True)
    except Exception as e:
        print("packed", (B, H, W, N, S), "FAILED", str(e)[:160], flush=True)
PY
