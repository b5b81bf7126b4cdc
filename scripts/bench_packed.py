"""Device-timed fwd+bwd of config 2 with packed bf16 image storage vs planar fp32
(srcs not requiring grad in both, since packed images carry no gradient)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets

DEV = torch.device("cuda:0")
B, H, W = 12, 256, 320


def run(packed, sets=6, steps=100, warm=10):
    bs = []
    for r in range(sets):
        d = make_triplets(B, H, W, seed=70 + r)
        tgt, srcs = d["tgt"], d["srcs"]
        if packed:
            tgt, srcs = coivo_b200.pack_images(tgt), coivo_b200.pack_images(srcs)
        bs.append(([x.to(DEV).requires_grad_() for x in d["depth"]], d["pose"].to(DEV).requires_grad_(), d["K"].to(DEV),
                   tgt.to(DEV), srcs.to(DEV)))
    graphs = [coivo_b200.GraphedStep(*b) for b in bs]
    for i in range(warm):
        graphs[i % sets].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % sets].replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    hw = H * W
    pyr = sum((H >> k) * (W >> k) for k in range(4))
    img_bytes = (8 if packed else 12) * hw * 3              # target + two sources
    alg = B * (2 * (img_bytes + 4 * pyr) + 4 * pyr)         # inputs read twice, depth gradients written once
    print(json.dumps({"storage": "bf16x4 packed" if packed else "fp32 planar (no grad_srcs)", "ms_per_step": ms,
                      "triplets_per_s": B / ms * 1e3, "algorithmic_MB_per_step": alg / 1e6,
                      "alg_GBps": alg / (ms * 1e-3) / 1e9}), flush=True)


if __name__ == "__main__":
    run(False)
    run(True)
