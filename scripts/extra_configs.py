"""Throughput of the other BASELINE configs on one B200 (device-timed, CUDA events):
   config 4: 4 x 1080x1350, N=2, S=4, fwd+bwd   (bandwidth stress, unaligned rows)
   config 5: 2000-frame consistency sweep, 256x320, forward only
   strong scaling shapes of config 3: per-GPU batches 24/12/6/3 (what 1/2/4/8 GPUs see at global batch 24)
Prints one JSON line per measurement."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets, make_sequence

DEV = torch.device("cuda:0")
PEAK = 6452.2


def alg_bytes(h, w, n=2, s=4):
    hw = h * w
    pyr = sum((h >> k) * (w >> k) for k in range(s))
    return 2 * 4 * (3 * hw + 3 * n * hw + pyr) + 4 * (3 * n * hw + pyr)


def time_steps(fn, warm=5, steps=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def train_config(B, H, W, name, sets=3):
    batches = []
    for r in range(sets):
        d = make_triplets(B, H, W, seed=50 + r)
        batches.append(([x.to(DEV).requires_grad_() for x in d["depth"]], d["pose"].to(DEV).requires_grad_(), d["K"].to(DEV),
                        d["tgt"].to(DEV), d["srcs"].to(DEV).requires_grad_()))
    graphs = [coivo_b200.GraphedStep(*b) for b in batches]
    i = [0]

    def eager():
        b = batches[i[0] % sets]; i[0] += 1
        for t in b[0] + [b[1], b[4]]:
            t.grad = None
        coivo_b200.photometric_loss(*b).backward()

    def graph():
        graphs[i[0] % sets].replay(); i[0] += 1

    ms_e, ms_g = time_steps(eager), time_steps(graph)
    ms = min(ms_e, ms_g)
    gbs = B * alg_bytes(H, W) / (ms * 1e-3) / 1e9
    print(json.dumps({"config": name, "B": B, "H": H, "W": W, "eager_ms": ms_e, "graph_ms": ms_g, "triplets_per_s": B / ms * 1e3,
                      "alg_GBps": gbs, "hbm_frac": gbs / PEAK}), flush=True)


def sweep_config(F, H, W):
    s = make_sequence(F, H, W, seed=7)
    args = (s["depth"].to(DEV), s["pose"].to(DEV), s["K"].to(DEV), s["frames"].to(DEV))
    ms = time_steps(lambda: coivo_b200.consistency(*args), warm=3, steps=10)
    gbs = (F - 1) * 4 * H * W * 7 / (ms * 1e-3) / 1e9
    print(json.dumps({"config": "5: consistency sweep", "frames": F, "H": H, "W": W, "ms": ms, "pairs_per_s": (F - 1) / ms * 1e3,
                      "alg_GBps": gbs, "hbm_frac": gbs / PEAK}), flush=True)


if __name__ == "__main__":
    train_config(4, 1080, 1350, "4: high-res C3VD-shaped, batch 4")
    sweep_config(2000, 256, 320)
    for B in (24, 12, 6, 3, 1):
        train_config(B, 256, 320, f"3-strong: per-GPU batch {B} (256x320)")
