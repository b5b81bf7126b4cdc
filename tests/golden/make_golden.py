"""Generate the committed golden fixtures from the oracle (run from the repo root):

    python tests/golden/make_golden.py

The upstream repository has no golden vectors (no code at all), so these freeze the oracle's own
outputs on small seeded inputs: any later edit to oracle/photometric.py that changes a number
shows up as a diff here, and the CUDA path is checked against the same files on the GPU box
(which has no /root/reference and needs none).  Inputs are stored too, so the fixtures do not
depend on the synthetic generator staying bit-stable.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from coivo_b200.synthetic import make_sequence, make_triplets  # noqa: E402
from oracle import photometric as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "b1_16x24_n2_s2": dict(B=1, H=16, W=24, N=2, S=2, seed=101, kw={}),
    "b2_24x32_n2_s4": dict(B=2, H=24, W=32, N=2, S=4, seed=102, kw={}),
    "b1_19x27_n1_s3": dict(B=1, H=19, W=27, N=1, S=3, seed=103, kw={}),
    "b1_16x24_nolcc": dict(B=1, H=16, W=24, N=2, S=2, seed=104, kw=dict(lcc=False)),
    "b1_16x24_detach": dict(B=1, H=16, W=24, N=2, S=2, seed=105, kw=dict(lcc_detach=True, smooth_weight=0.05)),
}


def run_case(c):
    torch.set_num_threads(1)
    d = make_triplets(c["B"], c["H"], c["W"], N=c["N"], S=c["S"], seed=c["seed"])
    depth = [x.clone().requires_grad_() for x in d["depth"]]
    pose = d["pose"].clone().requires_grad_()
    srcs = d["srcs"].clone().requires_grad_()
    loss, valid, sel, ab = O.photometric_loss(depth, pose, d["K"], d["tgt"], srcs, return_masks=True, **c["kw"])
    loss.backward()
    with torch.no_grad():
        gap = O.candidate_gap(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], lcc=c["kw"].get("lcc", True))
    out = {"tgt": d["tgt"], "srcs": d["srcs"], "K": d["K"], "pose": d["pose"], "loss": loss.detach(), "valid": valid,
           "sel": sel, "ab": ab, "gap": gap, "grad_pose": pose.grad, "grad_srcs": srcs.grad}
    for k in range(c["S"]):
        out[f"depth{k}"] = d["depth"][k]
        out[f"grad_depth{k}"] = depth[k].grad
    return {k: v.detach().numpy() for k, v in out.items()}


def main():
    for name, c in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(c))
        print("wrote", name)
    s = make_sequence(6, 24, 32, seed=201)
    out = O.consistency(s["depth"], s["pose"], s["K"], s["frames"])
    np.savez_compressed(os.path.join(HERE, "consistency_f6_24x32.npz"), frames=s["frames"].numpy(), depth=s["depth"].numpy(),
                        pose=s["pose"].numpy(), K=s["K"].numpy(), out=out.numpy())
    print("wrote consistency_f6_24x32")


if __name__ == "__main__":
    main()
