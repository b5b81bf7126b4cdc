"""Compare two dumps of gpu_dump.py (an old and a new build of the library on the same case)."""
import sys
import torch
a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
B, H, W, N, S = a["dims"]
rel = lambda x, y: ((x - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()
print("loss", a["loss"], b["loss"], "sel diff", int((a["sel"] != b["sel"]).sum()), "valid diff", int((a["valid"] != b["valid"]).sum()))
print("ab", rel(a["ab"], b["ab"]))
print("frame (n, mx, my, inv, a, b, Ga, Gb) max rel per column:", [rel(a["frame"][:, j], b["frame"][:, j]) for j in range(8)])
print(a["frame"][:, 6:], b["frame"][:, 6:])
print("scale", rel(a["scale"], b["scale"]))
for k in range(S):
    print("gd", k, rel(a["gd"][k], b["gd"][k]))
print("gT", rel(a["gT"], b["gT"]), "gs", rel(a["gs"], b["gs"]))
# coefficient fields where a source won (a = old: .w = winner index, zeros where identity won; b = new: flags in ch 0/1 .w)
sel = b["sel"]
for n in range(N):
    m = (sel == N + n)
    ca, cb = a["coef"][..., :3], b["coef"][..., :3]
    mm = m[:, :, None, :, :, None].expand_as(ca)
    if mm.any():
        print("coef src", n, "max rel over winner windows", ((ca - cb).abs()[mm].max() / ca.abs()[mm].max()).item(), "count", int(m.sum()))
print("new flags ch0.w sum", b["coef"][:, :, 0, :, :, 3].sum().item(), "ch1.w sum", b["coef"][:, :, 1, :, :, 3].sum().item(), "winners", [(int((sel == N + n).sum())) for n in range(N)])
if N == 1:
    print("geo", rel(a["geo"], b["geo"]))
if len(sys.argv) > 3:
    k = 0
    e = (a["gd"][k] - b["gd"][k]).abs()[:, 0]
    sc = a["gd"][k].abs().max()
    torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
    for bb in range(B):
        print("b", bb, "gd0 err map (x 1e-3 of max):")
        print((e[bb] / sc * 1e3))
    print("sel k=0:")
    for bb in range(B):
        print(b["sel"][bb, 0])
