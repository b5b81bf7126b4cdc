"""Experiment (round 2): does keeping the forward->backward intermediates L2-resident pay?
A batch of 12 triplets (256x320) is processed as 12/b sequential sub-batches of b triplets (forward + backward of a
sub-batch back to back, so its ~37 MB/triplet of saved texels can stay in the 126 MB L2), captured as ONE CUDA graph;
optionally the sub-batches alternate between two streams.  Prints ms per 12 triplets for each b.  Rotating inputs
larger than L2, as bench.py."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coivo_b200
from coivo_b200.synthetic import make_triplets

DEV = torch.device("cuda:0")
B, H, W = 12, 256, 320
SETS = 6


def build(b, nstreams):
    sets = []
    for r in range(SETS):
        d = make_triplets(B, H, W, seed=70 + r)
        subs = []
        for lo in range(0, B, b):
            s = slice(lo, lo + b)
            subs.append(([x[s].contiguous().to(DEV).requires_grad_() for x in d["depth"]], d["pose"][s].contiguous().to(DEV).requires_grad_(),
                         d["K"][s].contiguous().to(DEV), d["tgt"][s].contiguous().to(DEV), d["srcs"][s].contiguous().to(DEV).requires_grad_()))
        sets.append(subs)
    one = torch.ones((), device=DEV)
    streams = [torch.cuda.Stream(DEV) for _ in range(nstreams)]

    def run(subs):
        cur = torch.cuda.current_stream(DEV)
        if nstreams == 1:
            for a in subs:
                coivo_b200.photometric_loss(*a).backward(gradient=one)
            return
        for st in streams:
            st.wait_stream(cur)
        for i, a in enumerate(subs):
            with torch.cuda.stream(streams[i % nstreams]):
                coivo_b200.photometric_loss(*a).backward(gradient=one)
        for st in streams:
            cur.wait_stream(st)

    graphs = []
    side = torch.cuda.Stream(DEV)
    for subs in sets:
        side.wait_stream(torch.cuda.current_stream(DEV))
        with torch.cuda.stream(side):
            for _ in range(2):
                run(subs)
        torch.cuda.current_stream(DEV).wait_stream(side)
        torch.cuda.synchronize()
        for a in subs:
            for t in a[0] + [a[1], a[4]]:
                t.grad = None
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run(subs)
        graphs.append(g)
    return graphs, sets


def timeit(graphs, steps=60, warm=10):
    for i in range(warm):
        graphs[i % SETS].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % SETS].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if __name__ == "__main__":
    for b, ns in ((12, 1), (6, 1), (4, 1), (3, 1), (2, 1), (6, 2), (4, 2), (3, 2), (2, 2), (3, 4), (1, 4)):
        graphs, keep = build(b, ns)
        ms = timeit(graphs)
        print(json.dumps({"sub_batch": b, "streams": ns, "ms_per_12_triplets": ms, "triplets_per_s": B / ms * 1e3}), flush=True)
        del graphs, keep
        torch.cuda.empty_cache()
