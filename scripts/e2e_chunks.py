"""e2e (host buffers -> loss) rate of HostStepper for several chunk counts (B200; prints one line per setting)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, coivo_b200
from coivo_b200.synthetic import make_triplets
dev = torch.device("cuda:0")
d = make_triplets(12, 256, 320, seed=0)
pin = lambda t: t.pin_memory()
h = ([pin(x) for x in d["depth"]], pin(d["pose"]), pin(d["K"]), pin(d["tgt"]), pin(d["srcs"]))
to_u8 = lambda t: (t * 255.0).round().clamp_(0, 255).to(torch.uint8)
h8 = (h[0], h[1], h[2], pin(to_u8(d["tgt"])), pin(to_u8(d["srcs"])))
for mode, images in (("device", "f32"), ("host", "f32"), ("device", "u8")):
    h = h8 if images == "u8" else h
    for chunks in (1, 2, 3, 4, 6, 12):
        st = coivo_b200.HostStepper(12, 2, 4, 256, 320, device=dev, chunks=chunks, grads=mode, images=images)
        for _ in range(5):
            st.step(*h)
        st.finish()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            st.step(*h)
        st.join(); e1.record(); st.finish()
        ms = e0.elapsed_time(e1) / 200
        print(f"grads={mode} images={images} chunks={chunks}: {ms:.3f} ms/step = {12 / ms * 1e3:.0f} triplets/s", flush=True)
        del st
