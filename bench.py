#!/usr/bin/env python
"""bench.py -- the path's benchmark contract.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference [--steps K --warmup W]  (CPU reference arm)

A step = one forward + backward of `coivo_b200.photometric_loss` over one batch of synthetic
frame triplets.  Workload = BASELINE.json configs[1]: 12 triplets of 256x320, N = 2 sources,
S = 4 scales, fp32, per GPU (weak scaling: global batch 12 x N_gpus; the path shards by triplet
with no data-path collective -- the only exchange is the scalar-loss all-reduce).
`value` is whole-job triplets/s (= target frames/s) with the inputs resident in HBM; `e2e` is
the same metric through `colvo_photo_step_host` with pinned HOST buffers (H2D of every input +
fwd + bwd + D2H of the loss inside the timed region; the gradients stay on the device, where a
training step consumes them -- `e2e.full_d2h` is the variant that also copies every gradient
back).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, H, W, N_SRC, S = 12, 256, 320, 2, 4
METRIC = "photometric-loss fwd+bwd frames/s at 1/2/4/8 B200; HBM GB/s vs peak"
UNIT = "frames/s"
WORKLOAD = f"ColVO training loss, batch {B_PER_GPU} triplets {H}x{W}, N={N_SRC}, S={S}, fwd+bwd, fp32 (BASELINE configs[1])"


def alg_bytes_per_triplet(h=H, w=W, n=N_SRC, s=S):
    """B_alg = 2*bytes_in + bytes_grad; also the per-kernel split."""
    hw = h * w
    pyr = sum((h >> k) * (w >> k) for k in range(s))
    bytes_in = 4 * (3 * hw + 3 * n * hw + pyr)
    bytes_grad = 4 * (3 * n * hw + pyr)
    return {"step": 2 * bytes_in + bytes_grad, "fwd": bytes_in, "bwd": bytes_in + bytes_grad}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_oracle_rate(steps, warmup, seed=0, budget_s=None):
    """The CPU oracle (oracle/photometric.py, PyTorch, all host threads) on the same workload.
    With `budget_s`, each step is a bounded sample: the first `b` triplets of the batch, `b` chosen from one probe
    pass so that warmup + steps passes fit the budget (the rate is per triplet, so it extrapolates).
    Returns (triplets/s, mean ms per step, threads, triplets per step)."""
    import torch
    from coivo_b200.synthetic import make_triplets
    from oracle import photometric as O

    torch.set_num_threads(os.cpu_count() or 1)
    d = make_triplets(B_PER_GPU, H, W, N=N_SRC, S=S, seed=seed)

    def one(b):
        depth = [x[:b].clone().requires_grad_() for x in d["depth"]]
        pose = d["pose"][:b].clone().requires_grad_()
        srcs = d["srcs"][:b].clone().requires_grad_()
        t0 = time.perf_counter()
        O.photometric_loss(depth, pose, d["K"][:b], d["tgt"][:b], srcs).backward()
        return time.perf_counter() - t0

    b = B_PER_GPU
    if budget_s is not None:
        one(1)                                   # page in / thread pool
        per_triplet = one(2) / 2.0
        b = int(budget_s / max((steps + warmup) * per_triplet, 1e-9))
        b = max(1, min(B_PER_GPU, b))
    times = []
    for it in range(warmup + steps):
        dt = one(b)
        if it >= warmup:
            times.append(dt)
    return b / statistics.median(times), sum(times) / len(times) * 1e3, torch.get_num_threads(), b


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps or 3, (args.warmup if args.warmup is not None else 1)
    rate, ms, threads, b_s = cpu_oracle_rate(steps, warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "upstream ships no code: the CPU oracle port is the reference arm"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} fwd+bwd passes over the first {b_s} of the batch's {B_PER_GPU} triplets (bounded to ~150 s in total)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import coivo_b200
    from coivo_b200 import _lib
    from coivo_b200.synthetic import make_triplets

    steps = args.steps or 1000
    warmup = args.warmup if args.warmup is not None else 20
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the path has no CPU fallback")
    # keep stdout to the one JSON line: NCCL / torch banners written to fd 1 go to stderr until the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # Working set larger than L2: rotate over R independent batches (inputs + gradients + saved state).
    l2 = torch.cuda.get_device_properties(dev).L2_cache_size
    ab = alg_bytes_per_triplet()
    per_batch = B_PER_GPU * (ab["fwd"] + (ab["bwd"] - ab["fwd"]))
    R = max(2, -(-3 * l2 // per_batch))
    batches = []
    for r in range(R):
        d = make_triplets(B_PER_GPU, H, W, N=N_SRC, S=S, seed=1000 * rank + r)
        batches.append({
            "depth": [x.to(dev).requires_grad_() for x in d["depth"]],
            "pose": d["pose"].to(dev).requires_grad_(),
            "srcs": d["srcs"].to(dev).requires_grad_(),
            "K": d["K"].to(dev), "tgt": d["tgt"].to(dev), "host": d,
        })
    loss_buf = torch.zeros(1, device=dev)

    def step(i):
        b = batches[i % R]
        for t in b["depth"] + [b["pose"], b["srcs"]]:
            t.grad = None
        loss = coivo_b200.photometric_loss(b["depth"], b["pose"], b["K"], b["tgt"], b["srcs"])
        loss.backward()
        if world > 1:                      # the path's only collective: the scalar loss (logging)
            loss_buf.copy_(loss.detach().reshape(1))
            dist.all_reduce(loss_buf, async_op=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(warmup, 3)):
        step(i)
    barrier()

    # one (start, stop) event pair per timed step for the dominant kernel (k_photo_bwd)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b_ in kev:
        a.record(); b_.record()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    e0.record()
    for i in range(steps):
        lib.colvo_debug_time_kernel(args.kernel, kev[i][0].cuda_event, kev[i][1].cuda_event)
        step(warmup + i)
    e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    lib.colvo_debug_time_kernel(0, None, None)
    ms_total = e0.elapsed_time(e1)
    kern_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in kev)

    # CUDA-graph leg: the same K steps, each batch's forward+backward captured once and replayed
    graph_ms = None
    if not args.no_graph and not args.profile:
        graphs = [coivo_b200.GraphedStep(b["depth"], b["pose"], b["K"], b["tgt"], b["srcs"]) for b in batches]
        for i in range(max(warmup, 3)):
            graphs[i % R].replay()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(steps):
            loss = graphs[(warmup + i) % R].replay()
            if world > 1:
                loss_buf.copy_(loss.detach().reshape(1))
                dist.all_reduce(loss_buf, async_op=True)
        g1.record()
        barrier()
        graph_ms = g0.elapsed_time(g1)
    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    if args.profile:
        if sampler:
            sampler.stop()
        if rank == 0:
            emit({"profile_run": True, "ms_per_step": ms_total / steps, "kernel": args.kernel, "kernel_ms": kern_ms})
        if world > 1:
            dist.destroy_process_group()
        return
    # end-to-end legs: pinned host buffers -> H2D -> fwd -> bwd -> D2H, through the C ABI.
    #   "device": the loss is read back, the gradients stay in HBM (a training step consumes them there)
    #   "host":   every gradient is copied back as well
    hb = batches[0]["host"]
    pin = lambda t: t.pin_memory()
    h_in = ([pin(x) for x in hb["depth"]], pin(hb["pose"]), pin(hb["K"]), pin(hb["tgt"]), pin(hb["srcs"]))
    e2e_steps = steps
    e2e = {}
    for mode in ("device", "host"):
        stepper = coivo_b200.HostStepper(B_PER_GPU, N_SRC, S, H, W, device=dev, grads=mode)
        for _ in range(3):
            stepper.step(*h_in)
        stepper.finish()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(e2e_steps):
            stepper.step(*h_in)
        stepper.join()
        f1.record()
        barrier()
        stepper.finish()
        e2e[mode] = {"ms": f0.elapsed_time(f1), "h2d": stepper.h2d_bytes(*h_in), "d2h": stepper.d2h_bytes(),
                     "chunks": len(stepper.spans)}
        del stepper
    clocks = sampler.stop() if sampler else None
    e2e_ms, e2e_full_ms = e2e["device"]["ms"], e2e["host"]["ms"]

    # Reported separately (assumption A13): the training loop's gradient all-reduce of the
    # out-of-scope depth / pose CNNs (~28 M fp32 parameters) over NCCL / NVLink -- not part of the path or of `value`.
    ddp = None
    if world > 1:
        gbuf = torch.zeros(28_000_000, dtype=torch.float32, device=dev)
        for _ in range(3):
            dist.all_reduce(gbuf)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(gbuf)
        a1.record()
        barrier()
        tt = torch.tensor([a0.elapsed_time(a1) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ddp = {"bytes": gbuf.numel() * 4, "ms": tt.item(), "note": "dummy 28 M-parameter fp32 gradient all-reduce (NCCL), timed on its own"}
        del gbuf

    t = torch.tensor([ms_total, e2e_ms, kern_ms, graph_ms if graph_ms is not None else 0.0, e2e_full_ms], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kern_ms, graph_ms_max, e2e_full_ms = t.tolist()
    eager_ms = ms_total
    launch = "eager launches through the autograd.Function"
    if graph_ms is not None and graph_ms_max < ms_total:
        ms_total, launch = graph_ms_max, "CUDA-graph replay of the captured forward+backward (coivo_b200.GraphedStep)"

    if rank == 0:
        peak, peak_src = measured_peaks()
        rate = world * B_PER_GPU * steps / (ms_total * 1e-3)
        kern_bytes = B_PER_GPU * ab["bwd"]
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
        step_gbs = B_PER_GPU * ab["step"] / (ms_total / steps * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("k_photo_bwd_dram_bytes_per_launch")
        cpu_rate, cpu_ms, cpu_threads, _ = cpu_oracle_rate(3, 1) if world == 1 else (None, None, None, None)
        line = {
            "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B_PER_GPU, "parallelism": f"dp{world} (batch-sharded triplets)",
                       "l2_policy": f"inputs larger than L2: {R} rotating batches, {R * per_batch / 1e6:.0f} MB > L2 {l2 / 1e6:.0f} MB",
                       "frames_per_triplet": "1 target + 2 sources; frames/s counts target frames (= triplets/s)",
                       "launch": launch, "eager_ms_per_step": eager_ms / steps,
                       "graph_ms_per_step": (graph_ms_max / steps) if graph_ms is not None else None,
                       "eager_wall_ms_per_step": t_wall / steps * 1e3, "ddp_dummy_allreduce": ddp},
            "roofline": {"bound": "hbm", "kernel": "k_photo_bwd", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kern_bytes, "kernel_ms": kern_ms,
                         "kernel_share_of_step": kern_ms / (ms_total / steps),
                         "step": {"achieved": step_gbs, "frac": step_gbs / peak,
                                  "algorithmic_bytes_per_step": B_PER_GPU * ab["step"]}},
            "e2e": {"value": world * B_PER_GPU * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": e2e["device"]["h2d"], "d2h_bytes_per_step": e2e["device"]["d2h"],
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": f"colvo_photo_step_host (pinned host buffers; {e2e['device']['chunks']} batch chunks on their own streams "
                           "overlap H2D and kernels; the loss is read back, the gradients stay in HBM)",
                    "full_d2h": {"value": world * B_PER_GPU * e2e_steps / (e2e_full_ms * 1e-3), "unit": UNIT,
                                 "d2h_bytes_per_step": e2e["host"]["d2h"], "ms_per_step": e2e_full_ms / e2e_steps,
                                 "note": "every gradient (depth, pose, sources) copied back to pinned host memory as well"}},
            "gpu_launches": steps * (len(_lib.KERNELS_FWD) + len(_lib.KERNELS_BWD)),
            "clocks": clocks,
        }
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                    "sample": f"3 fwd+bwd passes of the full {B_PER_GPU}-triplet batch (oracle/photometric.py)",
                                    "ms_per_step": cpu_ms}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--profile", action="store_true", help="profiling run: skip the e2e and CPU-baseline legs")
    ap.add_argument("--kernel", type=int, default=2, help="which kernel the live event bracket times: 1 k_photo_fwd, 2 k_photo_bwd (default, the dominant one), 3 k_warp_stats")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches only (no CUDA-graph replay leg)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
