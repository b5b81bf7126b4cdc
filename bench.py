#!/usr/bin/env python
"""bench.py -- the path's benchmark contract.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --config {2,4,5} ...                     (BASELINE.json configs[1] / [3] / [4]; default 2)
    python bench.py --scaling strong --gpus N ...            (global batch 24 of config 2's shape split over N ranks)
    python bench.py --impl reference [--steps K --warmup W]  (CPU reference arm)

Config 2 / 4: a step = one forward + backward of `coivo_b200.photometric_loss` over one batch of synthetic frame
triplets (12 x 256x320, or 4 x 1080x1350; N = 2 sources, S = 4 scales, fp32) per GPU -- weak scaling: the path shards
by triplet with no data-path collective, the only exchange is the scalar-loss all-reduce.  `value` is whole-job
triplets/s (= target frames/s) with the inputs resident in HBM; `e2e` is the same metric through
`colvo_photo_step_host` with pinned HOST buffers (H2D of every input + fwd + bwd + D2H of the loss inside the timed
region; the gradients stay on the device, where a training step consumes them -- `e2e.full_d2h` also copies every
gradient back).  Config 5: a step = one forward-only consistency sweep over a 2000-frame sequence
(`coivo_b200.consistency`), `value` = frame pairs/s.  Prints ONE JSON line on rank 0.
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

N_SRC, S = 2, 4
METRIC = "photometric-loss fwd+bwd frames/s at 1/2/4/8 B200; HBM GB/s vs peak"
UNIT = "frames/s"
CONFIGS = {
    2: dict(kind="loss", B=12, H=256, W=320, tag="BASELINE configs[1]"),
    4: dict(kind="loss", B=4, H=1080, W=1350, tag="BASELINE configs[3], bandwidth stress, rows not 16-byte aligned"),
    5: dict(kind="sweep", F=2000, H=256, W=320, tag="BASELINE configs[4]"),
}
STRONG_GLOBAL_BATCH = 24          # 12 does not divide by 8
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # SMs x lanes x FMA x max SM clock = 74.4


def workload_name(cfg):
    if cfg["kind"] == "sweep":
        return (f"ColVO inference-time warp + LCC consistency sweep, {cfg['F']}-frame sequence {cfg['H']}x{cfg['W']}, "
                f"forward only, fp32 ({cfg['tag']})")
    return (f"ColVO training loss, batch {cfg['B']} triplets {cfg['H']}x{cfg['W']}, N={N_SRC}, S={S}, fwd+bwd, fp32 "
            f"({cfg['tag']})")


def alg_bytes_per_triplet(h, w, n=N_SRC, s=S):
    """B_alg = 2*bytes_in + bytes_grad; also the per-kernel split."""
    hw = h * w
    pyr = sum((h >> k) * (w >> k) for k in range(s))
    bytes_in = 4 * (3 * hw + 3 * n * hw + pyr)
    bytes_grad = 4 * (3 * n * hw + pyr)
    return {"step": 2 * bytes_in + bytes_grad, "fwd": bytes_in, "bwd": bytes_in + bytes_grad}


def alg_bytes_per_pair(h, w):
    """target 3 + source 3 + depth 1 planes per pair, no cross-pair reuse assumed."""
    return 4 * h * w * 7


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_record(cfg_id):
    """profiles/traffic.json: ncu-measured DRAM bytes and counted fp32 flops per kernel and per step, written by
    scripts/make_traffic.py from an `ncu --set full` capture; it records the hash of the kernel sources it was taken
    from, so a stale record is reported as such instead of silently."""
    from coivo_b200 import _lib
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, None
    with open(tp) as f:
        rec = json.load(f)
    ent = rec.get("configs", {}).get(str(cfg_id))
    if ent is None:
        return None, None
    stale = ent.get("src_sha16") != _lib.source_hash()
    return ent, stale


# ---------------------------------------------------------------------------------------------
# CPU legs (the oracle is the only CPU implementation of the path: upstream ships no code)
# ---------------------------------------------------------------------------------------------
def cpu_oracle_rate(cfg, steps, warmup, seed=0, budget_s=None):
    """The CPU oracle (oracle/photometric.py, PyTorch, all host threads) on the same workload.
    With `budget_s`, each step is a bounded sample: the first `b` triplets (pairs) of the batch, `b` chosen from one probe
    pass so that warmup + steps passes fit the budget (the rate is per triplet, so it extrapolates).
    Returns (units/s, mean ms per step, threads, units per step, description)."""
    import torch
    from coivo_b200.synthetic import make_triplets, make_sequence
    from oracle import photometric as O

    torch.set_num_threads(os.cpu_count() or 1)
    H, W = cfg["H"], cfg["W"]
    if cfg["kind"] == "sweep":
        full = 64                                             # the oracle sweep is per pair: a 65-frame slice stands for 2000
        s = make_sequence(full + 1, H, W, seed=seed)

        def one(b):
            t0 = time.perf_counter()
            O.consistency(s["depth"][:b + 1], s["pose"][:b], s["K"], s["frames"][:b + 1])
            return time.perf_counter() - t0
        what = "forward-only sweeps over the first {b} pairs of the sequence (oracle.consistency)"
    else:
        full = cfg["B"]
        d = make_triplets(full, H, W, N=N_SRC, S=S, seed=seed)

        def one(b):
            depth = [x[:b].clone().requires_grad_() for x in d["depth"]]
            pose = d["pose"][:b].clone().requires_grad_()
            srcs = d["srcs"][:b].clone().requires_grad_()
            t0 = time.perf_counter()
            O.photometric_loss(depth, pose, d["K"][:b], d["tgt"][:b], srcs).backward()
            return time.perf_counter() - t0
        what = "fwd+bwd passes over the first {b} of the batch's " + str(full) + " triplets (oracle/photometric.py)"

    b = full
    if budget_s is not None:
        one(1)                                   # page in / thread pool
        per_unit = one(2) / 2.0
        b = int(budget_s / max((steps + warmup) * per_unit, 1e-9))
        b = max(1, min(full, b))
    times = []
    for it in range(warmup + steps):
        dt = one(b)
        if it >= warmup:
            times.append(dt)
    return b / statistics.median(times), sum(times) / len(times) * 1e3, torch.get_num_threads(), b, what.format(b=b)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps or 3, (args.warmup if args.warmup is not None else 1)
    rate, ms, threads, b_s, what = cpu_oracle_rate(cfg, steps, warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "note": "upstream ships no code: the CPU oracle port is the reference arm"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} {what} (bounded to ~150 s in total)"},
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


def bind_numa(local_rank):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node BEFORE any pinned buffer is allocated, so the
    staging memory of the end-to-end step is local to the GPU's PCIe root (ranks sharing one socket's memory share its
    H2D rate).  Reads nvidia-smi's topology matrix; does nothing when that is unavailable."""
    if os.environ.get("COLVO_NO_AFFINITY"):
        return None
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for ln in out.splitlines():
            f = ln.replace("\x1b[4m", "").replace("\x1b[0m", "").split()
            if f and f[0] == f"GPU{local_rank}":
                for tok in f[1:]:
                    if tok.replace("-", "").replace(",", "").isdigit() and ("-" in tok or "," in tok):   # CPU affinity column, e.g. 0-31,64-95
                        cpus = set()
                        for part in tok.split(","):
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                        cpus &= os.sched_getaffinity(0)
                        if cpus:
                            os.sched_setaffinity(0, cpus)
                            return sorted(cpus)
                        return None
    except Exception:
        pass
    return None


class Env:
    """torch / distributed / library set-up shared by the GPU legs."""

    def __init__(self, args):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        # keep stdout to the one JSON line: NCCL / torch banners written to fd 1 go to stderr until the end
        sys.stdout.flush()
        self.saved_stdout = os.dup(1)
        os.dup2(2, 1)
        self.cpus = bind_numa(self.local)
        import torch
        import torch.distributed as dist
        from coivo_b200 import _lib
        self.torch, self.dist = torch, dist
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _lib.load()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def emit(self, obj):
        sys.stdout.flush()
        os.dup2(self.saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def time_loop(env, fn, steps, after=None):
    torch = env.torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record()
    for i in range(steps):
        fn(i)
    if after is not None:
        after()
    e1.record()
    env.barrier()
    return e0.elapsed_time(e1)


# ---------------------------------------------------------------------------------------------
# configs 2 / 4: training loss, forward + backward
# ---------------------------------------------------------------------------------------------
def run_loss(args, cfg, cfg_id):
    env = Env(args)
    torch, dist, dev, world, rank, lib = env.torch, env.dist, env.dev, env.world, env.rank, env.lib
    import coivo_b200
    from coivo_b200 import _lib
    from coivo_b200.synthetic import make_triplets

    steps = args.steps or (1000 if cfg_id == 2 else 100)
    warmup = args.warmup if args.warmup is not None else 20
    H, W = cfg["H"], cfg["W"]
    if args.scaling == "strong":
        q, r = divmod(STRONG_GLOBAL_BATCH, world)
        B_loc = q + (1 if rank < r else 0)
        B_glob = STRONG_GLOBAL_BATCH
    else:
        B_loc, B_glob = cfg["B"], cfg["B"] * world

    # Working set larger than L2: rotate over R independent batches (inputs + gradients + saved state).
    l2 = torch.cuda.get_device_properties(dev).L2_cache_size
    ab = alg_bytes_per_triplet(H, W)
    per_batch = B_loc * ab["bwd"]
    R = max(2, -(-3 * l2 // per_batch))
    batches = []
    for r_ in range(R):
        d = make_triplets(B_loc, H, W, N=N_SRC, S=S, seed=1000 * rank + r_)
        batches.append({
            "depth": [x.to(dev).requires_grad_() for x in d["depth"]],
            "pose": d["pose"].to(dev).requires_grad_(),
            "srcs": d["srcs"].to(dev).requires_grad_(),
            "K": d["K"].to(dev), "tgt": d["tgt"].to(dev), "host": d if r_ == 0 else None,
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

    for i in range(max(warmup, 3)):
        step(i)
    env.barrier()

    # one (start, stop) event pair per timed step for the dominant kernel (k_photo_bwd unless --kernel says otherwise)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b_ in kev:
        a.record(); b_.record()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(env.local) if rank == 0 else None
    t_wall = time.perf_counter()

    def timed_step(i):
        lib.colvo_debug_time_kernel(args.kernel, kev[i][0].cuda_event, kev[i][1].cuda_event)
        step(warmup + i)
    ms_total = time_loop(env, timed_step, steps)
    t_wall = time.perf_counter() - t_wall
    lib.colvo_debug_time_kernel(0, None, None)
    kern_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in kev)

    if args.profile:
        if sampler:
            sampler.stop()
        if rank == 0:
            env.emit({"profile_run": True, "config": cfg_id, "ms_per_step": ms_total / steps, "kernel": args.kernel, "kernel_ms": kern_ms})
        env.close()
        return

    # CUDA-graph leg: the same K steps, each batch's forward+backward captured once and replayed
    graph_ms = None
    if not args.no_graph:
        graphs = [coivo_b200.GraphedStep(b["depth"], b["pose"], b["K"], b["tgt"], b["srcs"]) for b in batches]
        for i in range(max(warmup, 3)):
            graphs[i % R].replay()

        def graph_step(i):
            loss = graphs[(warmup + i) % R].replay()
            if world > 1:
                loss_buf.copy_(loss.detach().reshape(1))
                dist.all_reduce(loss_buf, async_op=True)
        graph_ms = time_loop(env, graph_step, steps)
        del graphs

    # The same step as a training loop runs it: the frames are data, not parameters -- no gradient with respect to the
    # source images (no scatter, no zero-fill, no unpack).  Reported beside `value`, which follows SURVEY.md section 8(d)
    # (every gradient of row 11, grad_srcs included).
    nosrc_ms = merged_ms = None
    if not args.no_graph:
        srcs_const = [b["srcs"].detach() for b in batches]
        for b in batches:
            for t in b["depth"] + [b["pose"]]:
                t.grad = None
        graphs = [coivo_b200.GraphedStep(b["depth"], b["pose"], b["K"], b["tgt"], sc) for b, sc in zip(batches, srcs_const)]
        for i in range(max(warmup, 3)):
            graphs[i % R].replay()
        nosrc_ms = time_loop(env, lambda i: graphs[(warmup + i) % R].replay(), steps)
        del graphs
        # ... and with the warp-aggregated scatter (COLVO_F_SCATTER_MERGE), the form north_star names; it is the slower one
        for b in batches:
            for t in b["depth"] + [b["pose"], b["srcs"]]:
                t.grad = None
        graphs = [coivo_b200.GraphedStep(b["depth"], b["pose"], b["K"], b["tgt"], b["srcs"], scatter="merged") for b in batches]
        for i in range(max(warmup, 3)):
            graphs[i % R].replay()
        merged_ms = time_loop(env, lambda i: graphs[(warmup + i) % R].replay(), steps)
        del graphs

    # end-to-end legs: pinned host buffers -> H2D -> fwd -> bwd -> D2H, through the C ABI.
    #   "device": the loss is read back, the gradients stay in HBM (a training step consumes them there)
    #   "host":   every gradient is copied back as well
    #   "u8":     as "device", with the frames handed over as uint8 (what a video loader holds) and widened on the GPU
    hb = batches[0]["host"]
    pin = lambda t: t.pin_memory()
    h_in = ([pin(x) for x in hb["depth"]], pin(hb["pose"]), pin(hb["K"]), pin(hb["tgt"]), pin(hb["srcs"]))
    to_u8 = lambda t: (t * 255.0).round().clamp_(0, 255).to(torch.uint8)
    h_in_u8 = (h_in[0], h_in[1], h_in[2], pin(to_u8(hb["tgt"])), pin(to_u8(hb["srcs"])))
    e2e_steps = steps
    e2e = {}
    for mode in ("device", "host", "u8"):
        stepper = coivo_b200.HostStepper(B_loc, N_SRC, S, H, W, device=dev, grads="host" if mode == "host" else "device",
                                         images="u8" if mode == "u8" else "f32")
        inp = h_in_u8 if mode == "u8" else h_in
        for _ in range(3):
            stepper.step(*inp)
        stepper.finish()
        ms = time_loop(env, lambda _i: stepper.step(*inp), e2e_steps, after=stepper.join)
        stepper.finish()
        e2e[mode] = {"ms": ms, "h2d": stepper.h2d_bytes(*inp), "d2h": stepper.d2h_bytes(), "chunks": len(stepper.spans)}
        del stepper
    clocks = sampler.stop() if sampler else None

    # Reported separately (assumption A13): the training loop's gradient all-reduce of the
    # out-of-scope depth / pose CNNs (~28 M fp32 parameters) over NCCL / NVLink -- not part of the path or of `value`.
    ddp = None
    if world > 1:
        gbuf = torch.zeros(28_000_000, dtype=torch.float32, device=dev)
        for _ in range(3):
            dist.all_reduce(gbuf)
        ar_ms = time_loop(env, lambda _i: dist.all_reduce(gbuf), 10) / 10
        ddp = {"bytes": gbuf.numel() * 4, "ms": env.max_over_ranks([ar_ms])[0],
               "note": "dummy 28 M-parameter fp32 gradient all-reduce (NCCL), timed on its own"}
        del gbuf

    ms_total, e2e_ms, kern_ms, graph_ms_max, e2e_full_ms, e2e_u8_ms, nosrc_ms_max, merged_ms_max = env.max_over_ranks(
        [ms_total, e2e["device"]["ms"], kern_ms, graph_ms if graph_ms is not None else 0.0, e2e["host"]["ms"], e2e["u8"]["ms"],
         nosrc_ms if nosrc_ms is not None else 0.0, merged_ms if merged_ms is not None else 0.0])
    eager_ms = ms_total
    launch = "eager launches through the autograd.Function"
    if graph_ms is not None and graph_ms_max < ms_total:
        ms_total, launch = graph_ms_max, "CUDA-graph replay of the captured forward+backward (coivo_b200.GraphedStep)"

    if rank == 0:
        peak, peak_src = measured_peaks()
        rate = B_glob * steps / (ms_total * 1e-3)
        B_max = -(-B_glob // world)                 # the largest shard sets the step time
        kname = {1: "k_photo_fwd", 2: "k_photo_bwd", 3: "k_warp_stats"}[args.kernel]
        kern_bytes = B_max * (ab["bwd"] if args.kernel == 2 else ab["fwd"])
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
        step_ms = ms_total / steps
        step_gbs = B_max * ab["step"] / (step_ms * 1e-3) / 1e9
        rec, stale = ncu_record(cfg_id) if args.scaling == "weak" else (None, None)
        traffic = step_traffic = fp32 = None
        if rec is not None:
            traffic = rec["kernels"].get(kname, {}).get("dram_bytes_per_launch")
            step_traffic = rec["step"]["dram_bytes"]
            gf = rec["step"]["fp32_flop"] / B_max / 1e9
            tf = gf * B_max / (step_ms * 1e-3) / 1e3
            fp32 = {"gflop_per_triplet": gf, "achieved_tflops": tf, "peak_tflops": FP32_PEAK_TFLOPS, "frac": tf / FP32_PEAK_TFLOPS,
                    "warp_instructions_per_step": rec["step"]["warp_inst"],
                    "issue_slot_frac": rec["step"]["warp_inst"] / (148 * 4 * 1.965e9 * step_ms * 1e-3),
                    "source": "counted from the SASS of an ncu capture (profiles/traffic.json, scripts/make_traffic.py): "
                              "predicated-on thread instructions x flops per opcode (FFMA 2, FFMA2 4, FADD/FMUL 1, FADD2/FMUL2 2, MUFU 1)",
                    "stale": stale}
        scale = lambda ms, n: B_glob * n / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg) if args.scaling == "weak" else
                       f"ColVO training loss, GLOBAL batch {B_glob} triplets {H}x{W} split over {world} GPU(s), N={N_SRC}, S={S}, fwd+bwd, fp32 "
                       "(BASELINE configs[2], strong scaling: SURVEY.md section 8(e))",
                       "global_batch": B_glob, "per_gpu_batch": B_max, "parallelism": f"dp{world} (batch-sharded triplets)",
                       "l2_policy": f"inputs larger than L2: {R} rotating batches, {R * per_batch / 1e6:.0f} MB > L2 {l2 / 1e6:.0f} MB",
                       "frames_per_triplet": "1 target + 2 sources; frames/s counts target frames (= triplets/s)",
                       "launch": launch, "eager_ms_per_step": eager_ms / steps,
                       "graph_ms_per_step": (graph_ms_max / steps) if graph_ms is not None else None,
                       "eager_wall_ms_per_step": t_wall / steps * 1e3, "ddp_dummy_allreduce": ddp,
                       "without_image_gradient": None if nosrc_ms is None else {
                           "ms_per_step": nosrc_ms_max / steps, "frames_per_s": B_glob * steps / (nosrc_ms_max * 1e-3),
                           "note": "the same step with srcs not requiring grad (what a training loop asks for): graph replay; "
                                   "not the headline, which includes grad_srcs as SURVEY.md section 8(a) row 11 does"},
                       "warp_aggregated_scatter": None if merged_ms is None else {
                           "ms_per_step": merged_ms_max / steps,
                           "note": "the headline step with scatter='merged' (COLVO_F_SCATTER_MERGE: coincident taps summed in the warp, "
                                   "half the global reductions); graph replay.  Slower than the default vector-RED scatter, hence opt-in"},
                       "cpu_affinity": f"{len(env.cpus)} CPUs of the GPU's NUMA node" if env.cpus else "unchanged"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_stale": stale, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kern_bytes, "kernel_ms": kern_ms,
                         "kernel_share_of_step": kern_ms / step_ms,
                         "step": {"achieved": step_gbs, "frac": step_gbs / peak, "algorithmic_bytes_per_step": B_max * ab["step"],
                                  "traffic": step_traffic},
                         "fp32": fp32},
            "e2e": {"value": scale(e2e_ms, e2e_steps), "unit": UNIT,
                    "h2d_bytes_per_step": e2e["device"]["h2d"], "d2h_bytes_per_step": e2e["device"]["d2h"],
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": f"colvo_photo_step_host (pinned host buffers, fp32 frames; {e2e['device']['chunks']} batch chunks on their own "
                           "streams overlap H2D and kernels; the loss is read back, the gradients stay in HBM)",
                    "full_d2h": {"value": scale(e2e_full_ms, e2e_steps), "unit": UNIT,
                                 "d2h_bytes_per_step": e2e["host"]["d2h"], "ms_per_step": e2e_full_ms / e2e_steps,
                                 "note": "every gradient (depth, pose, sources) copied back to pinned host memory as well"},
                    "uint8_frames": {"value": scale(e2e_u8_ms, e2e_steps), "unit": UNIT,
                                     "h2d_bytes_per_step": e2e["u8"]["h2d"], "d2h_bytes_per_step": e2e["u8"]["d2h"],
                                     "ms_per_step": e2e_u8_ms / e2e_steps,
                                     "note": "same step with tgt / srcs handed over as uint8 (what a video loader holds) and widened to "
                                             "fp32 on the device (COLVO_F_HOST_U8): a different input contract, so reported beside the "
                                             "fp32 figure, not as it"}},
            "gpu_launches": steps * (len(_lib.KERNELS_FWD) + len(_lib.KERNELS_BWD)),
            "clocks": clocks,
        }
        if world == 1 and args.scaling == "weak":
            cpu_rate, cpu_ms, cpu_threads, b_s, what = cpu_oracle_rate(cfg, 3, 1, budget_s=None if cfg_id == 2 else 25.0)
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                    "sample": "3 " + what, "ms_per_step": cpu_ms}
        env.emit(line)
    env.close()


# ---------------------------------------------------------------------------------------------
# config 5: inference-time consistency sweep, forward only
# ---------------------------------------------------------------------------------------------
def run_sweep(args, cfg, cfg_id):
    env = Env(args)
    torch, dev, world, rank, lib = env.torch, env.dev, env.world, env.rank, env.lib
    import coivo_b200
    from coivo_b200.synthetic import make_sequence

    steps = args.steps or 20
    warmup = args.warmup if args.warmup is not None else 3
    F, H, W = cfg["F"], cfg["H"], cfg["W"]
    s = make_sequence(F, H, W, seed=7 + rank)           # every rank sweeps its own sequence (independent sequences: replicas)
    d_in = (s["depth"].to(dev), s["pose"].to(dev), s["K"].to(dev), s["frames"].to(dev))

    def step(_i):
        coivo_b200.consistency(*d_in)
    for i in range(max(warmup, 3)):
        step(i)
    env.barrier()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b_ in kev:
        a.record(); b_.record()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(env.local) if rank == 0 else None
    kid = args.kernel if args.kernel in (3, 4) else 4

    def timed_step(i):
        lib.colvo_debug_time_kernel(kid, kev[i][0].cuda_event, kev[i][1].cuda_event)
        step(i)
    ms_total = time_loop(env, timed_step, steps)
    lib.colvo_debug_time_kernel(0, None, None)
    kern_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in kev)
    if args.profile:
        if sampler:
            sampler.stop()
        if rank == 0:
            env.emit({"profile_run": True, "config": cfg_id, "ms_per_step": ms_total / steps, "kernel": kid, "kernel_ms": kern_ms})
        env.close()
        return

    # end to end: the sequence sits in pinned host memory; H2D of frames + depth + poses, sweep, D2H of the [F-1, 4] result
    h = [t.pin_memory() for t in (s["depth"], s["pose"], s["K"], s["frames"])]
    d_buf = [torch.empty_like(t, device=dev) for t in h]
    h_out = torch.empty(F - 1, 4).pin_memory()

    def e2e_step(_i):
        for a, b_ in zip(d_buf, h):
            a.copy_(b_, non_blocking=True)
        h_out.copy_(coivo_b200.consistency(*d_buf), non_blocking=True)
    for i in range(2):
        e2e_step(i)
    e2e_steps = max(3, steps // 4)
    e2e_ms = time_loop(env, e2e_step, e2e_steps)
    clocks = sampler.stop() if sampler else None
    ms_total, kern_ms, e2e_ms = env.max_over_ranks([ms_total, kern_ms, e2e_ms])
    if rank == 0:
        peak, peak_src = measured_peaks()
        pairs = F - 1
        step_ms = ms_total / steps
        rate = world * pairs * steps / (ms_total * 1e-3)
        pass_pairs = min(pairs, (256 << 20) // (H * W * 16))      # colvo_api.cu::consistency_pairs_per_pass
        kname = {3: "k_warp_stats", 4: "k_consistency_pe"}[kid]
        kern_bytes = pass_pairs * alg_bytes_per_pair(H, W)
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
        step_gbs = pairs * alg_bytes_per_pair(H, W) / (step_ms * 1e-3) / 1e9
        rec, stale = ncu_record(cfg_id)
        passes = -(-pairs // pass_pairs)
        line = {
            "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "pairs_per_sweep": pairs,
                       "parallelism": f"{world} independent sequence(s), one per GPU (replicas)",
                       "l2_policy": f"inputs larger than L2: one sweep reads {pairs * alg_bytes_per_pair(H, W) / 1e6:.0f} MB",
                       "frames_per_unit": "frames/s counts frame pairs (= frames of the sequence) per second",
                       "launch": "eager launches through coivo_b200.consistency"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": rec["kernels"].get(kname, {}).get("dram_bytes_per_launch") if rec else None,
                         "traffic_stale": stale, "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes,
                         "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms * passes / step_ms,
                         "note": f"the sweep runs in {passes} passes of <= {pass_pairs} pairs; the timed launch is the first pass's",
                         "step": {"achieved": step_gbs, "frac": step_gbs / peak, "algorithmic_bytes_per_step": pairs * alg_bytes_per_pair(H, W),
                                  "traffic": rec["step"]["dram_bytes"] if rec else None}},
            "e2e": {"value": world * pairs * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": sum(t.numel() * 4 for t in h), "d2h_bytes_per_step": h_out.numel() * 4,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": "coivo_b200.consistency on device copies of a pinned-host sequence (frames, depth, poses copied in, [F-1,4] result copied back)"},
            "gpu_launches": steps * (3 * passes + 1),
            "clocks": clocks,
        }
        if world == 1:
            cpu_rate, cpu_ms, cpu_threads, b_s, what = cpu_oracle_rate(cfg, 3, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                    "sample": "3 " + what, "ms_per_step": cpu_ms}
        env.emit(line)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config: 2 (default, the headline), 4 (1080x1350), 5 (2000-frame sweep)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: global batch 24 of config 2's shape split over the ranks")
    ap.add_argument("--profile", action="store_true", help="profiling run: skip the graph, e2e and CPU-baseline legs")
    ap.add_argument("--kernel", type=int, default=2,
                    help="which kernel the live event bracket times: 1 k_photo_fwd, 2 k_photo_bwd (default, the dominant one), "
                         "3 k_warp_stats, 4 k_consistency_pe (config 5)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches only (no CUDA-graph replay leg)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.scaling == "strong" and args.config != 2:
        ap.error("--scaling strong is defined on config 2's shape")
    if args.impl == "reference":
        run_reference(args, cfg)
    elif cfg["kind"] == "sweep":
        run_sweep(args, cfg, args.config)
    else:
        run_loss(args, cfg, args.config)


if __name__ == "__main__":
    main()
