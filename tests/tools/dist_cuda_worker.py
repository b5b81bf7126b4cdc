"""One rank of the multi-GPU parity check of `coivo_b200.dist.sharded_loss` (SURVEY.md section 4 "Distributed: per-rank
shard parity vs oracle on that shard; all-reduced loss == oracle loss on full batch").  Launched by
tests/test_dist_cuda.py under `python -m torch.distributed.run`, one process per rank:

  --backend nccl   one GPU per rank (LOCAL_RANK), the production set-up
  --backend gloo   every rank on cuda:0 (a one-GPU box): the CUDA path computes, gloo carries the 16-byte all-reduce

Each rank runs the CUDA operator on its shard of a B-triplet batch (B = 5 over 2 ranks: uneven shards), back-propagates
the local loss that sharded_loss returns, and the per-sample gradients are assembled across ranks; rank 0 then checks
the all-reduced loss and every assembled gradient against the CPU oracle on the FULL batch (rel 1e-4, the oracle run
with the kernels' own arg-min / (a, b) as everywhere else)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--batch", type=int, default=5)
    ap.add_argument("--reduce", default="sum")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)) if a.backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(a.backend, rank=rank, world_size=world)
    import coivo_b200
    from coivo_b200 import dist as cdist
    from coivo_b200.synthetic import make_triplets

    B, H, W = a.batch, 48, 64
    d = make_triplets(B, H, W, seed=77)
    batch = {k: ([x.to(dev) for x in v] if isinstance(v, list) else v.to(dev)) for k, v in d.items()}
    lo, hi = cdist.shard_range(B, rank, world)
    # full-batch leaves: sharded_loss slices out this rank's shard, so the gradients are non-zero on [lo, hi) only
    batch["depth"] = [x.requires_grad_() for x in batch["depth"]]
    batch["pose"].requires_grad_()
    batch["srcs"].requires_grad_()
    scaled, glob = cdist.sharded_loss(coivo_b200.photometric_loss, batch, reduce=a.reduce)
    scaled.backward()
    sh = cdist.shard_batch({k: ([x.detach() for x in v] if isinstance(v, list) else v.detach()) for k, v in batch.items()}, rank, world)
    with torch.no_grad():
        _, valid, sel, ab = coivo_b200.photometric_loss(sh["depth"], sh["pose"], sh["K"], sh["tgt"], sh["srcs"], return_masks=True)

    def assemble(local, full_shape, dtype=torch.float32):
        buf = torch.zeros(full_shape, dtype=dtype, device=dev)
        buf[lo:hi] = local.to(dtype)
        dist.all_reduce(buf)
        return buf.cpu()

    gscale = 1.0 if a.reduce == "sum" else 1.0 / world      # 'mean': the caller averages the per-rank gradients
    for x in batch["depth"] + [batch["pose"], batch["srcs"]]:      # nothing leaks outside the shard
        assert x.grad[:lo].abs().sum().item() == 0 and x.grad[hi:].abs().sum().item() == 0
    g_depth = [assemble(x.grad[lo:hi] * gscale, x.shape) for x in batch["depth"]]
    g_pose = assemble(batch["pose"].grad[lo:hi] * gscale, batch["pose"].shape)
    g_srcs = assemble(batch["srcs"].grad[lo:hi] * gscale, batch["srcs"].shape)
    sel_f = assemble(sel, (B,) + tuple(sel.shape[1:]), torch.int32).to(torch.uint8)
    ab_f = assemble(ab, (B,) + tuple(ab.shape[1:]))
    valid_f = assemble(valid, (B,) + tuple(valid.shape[1:]), torch.int32).to(torch.uint8)
    ok = True
    if rank == 0:
        from oracle import photometric as O
        from test_gpu_parity import assert_close_upto_kinks, relinf, TOL
        with torch.no_grad():
            l0, v0, s0, ab0 = O.photometric_loss(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], return_masks=True)
        assert torch.equal(valid_f, v0), "valid mask must be bit-exact on every shard"
        assert abs(glob.item() - l0.item()) <= TOL * abs(l0.item()), (glob.item(), l0.item())
        od = [x.clone().requires_grad_() for x in d["depth"]]
        op, osr = d["pose"].clone().requires_grad_(), d["srcs"].clone().requires_grad_()
        O.photometric_loss(od, op, d["K"], d["tgt"], osr, sel_override=sel_f, ab_override=ab_f).backward()
        kinks = O.l1_kink_count(d["depth"], d["pose"], d["K"], d["tgt"], d["srcs"], sel_f, ab_f)
        for k in range(len(od)):
            assert_close_upto_kinks(g_depth[k], od[k].grad, kinks, f"grad_depth[{k}]")
        assert relinf(g_pose[:, :, :3], op.grad[:, :, :3]) < TOL
        assert_close_upto_kinks(g_srcs, osr.grad, kinks, "grad_srcs")
        print(f"DIST_OK backend={a.backend} world={world} B={B} reduce={a.reduce} global_loss={glob.item():.7f} oracle={l0.item():.7f} "
              f"sel_mism={int((sel_f != s0).sum())} kinks={kinks}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
