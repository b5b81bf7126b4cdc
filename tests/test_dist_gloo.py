"""N > 1 host logic on CPU: world_size-2 gloo, batch sharding + the scalar-loss all-reduce.
The compute inside each rank is the oracle (the CUDA path cannot run here); what is under test
is coivo_b200.dist."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coivo_b200 import dist as cdist


def test_shard_range_partitions_every_batch():
    for B in (1, 3, 12, 24):
        for world in (1, 2, 4, 8):
            spans = [cdist.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert cdist.shard_range(12, 7, 8) == (11, 12)
    with pytest.raises(ValueError):
        cdist.shard_range(12, 8, 8)


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from coivo_b200.synthetic import make_triplets
    from oracle import photometric as O
    d = make_triplets(B, 24, 32, seed=5)
    batch = {k: ([x.clone().requires_grad_() for x in v] if isinstance(v, list) else v) for k, v in d.items()}
    scaled, glob = cdist.sharded_loss(O.photometric_loss, batch)
    scaled.backward()
    lo, hi = cdist.shard_range(B, rank, world)
    g = torch.zeros_like(d["depth"][0])
    g[lo:hi] = batch["depth"][0].grad[lo:hi] if batch["depth"][0].grad is not None else 0
    dist.all_reduce(g)          # assemble the per-sample gradients (stay local in production)
    if rank == 0:
        q.put((glob.item(), g))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 3])
def test_world2_sharded_loss_equals_full_batch(B):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    glob, g = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from coivo_b200.synthetic import make_triplets
    from oracle import photometric as O
    d = make_triplets(B, 24, 32, seed=5)
    depth = [x.clone().requires_grad_() for x in d["depth"]]
    full = O.photometric_loss(depth, d["pose"], d["K"], d["tgt"], d["srcs"])
    full.backward()
    assert abs(glob - full.item()) < 1e-6
    assert torch.allclose(g, depth[0].grad, rtol=1e-4, atol=1e-9)


def _ddp_worker(rank, world, port, B, q):
    """sharded_loss(reduce="mean") under torch DDP: DDP AVERAGES the per-rank parameter gradients, so the local loss is
    scaled by B_g * world / B.  The "network" is one learnable log-scale on a fixed depth pyramid."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from coivo_b200.synthetic import make_triplets
    from oracle import photometric as O
    d = make_triplets(B, 24, 32, seed=6)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.log_scale = torch.nn.Parameter(torch.zeros(()))

        def forward(self, depth):
            return [x * torch.exp(self.log_scale) for x in depth]

    net = torch.nn.parallel.DistributedDataParallel(Net())
    batch = dict(d)
    batch["depth"] = net(d["depth"])
    scaled, glob = cdist.sharded_loss(O.photometric_loss, batch, reduce="mean")
    scaled.backward()                       # DDP all-reduces and divides by world
    if rank == 0:
        q.put((glob.item(), net.module.log_scale.grad.item()))
    dist.destroy_process_group()


def test_world2_sharded_loss_mean_under_ddp_matches_the_global_gradient():
    B = 3                                   # uneven shards: 2 + 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    glob, g = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from coivo_b200.synthetic import make_triplets
    from oracle import photometric as O
    d = make_triplets(B, 24, 32, seed=6)
    ls = torch.zeros((), requires_grad=True)
    full = O.photometric_loss([x * torch.exp(ls) for x in d["depth"]], d["pose"], d["K"], d["tgt"], d["srcs"])
    full.backward()
    assert abs(glob - full.item()) < 1e-6
    assert abs(g - ls.grad.item()) <= 1e-4 * abs(ls.grad.item()) + 1e-9
