"""Batch sharding of the path across GPUs.

Every triplet's loss and gradients depend on that triplet only, so the path shards by the
batch axis with no data-path collective: rank g owns triplets [lo, hi), computes its local
mean loss and local gradients; the only exchange is one scalar all-reduce for the global loss
(the bulk gradient all-reduce of a training loop belongs to the out-of-scope CNNs).
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of B triplets: the first B % world ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    q, r = divmod(B, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_batch(batch: Dict[str, object], rank: int, world: int) -> Dict[str, object]:
    B = batch["tgt"].shape[0]
    lo, hi = shard_range(B, rank, world)
    out = {}
    for k, v in batch.items():
        out[k] = [x[lo:hi].contiguous() for x in v] if isinstance(v, (list, tuple)) else v[lo:hi].contiguous()
    return out


def global_mean_loss(local_loss: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """Mean over the global batch from per-rank means (weights = shard sizes): one tiny
    all-reduce of (loss * count, count).  Works on NCCL (CUDA tensors) and gloo (CPU)."""
    buf = torch.stack([local_loss.detach().to(torch.float64) * local_count,
                       torch.tensor(float(local_count), dtype=torch.float64, device=local_loss.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(local_loss.dtype)


def sharded_loss(loss_fn: Callable[..., torch.Tensor], batch: Dict[str, object], *, group=None, reduce: str = "sum", **kw):
    """Run `loss_fn(depth, pose, K, tgt, srcs)` on this rank's shard and return
    `(local loss to back-propagate, global mean loss)`.  `loss_fn` is `coivo_b200.photometric_loss` in production.

    The scale of the returned local loss depends on how the caller combines gradients across ranks:

    * `reduce="sum"` (default): `local * (B_g / B)`.  SUMMING the per-rank gradients (all-reduce with
      `ReduceOp.SUM`, or simply keeping per-sample gradients local) gives the gradient of the global mean.
    * `reduce="mean"`: `local * (B_g * world / B)`.  AVERAGING the per-rank gradients gives the gradient of the
      global mean -- this is what `torch.nn.parallel.DistributedDataParallel` does (it divides the summed
      gradients by `world_size`), so use it when the depth / pose networks are wrapped in DDP.  For balanced
      shards the factor is exactly 1.
    """
    if reduce not in ("sum", "mean"):
        raise ValueError("reduce must be 'sum' or 'mean'")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = batch["tgt"].shape[0]
    sh = shard_batch(batch, rank, world)
    nloc = sh["tgt"].shape[0]
    local = loss_fn(sh["depth"], sh["pose"], sh["K"], sh["tgt"], sh["srcs"], **kw)
    glob = global_mean_loss(local, nloc, group)
    scale = nloc / B if reduce == "sum" else nloc * world / B
    return local * scale, glob
