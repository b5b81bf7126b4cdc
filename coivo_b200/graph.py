"""CUDA-graph capture of one forward+backward step of the path.

The C ABI only enqueues work on the caller's stream (no allocation, no synchronisation), so a
whole `photometric_loss(...).backward()` step captures into one CUDA graph: nine kernel launches
(with their programmatic-dependent-launch edges) and the autograd bookkeeping collapse into a single
`cudaGraphLaunch`.  Inputs are static
tensors: copy new data into them (`tensor.copy_`) between replays.
"""
from __future__ import annotations

from typing import Sequence

import torch

from .loss import photometric_loss


class GraphedStep:
    """Capture `loss = photometric_loss(...); loss.backward()` once, replay many times.

    After `replay()`, `self.loss` holds the loss and the `.grad` of every input that requires grad
    holds its gradient (overwritten by each replay)."""

    def __init__(self, depth: Sequence[torch.Tensor], pose, K, tgt, srcs, warmup: int = 3, **kw):
        self.depth, self.pose, self.K, self.tgt, self.srcs = list(depth), pose, K, tgt, srcs
        self.kw = kw
        self._leaves = [t for t in self.depth + [pose, srcs] if t.requires_grad]
        dev = tgt.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for t in self._leaves:
            t.grad = None
        self._one = torch.ones((), dtype=torch.float32, device=dev)     # d loss / d loss, made once (no fill in the graph)
        self.graph = torch.cuda.CUDAGraph()
        # the leaves' AccumulateGrad nodes were created on the warm-up stream and capture runs on its own stream: the
        # stream-mismatch warning is silenced for the capture only and put back afterwards
        setter = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if setter is not None:
            setter(False)
        try:
            with torch.cuda.graph(self.graph):
                self.loss = photometric_loss(self.depth, self.pose, self.K, self.tgt, self.srcs, **self.kw)
                self.loss.backward(gradient=self._one)
        finally:
            if setter is not None:
                setter(True)

    def _eager(self):
        for t in self._leaves:
            t.grad = None
        loss = photometric_loss(self.depth, self.pose, self.K, self.tgt, self.srcs, **self.kw)
        loss.backward()
        return loss

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
