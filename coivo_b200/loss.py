"""`photometric_loss(depth, pose, K, tgt, srcs)` -- the drop-in operator of the path.

ColVO couples depth and pose through a view-synthesis loss ("loss function constraints ...
alignment of geometric projections between consecutive frames", /root/reference/README.md:7)
with the LCC brightness recalibration of the adjacent frames (README.md:5,7).  Upstream ships
no code, so the operator interface is the one BASELINE.json's north_star prescribes and
SURVEY.md section 8(b) spells out; `oracle/photometric.py` has the same signature and is the
parity reference.

A `torch.autograd.Function` over the C ABI in include/colvo.h: PyTorch only provides device
memory, the current stream and autograd plumbing.  CUDA-only: CPU tensors are rejected.
"""
from __future__ import annotations

import collections
import ctypes
import functools
from typing import List, Optional, Sequence

import torch

from . import _lib


def pack_images(x: torch.Tensor) -> torch.Tensor:
    """`[...,3,H,W]` float -> `[...,H,W,4]` bfloat16, RGBA-interleaved (A = 0): the packed image storage of
    SURVEY.md section 8(f)-3.  8 bytes per pixel instead of 12; one load per bilinear tap instead of three.
    The loss computes in fp32 on the widened values, so parity holds against the oracle run on
    `unpack_images(pack_images(x))`."""
    if x.shape[-3] != 3:
        raise ValueError("expected [...,3,H,W]")
    pad = torch.zeros_like(x[..., :1, :, :])
    return torch.cat([x, pad], dim=-3).movedim(-3, -1).contiguous().to(torch.bfloat16)


def unpack_images(p: torch.Tensor) -> torch.Tensor:
    """Inverse of `pack_images` (exact): `[...,H,W,4]` bfloat16 -> `[...,3,H,W]` float32."""
    if p.shape[-1] != 4 or p.dtype != torch.bfloat16:
        raise ValueError("expected [...,H,W,4] bfloat16")
    return p[..., :3].to(torch.float32).movedim(-1, -3).contiguous()


def _check_inputs(depth, pose, K, tgt, srcs):
    """Returns (B, N, S, H, W, packed)."""
    if not isinstance(depth, (list, tuple)) or not 1 <= len(depth) <= _lib.MAX_SCALES:
        raise ValueError(f"depth must be a sequence of 1..{_lib.MAX_SCALES} tensors [B,1,h_k,w_k]")
    tensors = list(depth) + [pose, K, tgt, srcs]
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError("all inputs must be torch.Tensor")
    packed = tgt.dtype == torch.bfloat16
    if packed:
        if tgt.dim() != 4 or tgt.shape[-1] != 4:
            raise ValueError("packed tgt must be [B,H,W,4] bfloat16 (see pack_images)")
        B, H, W, _ = tgt.shape
        if srcs.dtype != torch.bfloat16 or srcs.dim() != 5 or srcs.shape[0] != B or tuple(srcs.shape[2:]) != (H, W, 4):
            raise ValueError("packed srcs must be [B,N,H,W,4] bfloat16")
        if srcs.requires_grad or tgt.requires_grad:
            raise ValueError("packed bf16 images are data: they cannot require grad")
    else:
        if tgt.dim() != 4 or tgt.shape[1] != 3:
            raise ValueError("tgt must be [B,3,H,W]")
        B, _, H, W = tgt.shape
        if srcs.dim() != 5 or srcs.shape[0] != B or tuple(srcs.shape[2:]) != (3, H, W):
            raise ValueError("srcs must be [B,N,3,H,W]")
    N = srcs.shape[1]
    if not 1 <= N <= _lib.MAX_SOURCES:
        raise ValueError(f"1 <= N <= {_lib.MAX_SOURCES} source frames are supported")
    if tuple(pose.shape) != (B, N, 4, 4):
        raise ValueError("pose must be [B,N,4,4]")
    if tuple(K.shape) != (B, 3, 3):
        raise ValueError("K must be [B,3,3]")
    S = len(depth)
    for k, d in enumerate(depth):
        if tuple(d.shape) != (B, 1, H >> k, W >> k):
            raise ValueError(f"depth[{k}] must be [B,1,{H >> k},{W >> k}], got {tuple(d.shape)}")
    for t in list(depth) + [pose, K] + ([] if packed else [tgt, srcs]):
        if t.dtype != torch.float32:
            raise TypeError("all inputs must be float32 (images may also be packed bfloat16, see pack_images)")
    dev = tgt.device
    for t in tensors:
        if t.device.type != "cuda":
            raise ValueError("photometric_loss is CUDA-only (no CPU fallback): move inputs to a B200")
        if t.device != dev:
            raise ValueError("all inputs must live on the same device")
        if not t.is_contiguous():
            raise ValueError("inputs must be contiguous (NCHW); call .contiguous() outside the timed path")
    return B, N, S, H, W, packed


class _Workspace:
    """One scratch buffer per (device, stream), grown on demand; never shared across streams.  At most `MAX_STREAMS`
    buffers are kept (least recently used first out), `clear()` drops them all."""

    MAX_STREAMS = 4
    _bufs = collections.OrderedDict()

    @classmethod
    def get(cls, nbytes: int, device: torch.device) -> torch.Tensor:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        cls._bufs.move_to_end(key)
        while len(cls._bufs) > cls.MAX_STREAMS:
            cls._bufs.popitem(last=False)
        return buf

    @classmethod
    def clear(cls) -> None:
        cls._bufs.clear()


@functools.lru_cache(maxsize=64)
def _plan(B, N, S, H, W, flags, alpha, smooth_weight, geo_weight):
    """Descriptor and buffer sizes of one problem shape (cached: two ctypes round trips per shape, not per call)."""
    lib = _lib.load()
    desc = _lib.make_desc(B, N, S, H, W, flags, alpha, smooth_weight, geo_weight)
    nbytes, nsaved = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(lib.colvo_workspace_bytes(ctypes.byref(desc), ctypes.byref(nbytes)), "colvo_workspace_bytes")
    _lib.check(lib.colvo_saved_doubles(ctypes.byref(desc), ctypes.byref(nsaved)), "colvo_saved_doubles")
    return desc, nbytes.value, nsaved.value


class _PhotoLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, K, tgt, srcs, alpha, smooth_weight, lcc, lcc_detach, want_valid, src_depth, geo_weight, *depth):
        merge = bool(want_valid & 16)         # bit 4: warp-aggregated scatter in the backward (COLVO_F_SCATTER_MERGE)
        want_valid &= 15
        want_occ = want_valid == 2            # 2: masks and the soft occlusion mask of the geometric term
        want_valid = bool(want_valid)
        B, N, S, H, W, packed = _check_inputs(depth, pose, K, tgt, srcs)
        if src_depth is not None:
            if tuple(src_depth.shape) != (B, N, 1, H, W):
                raise ValueError("src_depth must be [B,N,1,H,W]")
            if src_depth.dtype != torch.float32:
                raise TypeError("src_depth must be float32")
            if src_depth.device != tgt.device or not src_depth.is_contiguous():
                raise ValueError("src_depth must be a contiguous tensor on the inputs' device")
        else:
            geo_weight = 0.0
        if geo_weight == 0.0:
            src_depth = None          # the term is off: the source depth maps take no part (and get no gradient)
        if want_occ and src_depth is None:
            raise ValueError("return_occlusion needs src_depth and geo_weight != 0 (the mask is a by-product of that term)")
        if K.requires_grad or tgt.requires_grad:
            raise ValueError("photometric_loss gives no gradient to K or tgt (oracle A14): detach them")
        lib = _lib.load()
        dev = tgt.device
        needs_grad = any(ctx.needs_input_grad[i] for i in (0, 3, 9)) or any(ctx.needs_input_grad[11:])
        flags = (_lib.F_LCC if lcc else 0) | (_lib.F_LCC_DETACH if lcc_detach else 0) | (_lib.F_PACKED_BF16 if packed else 0)
        if needs_grad:
            flags |= _lib.F_SAVE_FOR_BWD
        if merge:
            flags |= _lib.F_SCATTER_MERGE
        desc, nbytes, nsaved = _plan(B, N, S, H, W, flags, alpha, smooth_weight, geo_weight)
        with torch.cuda.device(dev):
            ws = _Workspace.get(nbytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ab = torch.empty(B, N, S, 2, dtype=torch.float32, device=dev)
            sel = torch.empty(B, S, H, W, dtype=torch.uint8, device=dev)
            # forward -> backward state: only when a backward can follow (validation / no_grad allocates nothing)
            saved = torch.empty(nsaved, dtype=torch.float64, device=dev) if needs_grad else None
            valid = torch.empty(B, N, S, H, W, dtype=torch.uint8, device=dev) if want_valid else None
            occ = torch.empty(B, N, S, H, W, dtype=torch.float32, device=dev) if want_occ else None
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.colvo_photo_forward_occ(
                ctypes.byref(desc), tgt.data_ptr(), srcs.data_ptr(), _lib.ptr_array([d.data_ptr() for d in depth]),
                K.data_ptr(), pose.data_ptr(), src_depth.data_ptr() if src_depth is not None else None, loss.data_ptr(),
                ab.data_ptr(), valid.data_ptr() if valid is not None else None, sel.data_ptr(),
                occ.data_ptr() if occ is not None else None,
                saved.data_ptr() if saved is not None else None, ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "colvo_photo_forward")
        ctx.desc_args = (B, N, S, H, W, flags, alpha, smooth_weight, geo_weight)
        ctx.n_depth = S
        ctx.has_src_depth = src_depth is not None
        if needs_grad:
            extra = (src_depth,) if src_depth is not None else ()
            ctx.save_for_backward(pose, K, tgt, srcs, sel, saved, *extra, *depth)
        ctx.mark_non_differentiable(ab, sel)
        ctx.set_materialize_grads(False)      # no zero-filled "gradients" for the mask outputs (a 4 MB fill per step)
        if occ is not None:
            ctx.mark_non_differentiable(valid, occ)
            return loss, ab, sel, valid, occ
        if valid is not None:
            ctx.mark_non_differentiable(valid)
            return loss, ab, sel, valid
        return loss, ab, sel

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, *unused):
        if grad_loss is None:                 # the loss itself was not used
            return (None,) * (11 + ctx.n_depth)
        pose, K, tgt, srcs, sel, saved = ctx.saved_tensors[:6]
        src_depth = ctx.saved_tensors[6] if ctx.has_src_depth else None
        depth = ctx.saved_tensors[7:] if ctx.has_src_depth else ctx.saved_tensors[6:]
        B, N, S, H, W, flags, alpha, smooth_weight, geo_weight = ctx.desc_args
        want_src = ctx.needs_input_grad[3] and not (flags & _lib.F_PACKED_BF16)
        if not want_src:
            flags |= _lib.F_NO_SRC_GRAD
        lib = _lib.load()
        dev = tgt.device
        desc, nbytes, _ = _plan(B, N, S, H, W, flags, alpha, smooth_weight, geo_weight)
        with torch.cuda.device(dev):
            ws = _Workspace.get(nbytes, dev)
            go = grad_loss.to(torch.float32).contiguous()
            grad_depth = [torch.empty_like(d) for d in depth]
            grad_T = torch.empty_like(pose)
            grad_srcs = torch.empty_like(srcs) if want_src else None
            want_sd = src_depth is not None and ctx.needs_input_grad[9]
            grad_sd = torch.empty_like(src_depth) if want_sd else None
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.colvo_photo_backward(
                ctypes.byref(desc), tgt.data_ptr(), srcs.data_ptr(), _lib.ptr_array([d.data_ptr() for d in depth]),
                K.data_ptr(), pose.data_ptr(), src_depth.data_ptr() if src_depth is not None else None, go.data_ptr(),
                sel.data_ptr(), saved.data_ptr(), _lib.ptr_array([g.data_ptr() for g in grad_depth]), grad_T.data_ptr(),
                grad_srcs.data_ptr() if want_src else None, grad_sd.data_ptr() if want_sd else None, ws.data_ptr(),
                ws.numel(), stream)
        _lib.check(rc, "colvo_photo_backward")
        gd = [g if ctx.needs_input_grad[11 + i] else None for i, g in enumerate(grad_depth)]
        return (grad_T if ctx.needs_input_grad[0] else None, None, None, grad_srcs, None, None, None, None, None,
                grad_sd, None, *gd)


def photometric_loss(
    depth: Sequence[torch.Tensor],
    pose: torch.Tensor,
    K: torch.Tensor,
    tgt: torch.Tensor,
    srcs: torch.Tensor,
    *,
    alpha: float = 0.85,
    smooth_weight: float = 1e-3,
    lcc: bool = True,
    lcc_detach: bool = False,
    return_masks: bool = False,
    src_depth: Optional[torch.Tensor] = None,
    geo_weight: float = 0.0,
    return_occlusion: bool = False,
    scatter: str = "atomic",
):
    """View-synthesis photometric loss with LCC, min-reprojection / auto-mask and edge-aware
    smoothness over S scales and N neighbouring frames.

    depth: S tensors `[B,1,H>>k,W>>k]`; pose `[B,N,4,4]` (T target->source); K `[B,3,3]`;
    tgt `[B,3,H,W]`; srcs `[B,N,3,H,W]`.  All CUDA, fp32, contiguous.  Differentiable in
    depth, pose and srcs; K and tgt get no gradient (oracle A14) and must not require grad.
    First-order only (`once_differentiable`: a double backward raises instead of returning zeros).

    `src_depth [B,N,1,H,W]` (the depth maps of the source frames) with `geo_weight > 0` adds the
    geometric-consistency term of SURVEY.md section 8(f)-2 (oracle A16); it is differentiable too.
    With `geo_weight == 0` the term is off and `src_depth` is ignored (its gradient is `None`).

    Returns the 0-dim loss, or `(loss, valid u8 [B,N,S,H,W], sel u8 [B,S,H,W], ab [B,N,S,2])`; with
    `return_occlusion=True` (needs the geometric term) a fifth item `occ [B,N,S,H,W]`: SC-Depth's soft occlusion mask
    `1 - diff` (0 where the projection is invalid), a constant by-product of that term.

    `scatter="merged"` makes the backward sum coincident bilinear taps of neighbouring pixels inside the warp before they go
    to `grad_srcs` (half the global reductions, no contended atomics; same results, measured slower on B200 than the
    default `"atomic"` vector-RED scatter -- DESIGN.md section 4).
    """
    if scatter not in ("atomic", "merged"):
        raise ValueError("scatter must be 'atomic' or 'merged'")
    mode = (2 if return_occlusion else int(bool(return_masks))) | (16 if scatter == "merged" else 0)
    out = _PhotoLossFn.apply(pose, K, tgt, srcs, float(alpha), float(smooth_weight), bool(lcc), bool(lcc_detach),
                             mode, src_depth, float(geo_weight), *depth)
    if return_occlusion:
        loss, ab, sel, valid, occ = out
        return loss, valid, sel, ab, occ
    if return_masks:
        loss, ab, sel, valid = out
        return loss, valid, sel, ab
    return out[0]


class HostStepper:
    """End-to-end step on pinned HOST buffers through `colvo_photo_step_host`: H2D of the inputs,
    forward, backward, D2H of the loss and gradients.  This is the call `bench.py` times for the
    `e2e` number.

    The batch is split into `chunks` contiguous sub-batches, each on its own CUDA stream with its
    own device arena, so the H2D copy of one chunk overlaps the kernels of another and the D2H
    copy of a third (the three engines of the GPU work concurrently); gradients are those of the
    whole-batch mean loss (each chunk runs with grad_loss = B_chunk / B).

    `grads="host"` copies every gradient back to pinned host buffers; `grads="device"` leaves them in
    the device arenas (what a training step does: the depth / pose networks consume them on the GPU)
    and reads back the loss only -- `device_grads()` returns views of them.

    `images="u8"`: `tgt` / `srcs` are uint8 frames (what a video loader holds); they cross PCIe as bytes -- a quarter of
    the image traffic of the H2D-bound step -- and are widened on the device, `x = u8 * (1 / 255)` in fp32."""

    def __init__(self, B, N, S, H, W, *, device="cuda:0", lcc=True, lcc_detach=False, want_src_grad=True,
                 alpha=0.85, smooth_weight=1e-3, chunks=None, grads="host", images="f32"):
        if grads not in ("host", "device"):
            raise ValueError("grads must be 'host' or 'device'")
        if images not in ("f32", "u8"):
            raise ValueError("images must be 'f32' or 'u8'")
        self.images = images
        if chunks is None:       # measured on B200 / PCIe 5 (scripts/e2e_chunks.py): two streams keep the H2D engine busy
            chunks = 3 if grads == "host" else 2     # when only the loss comes back, three also overlap the gradient D2H
        self.grads = grads
        self.N, self.H, self.W = N, H, W
        self.lib = _lib.load()
        self.device = torch.device(device)
        flags = (_lib.F_LCC if lcc else 0) | (_lib.F_LCC_DETACH if lcc_detach else 0)
        if not want_src_grad:
            flags |= _lib.F_NO_SRC_GRAD
        if images == "u8":         # frames cross PCIe as bytes and are widened on the device: x = u8 * (1 / 255)
            flags |= _lib.F_HOST_U8
        chunks = max(1, min(int(chunks), B))
        q, r = divmod(B, chunks)
        sizes = [q + (1 if i < r else 0) for i in range(chunks)]
        self.spans, lo = [], 0
        for n in sizes:
            self.spans.append((lo, lo + n))
            lo += n
        self.B, self.S = B, S
        pin = dict(dtype=torch.float32, pin_memory=True)
        self.h_loss_parts = torch.zeros(chunks, **pin)
        self.h_grad_depth = [torch.zeros(B, 1, H >> k, W >> k, **pin) for k in range(S)]
        self.h_grad_T = torch.zeros(B, N, 4, 4, **pin)
        self.h_grad_srcs = torch.zeros(B, N, 3, H, W, **pin) if want_src_grad else None
        self.descs, self.arenas, self.streams = [], [], []
        with torch.cuda.device(self.device):
            for n in sizes:
                d = _lib.make_desc(n, N, S, H, W, flags, alpha, smooth_weight)
                nb = ctypes.c_size_t()
                _lib.check(self.lib.colvo_step_host_arena_bytes(ctypes.byref(d), ctypes.byref(nb)), "arena_bytes")
                self.descs.append(d)
                self.arenas.append(torch.empty(nb.value, dtype=torch.uint8, device=self.device))
                self.streams.append(torch.cuda.Stream(self.device))
        self._done = [torch.cuda.Event() for _ in sizes]

    def h2d_bytes(self, depth, pose, K, tgt, srcs) -> int:
        return (tgt.numel() * tgt.element_size() + srcs.numel() * srcs.element_size()
                + 4 * (sum(d.numel() for d in depth) + K.numel() + pose.numel()))

    def d2h_bytes(self) -> int:
        n = len(self.spans)
        if self.grads == "host":
            n += sum(g.numel() for g in self.h_grad_depth) + self.h_grad_T.numel()
            if self.h_grad_srcs is not None:
                n += self.h_grad_srcs.numel()
        return 4 * n

    def device_grads(self):
        """Per chunk: `(grad_depth[S], grad_T, grad_srcs)` as views into that chunk's device arena (valid after the
        chunk's stream has finished; overwritten by the next `step`)."""
        out = []
        for i, (lo, hi) in enumerate(self.spans):
            n = hi - lo
            offs = (ctypes.c_size_t * _lib.MAX_SCALES)()
            oT, oS = ctypes.c_size_t(), ctypes.c_size_t()
            _lib.check(self.lib.colvo_step_host_arena_grads(ctypes.byref(self.descs[i]), offs, ctypes.byref(oT), ctypes.byref(oS)),
                       "colvo_step_host_arena_grads")
            a = self.arenas[i]

            def view(off, shape):
                cnt = 1
                for s_ in shape:
                    cnt *= s_
                return a[off:off + 4 * cnt].view(torch.float32).view(*shape)

            gd = [view(offs[k], (n, 1, self.H >> k, self.W >> k)) for k in range(self.S)]
            gs = view(oS.value, (n, self.N, 3, self.H, self.W)) if self.h_grad_srcs is not None else None
            out.append((gd, view(oT.value, (n, self.N, 4, 4)), gs))
        return out

    def step(self, depth, pose, K, tgt, srcs):
        """Inputs are CPU tensors (pinned for full speed).  Enqueues every chunk on its own stream, ordered
        after the caller's current stream; call `finish()` (or synchronise) before reading the host outputs."""
        img_dtype = torch.uint8 if self.images == "u8" else torch.float32
        for t, dt in [(x, torch.float32) for x in list(depth) + [pose, K]] + [(tgt, img_dtype), (srcs, img_dtype)]:
            if t.device.type != "cpu" or t.dtype != dt or not t.is_contiguous():
                raise ValueError("HostStepper.step takes contiguous CPU tensors: float32, and uint8 frames with images='u8'")
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            for i, (lo, hi) in enumerate(self.spans):
                st = self.streams[i]
                st.wait_stream(cur)
                to_host = self.grads == "host"
                rc = self.lib.colvo_photo_step_host(
                    ctypes.byref(self.descs[i]), tgt[lo:hi].data_ptr(), srcs[lo:hi].data_ptr(),
                    _lib.ptr_array([d[lo:hi].data_ptr() for d in depth]), K[lo:hi].data_ptr(), pose[lo:hi].data_ptr(),
                    self.h_loss_parts[i:i + 1].data_ptr(),
                    _lib.ptr_array([g[lo:hi].data_ptr() for g in self.h_grad_depth]) if to_host else None,
                    self.h_grad_T[lo:hi].data_ptr() if to_host else None,
                    self.h_grad_srcs[lo:hi].data_ptr() if (to_host and self.h_grad_srcs is not None) else None,
                    ctypes.c_float((hi - lo) / self.B), self.arenas[i].data_ptr(), self.arenas[i].numel(), st.cuda_stream)
                _lib.check(rc, "colvo_photo_step_host")
                self._done[i].record(st)
        return self

    def join(self):
        """Make the caller's current stream wait for every chunk (for event timing); no host sync."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self._done:
            cur.wait_event(ev)

    def finish(self) -> torch.Tensor:
        """Synchronise all chunk streams and return the whole-batch mean loss (host tensor)."""
        for st in self.streams:
            st.synchronize()
        w = torch.tensor([(hi - lo) / self.B for lo, hi in self.spans])
        return (self.h_loss_parts * w).sum()
