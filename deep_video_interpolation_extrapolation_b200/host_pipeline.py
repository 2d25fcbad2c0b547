"""warp_blend for HOST-resident clips: batch-chunked H2D -> forward+backward -> D2H pipeline on three CUDA streams.

The reference's data loader hands the trainer pinned host batches (`folder.py`, `DataLoader(pin_memory=True)`) and the
runners call `.cuda()` on them before the warp (`runners/InterTrainer.py:399-408`).  When the inputs and the wanted
results both live on the host, the op is PCIe-bound (about 300 B/pixel in, 300 B/pixel out against 808 B/pixel of HBM
traffic on a link 100x slower), so the only thing that matters is to keep both directions of the link busy at once:
the batch is cut into chunks along N (the op has no cross-sample term, utils/net_utils.py:93-114) and chunk k+1 is
uploaded while chunk k computes and chunk k-1 is downloaded.

All arithmetic runs in libflowwarp_b200.so (the three C-ABI calls the autograd op makes, on preallocated device slots);
there is no CPU path here either.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import ctypes

import torch

from . import _lib as L
from ._problem import fill_grads, fill_problem
from .sharding import chunk_slices

Tensor = torch.Tensor


def _numel(shape) -> int:
    n = 1
    for x in shape:
        n *= int(x)
    return n


def chunk_layout(n: int, Cs: Sequence[int], H: int, W: int):
    """Shapes, in arena order, of the inputs and of the results of one chunk of n clips:
    inputs  frames0[g].., frames1[g].., for_flow, back_flow, for_mask, back_mask, grad_outs[g]..
    results outs[g].., grad_frames0[g].., grad_frames1[g].., grad_for_flow, grad_back_flow, grad_for_mask, grad_back_mask"""
    fr = [(n, c, H, W) for c in Cs]
    fl, mk = [(n, 2, H, W)] * 2, [(n, 1, H, W)] * 2
    return [*fr, *fr, *fl, *mk, *fr], [*fr, *fr, *fr, *fl, *mk]


def _carve(buf: Tensor, shapes) -> List[Tensor]:
    out, off = [], 0
    for sh in shapes:
        k = _numel(sh)
        out.append(buf[off:off + k].view(*sh))
        off += k
    return out


class _Slot:
    """Device buffers of one in-flight chunk (inputs, outputs, gradients, workspace) and its C-ABI structs.  Preallocated
    once: the timed path neither allocates nor frees (no caching-allocator traffic between the three streams)."""

    def __init__(self, dev, n, Cs, H, W, kw):
        f32 = dict(dtype=torch.float32, device=dev)
        self.n = n
        # every input of a chunk lives in ONE device buffer and every result in another, in the order and at the offsets of
        # the host arenas (chunk_layout): a chunk moves with one cudaMemcpyAsync each way
        in_shapes, out_shapes = chunk_layout(n, Cs, H, W)
        self.in_buf = torch.empty(sum(_numel(sh) for sh in in_shapes), **f32)
        self.out_buf = torch.empty(sum(_numel(sh) for sh in out_shapes), **f32)
        iv, ov = _carve(self.in_buf, in_shapes), _carve(self.out_buf, out_shapes)
        G = len(Cs)
        u5 = lambda t: t.unsqueeze(1)  # noqa: E731  [n,C,H,W] -> [n,1,C,H,W]
        self.f0, self.f1 = [u5(t) for t in iv[:G]], [u5(t) for t in iv[G:2 * G]]
        self.flows = [t.unsqueeze(2) for t in iv[2 * G:2 * G + 2]]
        self.masks = list(iv[2 * G + 2:2 * G + 4])
        self.gos = [u5(t) for t in iv[2 * G + 4:]]
        self.outs = [u5(t) for t in ov[:G]]
        self.g_f0, self.g_f1 = [u5(t) for t in ov[G:2 * G]], [u5(t) for t in ov[2 * G:3 * G]]
        self.g_flows = [t.unsqueeze(2) for t in ov[3 * G:3 * G + 2]]
        self.g_masks = list(ov[3 * G + 2:3 * G + 4])
        self.ev_cmp = torch.cuda.Event()  # compute of the chunk that last used this slot is done (inputs free)
        self.ev_out = torch.cuda.Event()  # its results have left the device (outputs / gradients free)
        self.used = False
        self.lib = L.load()
        pad = L.FWB_PAD_BORDER if kw["padding_mode"] == "border" else L.FWB_PAD_ZEROS
        self.fused = not kw["deterministic"]  # fused backward: the forward zero-fills grad_src (no memset pass)
        flags = L.FWB_FLAG_DETERMINISTIC if kw["deterministic"] else (L.FWB_FLAG_FUSED_BWD | L.FWB_FLAG_GRAD_SRC_ZEROED)
        ptr, st = (lambda t: t.data_ptr()), (lambda t: t.stride())
        self.structs = {}
        for m in sorted({n} | ({kw["tail"]} if kw.get("tail") else set())):  # full chunk and the ragged last chunk
            v = lambda ts: [t[:m] for t in ts]  # noqa: E731
            p = fill_problem(N=m, T=1, H=H, W=W, flows=v(self.flows), gates=[None, None], blends=v(self.masks),
                             signs=[-1.0, 1.0], srcs=[[a[:m], b[:m]] for a, b in zip(self.f0, self.f1)], outs=v(self.outs),
                             padding_mode=pad, align_corners=kw["align_corners"], flags=flags, ptr=ptr, strides=st)
            q = fill_grads(p, grad_outs=v(self.gos), grad_srcs=[[a[:m], b[:m]] for a, b in zip(self.g_f0, self.g_f1)],
                           grad_flows=v(self.g_flows), grad_gates=[None, None], grad_blends=v(self.g_masks), ptr=ptr, strides=st)
            self.structs[m] = (p, q)
        self.ws_bytes = max(int(self.lib.fwb_workspace_bytes(ctypes.byref(self.structs[n][0]))), 1)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def inputs(self):
        return [*self.f0, *self.f1, *self.flows, *self.masks, *self.gos]

    def results(self):
        return [*self.outs, *self.g_f0, *self.g_f1, *self.g_flows, *self.g_masks]

    def launch(self, m, stream_ptr):
        p, q = self.structs[m]
        lib = self.lib
        if self.fused:
            L.check(lib.fwb_warp_blend_forward_zero(ctypes.byref(p), ctypes.byref(q), stream_ptr), "fwb_warp_blend_forward_zero")
        else:
            L.check(lib.fwb_warp_blend_forward(ctypes.byref(p), stream_ptr), "fwb_warp_blend_forward")
        L.check(lib.fwb_warp_blend_backward_flow(ctypes.byref(p), ctypes.byref(q), self.ws.data_ptr(), self.ws_bytes, stream_ptr),
                "fwb_warp_blend_backward_flow")
        L.check(lib.fwb_warp_blend_backward_src(ctypes.byref(p), ctypes.byref(q), self.ws.data_ptr(), self.ws_bytes, stream_ptr),
                "fwb_warp_blend_backward_src")


class HostWarpBlend:
    """Reusable pipeline object: streams, device slots and pinned result buffers are kept between calls.

    run(frames0, frames1, for_flow, back_flow, for_mask, back_mask, grad_outs) with every argument a (pinned) host
    tensor laid out as `warp_blend` expects ([N,C,H,W] frames and grad_outs, [N,2,H,W] flows, [N,1,H,W] masks);
    returns a dict of pinned host tensors:
        outs[g], grad_frames0[g], grad_frames1[g], grad_for_flow, grad_back_flow, grad_for_mask, grad_back_mask
    i.e. exactly what `warp_blend(...)` followed by `torch.autograd.backward(outs, grad_outs)` produces, computed by the
    same three C-ABI calls the autograd op makes, on preallocated device slots.  The returned buffers are reused by the
    next call.  `synchronize=False` leaves the last copies in flight (the caller's current stream is made to wait for
    them; `self.done` is recorded behind the last copy).
    """

    SLOTS = 3

    def __init__(self, device, chunk: int = 2, padding_mode: str = "border", align_corners: bool = False,
                 deterministic: bool = False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostWarpBlend: a CUDA device is required (this library has no CPU path)")
        if chunk < 1:
            raise ValueError("chunk must be >= 1")
        if padding_mode not in ("zeros", "border"):
            raise ValueError(f"padding_mode must be 'zeros' or 'border', got {padding_mode!r}")
        L.load()  # fail loudly when the CUDA library is missing
        self.chunk = int(chunk)
        self.kw = dict(padding_mode=padding_mode, align_corners=bool(align_corners), deterministic=bool(deterministic))
        self.s_in = torch.cuda.Stream(self.device)
        self.s_cmp = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.done = torch.cuda.Event()
        self._res: Optional[Dict[str, object]] = None
        self._slots: List[_Slot] = []
        self._key = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._arena_in: Optional[Tensor] = None   # [chunks, floats per chunk] pinned: all inputs of a chunk back to back
        self._arena_out: Optional[Tensor] = None  # [chunks, floats per chunk] pinned: all results of a chunk back to back
        self._arena_key = None

    def _setup(self, key, N, Cs, H, W):
        if self._key == key:
            return
        mk = lambda *s: torch.empty(s, dtype=torch.float32).pin_memory()  # noqa: E731
        self._res = {
            "outs": [mk(N, c, H, W) for c in Cs],
            "grad_frames0": [mk(N, c, H, W) for c in Cs], "grad_frames1": [mk(N, c, H, W) for c in Cs],
            "grad_for_flow": mk(N, 2, H, W), "grad_back_flow": mk(N, 2, H, W),
            "grad_for_mask": mk(N, 1, H, W), "grad_back_mask": mk(N, 1, H, W),
        }
        n = min(self.chunk, max(N, 1))
        kw = dict(self.kw, tail=(N % n) or None)
        with torch.cuda.device(self.device):
            self._slots = [_Slot(self.device, n, Cs, H, W, kw) for _ in range(min(self.SLOTS, -(-max(N, 1) // n)))]
        self._key = key

    def run(self, frames0: Sequence[Tensor], frames1: Sequence[Tensor], for_flow: Tensor, back_flow: Tensor,
            for_mask: Tensor, back_mask: Tensor, grad_outs: Sequence[Tensor], synchronize: bool = True):
        G = len(frames0)
        if G < 1 or len(frames1) != G or len(grad_outs) != G:
            raise ValueError("HostWarpBlend: frames0, frames1 and grad_outs need one tensor per channel group")
        host_in: List[Tensor] = [*frames0, *frames1, for_flow, back_flow, for_mask, back_mask, *grad_outs]
        for t in host_in:
            if not isinstance(t, Tensor):
                raise TypeError("HostWarpBlend: tensors expected")
            if t.is_cuda:
                raise RuntimeError("HostWarpBlend: host tensors expected (use warp_blend for device tensors)")
            if t.dtype != torch.float32:
                raise RuntimeError(f"HostWarpBlend: float32 required, got {t.dtype}")
        if for_flow.dim() != 4 or for_flow.shape[1] != 2:
            raise RuntimeError(f"flow must be [N,2,H,W], got {tuple(for_flow.shape)}")
        N, _, H, W = for_flow.shape
        Cs = tuple(int(t.shape[1]) for t in frames0)
        want = ([(N, c, H, W) for c in Cs] * 2 + [(N, 2, H, W)] * 2 + [(N, 1, H, W)] * 2 + [(N, c, H, W) for c in Cs])
        for t, w in zip(host_in, want):
            if tuple(t.shape) != w:
                raise RuntimeError(f"HostWarpBlend: expected shape {w}, got {tuple(t.shape)}")
        host_in = [(t if t.is_contiguous() else t.contiguous()) for t in host_in]
        host_in = [t if t.is_pinned() else t.pin_memory() for t in host_in]
        self._setup((N, Cs, H, W), N, Cs, H, W)
        res = self._res
        dst = self._flat(res)
        self.h2d_bytes = sum(t.numel() * 4 for t in host_in)
        self.d2h_bytes = sum(t.numel() * 4 for t in dst)
        if N == 0:
            return res
        dev = self.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            for s in (self.s_in, self.s_cmp, self.s_out):
                s.wait_stream(cur)
            for k, sl in enumerate(chunk_slices(N, self.chunk)):
                slot = self._slots[k % len(self._slots)]
                m = sl.stop - sl.start
                with torch.cuda.stream(self.s_in):
                    if slot.used:
                        self.s_in.wait_event(slot.ev_cmp)
                    for dbuf, h in zip(slot.inputs(), host_in):
                        dbuf[:m].view(h[sl].shape).copy_(h[sl], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_in)
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(ev_in)
                    if slot.used:
                        self.s_cmp.wait_event(slot.ev_out)
                    slot.launch(m, self.s_cmp.cuda_stream)
                    slot.ev_cmp.record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(slot.ev_cmp)
                    for h, dbuf in zip(dst, slot.results()):
                        h[sl].copy_(dbuf[:m].view(h[sl].shape), non_blocking=True)
                    slot.ev_out.record(self.s_out)
                slot.used = True
            self.done.record(self.s_out)
            cur.wait_stream(self.s_out)
            cur.wait_stream(self.s_cmp)
            cur.wait_stream(self.s_in)
        if synchronize:
            self.done.synchronize()
        return res

    # ------------------------------------------------------------------ arena mode: one copy per chunk and direction
    def arena(self, N: int, Cs: Sequence[int], H: int, W: int) -> Dict[str, object]:
        """Pinned host ARENAS for a batch of N clips: the producer (data loader / previous stage) writes its tensors straight
        into the returned per-chunk views and reads the results from them, so that `run_arena()` moves every chunk with ONE
        cudaMemcpyAsync host->device and ONE device->host instead of 13 + 11 (the per-tensor path of `run`).

        Returns {"inputs": [per chunk: list of views in chunk_layout order], "results": [per chunk: list of views],
                 "chunks": [slice of the batch each chunk covers]}.  The last chunk may be ragged: its views hold m < chunk clips
        (the arena row is still laid out for a full chunk)."""
        Cs = tuple(int(c) for c in Cs)
        key = (N, Cs, H, W)
        n = min(self.chunk, max(N, 1))
        sls = chunk_slices(N, n)
        if self._arena_key != key:
            self._setup(key, N, Cs, H, W)
            in_shapes, out_shapes = chunk_layout(n, Cs, H, W)
            fi, fo = sum(_numel(s) for s in in_shapes), sum(_numel(s) for s in out_shapes)
            self._arena_in = torch.empty((max(len(sls), 1), fi), dtype=torch.float32).pin_memory()
            self._arena_out = torch.empty((max(len(sls), 1), fo), dtype=torch.float32).pin_memory()
            self._arena_key = key
        in_shapes, out_shapes = chunk_layout(n, Cs, H, W)
        ins, outs = [], []
        for k, sl in enumerate(sls):
            m = sl.stop - sl.start
            ins.append([v[:m] for v in _carve(self._arena_in[k], in_shapes)])
            outs.append([v[:m] for v in _carve(self._arena_out[k], out_shapes)])
        return {"inputs": ins, "results": outs, "chunks": sls}

    def fill_arena(self, frames0, frames1, for_flow, back_flow, for_mask, back_mask, grad_outs) -> Dict[str, object]:
        """Convenience (CPU copies): place ordinary host tensors ([N,...]) into the arena.  A real producer writes the views."""
        N, _, H, W = for_flow.shape
        a = self.arena(N, [int(t.shape[1]) for t in frames0], H, W)
        host = [*frames0, *frames1, for_flow, back_flow, for_mask, back_mask, *grad_outs]
        for views, sl in zip(a["inputs"], a["chunks"]):
            for v, t in zip(views, host):
                v.copy_(t[sl])
        return a

    def run_arena(self, synchronize: bool = True):
        """forward + backward of the batch that sits in the arena (see `arena`): per chunk ONE H2D copy of its inputs, the
        three C-ABI calls, ONE D2H copy of its results; H2D | compute | D2H of consecutive chunks overlap on three streams."""
        if self._arena_key is None:
            raise RuntimeError("HostWarpBlend.run_arena: call arena(...) / fill_arena(...) first")
        N = self._arena_key[0]
        sls = chunk_slices(N, min(self.chunk, max(N, 1)))
        self.h2d_bytes = len(sls) * self._arena_in.shape[1] * 4
        self.d2h_bytes = len(sls) * self._arena_out.shape[1] * 4
        if N == 0:
            return
        dev = self.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            for s in (self.s_in, self.s_cmp, self.s_out):
                s.wait_stream(cur)
            for k, sl in enumerate(sls):
                slot = self._slots[k % len(self._slots)]
                m = sl.stop - sl.start
                with torch.cuda.stream(self.s_in):
                    if slot.used:
                        self.s_in.wait_event(slot.ev_cmp)
                    slot.in_buf.copy_(self._arena_in[k], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_in)
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(ev_in)
                    if slot.used:
                        self.s_cmp.wait_event(slot.ev_out)
                    slot.launch(m, self.s_cmp.cuda_stream)
                    slot.ev_cmp.record(self.s_cmp)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(slot.ev_cmp)
                    self._arena_out[k].copy_(slot.out_buf, non_blocking=True)
                    slot.ev_out.record(self.s_out)
                slot.used = True
            self.done.record(self.s_out)
            cur.wait_stream(self.s_out)
            cur.wait_stream(self.s_cmp)
            cur.wait_stream(self.s_in)
        if synchronize:
            self.done.synchronize()

    @staticmethod
    def _flat(res) -> List[Tensor]:
        return [*res["outs"], *res["grad_frames0"], *res["grad_frames1"], res["grad_for_flow"], res["grad_back_flow"],
                res["grad_for_mask"], res["grad_back_mask"]]


def warp_blend_host(frames0, frames1, for_flow, back_flow, for_mask, back_mask, grad_outs, device="cuda:0", chunk: int = 2,
                    **kw):
    """One-shot form of `HostWarpBlend(device, chunk, **kw).run(...)`."""
    return HostWarpBlend(device, chunk, **kw).run(frames0, frames1, for_flow, back_flow, for_mask, back_mask, grad_outs)
