"""numpy front end of the CPU oracle (oracle/flowwarp_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs as the checker; never by the product package.  It reuses the product's ctypes struct layout
(one problem description for both sides) but none of its compute.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from deep_video_interpolation_extrapolation_b200 import _lib as L
from deep_video_interpolation_extrapolation_b200._problem import fill_grads, fill_problem

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libflowwarp_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flowwarp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libflowwarp_oracle.so"], check=True, capture_output=True)
    return _SO


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = C.CDLL(_SO)
        pp, gp, vp = C.POINTER(L.fwb_problem), C.POINTER(L.fwb_grads), C.c_void_p
        lib.fwo_warp_blend_forward.argtypes = [pp]
        lib.fwo_warp_blend_backward.argtypes = [pp, gp]
        lib.fwo_sample_indices.argtypes = [pp, C.c_int32, vp, vp, vp, vp, vp]
        lib.fwo_set_num_threads.argtypes = [C.c_int32]
        lib.fwo_base_coord.argtypes = [C.c_int, C.c_int]
        lib.fwo_base_coord.restype = C.c_float
        fpp = C.POINTER(C.c_void_p)
        lib.fwo_bidir_contig.argtypes = [C.c_int32] * 4 + [C.POINTER(C.c_int32), fpp, fpp, vp, vp, vp, vp, C.c_int32, C.c_int32,
                                                           fpp, fpp, fpp, fpp, vp, vp, vp, vp]
        _lib = lib
    return _lib


def set_num_threads(k: int) -> None:
    load().fwo_set_num_threads(int(k))


def base_coord(i: int, n: int) -> float:
    return float(load().fwo_base_coord(i, n))


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _strides(a: np.ndarray):
    return tuple(s // a.itemsize for s in a.strides)


def _f32(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)
    if a.ndim and a.strides[-1] != a.itemsize:
        a = np.ascontiguousarray(a)
    return a


def _canon(srcs, flows, gates, blends):
    flows = [_f32(f) for f in flows]
    flows = [f[:, :, None] if f.ndim == 4 else f for f in flows]
    N, _, T, H, W = flows[0].shape
    D = len(flows)

    def mask(m):
        if m is None:
            return None
        m = _f32(m)
        if m.ndim == 3:
            m = m[:, None]
        if m.ndim == 5:
            m = m[:, :, 0]
        if m.shape[1] != T:
            m = np.broadcast_to(m, (N, T, H, W))
        return m

    gates = [mask(m) for m in (gates or [None] * D)]
    blends = [mask(m) for m in (blends or [None] * D)]
    groups = []
    for g in srcs:
        row = []
        for s in ([g] if isinstance(g, np.ndarray) else list(g)):
            s = _f32(s)
            if s.ndim == 4:
                s = s[:, None]
            if s.shape[1] != T:
                s = np.broadcast_to(s, (N, T) + s.shape[2:])
            row.append(s)
        groups.append(row)
    return groups, flows, gates, blends, (N, T, H, W)


def _pad(mode) -> int:
    return L.FWB_PAD_BORDER if mode == "border" else L.FWB_PAD_ZEROS


def forward(srcs, flows, gates=None, blends=None, signs=None, padding_mode="zeros", align_corners=False) -> List[np.ndarray]:
    """Same semantics as deep_video_interpolation_extrapolation_b200.ops.flow_warp_blend, numpy in/out, [N,T,C,H,W]."""
    groups, flows, gates, blends, (N, T, H, W) = _canon(srcs, flows, gates, blends)
    D = len(flows)
    signs = [-1.0] * D if signs is None else ([float(signs)] * D if np.isscalar(signs) else list(signs))
    outs = [np.empty((N, T, g[0].shape[2], H, W), np.float32) for g in groups]
    p = fill_problem(N=N, T=T, H=H, W=W, flows=flows, gates=gates, blends=blends, signs=signs, srcs=groups, outs=outs,
                     padding_mode=_pad(padding_mode), align_corners=align_corners, flags=0, ptr=_ptr, strides=_strides)
    rc = load().fwo_warp_blend_forward(C.byref(p))
    if rc:
        raise ValueError(f"oracle forward: code {rc}")
    return outs


def backward(srcs, flows, grad_outs, gates=None, blends=None, signs=None, padding_mode="zeros", align_corners=False):
    """All gradients of `forward` for upstream grad_outs[g] ([N,T,C,H,W] or [N,C,H,W]).

    Returns dict(grad_srcs=[[...per dir] per group], grad_flows=[...], grad_gates=[...], grad_blends=[...]).
    A source shared by all T frames gets a [N,1,C,H,W] gradient (summed over T)."""
    groups, flows, gates, blends, (N, T, H, W) = _canon(srcs, flows, gates, blends)
    D = len(flows)
    signs = [-1.0] * D if signs is None else ([float(signs)] * D if np.isscalar(signs) else list(signs))
    gos = []
    for g, go in zip(groups, grad_outs):
        if go is None:
            gos.append(None)
            continue
        go = _f32(go)
        gos.append(go[:, None] if go.ndim == 4 else go)
    g_srcs, g_bufs = [], []
    for g in groups:
        row, brow = [], []
        for s in g:
            shared = s.strides[1] == 0 and T > 1
            buf = np.zeros((N, 1 if shared else T, s.shape[2], H, W), np.float32)
            brow.append(buf)
            row.append(np.broadcast_to(buf, (N, T, s.shape[2], H, W)) if shared else buf)
        g_srcs.append(row)
        g_bufs.append(brow)
    g_flows = [np.zeros((N, 2, T, H, W), np.float32) for _ in range(D)]
    g_gates = [np.zeros((N, T, H, W), np.float32) if gates[d] is not None else None for d in range(D)]
    g_blends = [np.zeros((N, T, H, W), np.float32) if blends[d] is not None else None for d in range(D)]
    p = fill_problem(N=N, T=T, H=H, W=W, flows=flows, gates=gates, blends=blends, signs=signs, srcs=groups, outs=None,
                     padding_mode=_pad(padding_mode), align_corners=align_corners, flags=0, ptr=_ptr, strides=_strides)
    q = fill_grads(p, grad_outs=gos, grad_srcs=g_srcs, grad_flows=g_flows, grad_gates=g_gates, grad_blends=g_blends,
                   ptr=_ptr, strides=_strides)
    rc = load().fwo_warp_blend_backward(C.byref(p), C.byref(q))
    if rc:
        raise ValueError(f"oracle backward: code {rc}")
    return dict(grad_srcs=g_bufs, grad_flows=g_flows, grad_gates=g_gates, grad_blends=g_blends)


def sample_indices(flow, gate=None, sign=-1.0, padding_mode="zeros", align_corners=False):
    """(x0, y0, valid, ix, iy), each [N,T,H,W] — the integer taps / validity bits / float coordinates."""
    dummy = np.zeros((flow.shape[0], 1, flow.shape[-2], flow.shape[-1]), np.float32)
    groups, flows, gates, blends, (N, T, H, W) = _canon([dummy], [flow], [gate], None)
    p = fill_problem(N=N, T=T, H=H, W=W, flows=flows, gates=gates, blends=blends, signs=[sign], srcs=groups, outs=None,
                     padding_mode=_pad(padding_mode), align_corners=align_corners, flags=0, ptr=_ptr, strides=_strides)
    x0 = np.empty((N, T, H, W), np.int32)
    y0 = np.empty_like(x0)
    valid = np.empty((N, T, H, W), np.uint8)
    ix = np.empty((N, T, H, W), np.float32)
    iy = np.empty_like(ix)
    rc = load().fwo_sample_indices(C.byref(p), 0, _ptr(x0), _ptr(y0), _ptr(valid), _ptr(ix), _ptr(iy))
    if rc:
        raise ValueError(f"oracle sample_indices: code {rc}")
    return x0, y0, valid, ix, iy


# ---------------------------------------------------------------------------------------------
# `refine`'s mask blend (utils/net_utils.py:131-143), numpy, same op order as the torch expression at :141-142:
#     input[:, i] * mask[:, i:i+1] + noise * (1. - mask[:, i:i+1])
# with noise = cat([noise_bg, zeros(bs, 20, h, w)], 1) when opt.seg (:134-136).  fp32 ops, one rounding each.
# ---------------------------------------------------------------------------------------------
def mask_blend_forward(inp, mask, noise_bg=None) -> np.ndarray:
    """inp [N,T,C,H,W], mask [N,T,H,W], noise_bg [N,Cn<=C,H,W] or None -> [N,T,C,H,W] (float32)."""
    inp, mask = _f32(inp), _f32(mask)
    N, T, C, H, W = inp.shape
    noise = np.zeros((N, C, H, W), np.float32)
    if noise_bg is not None:
        nb = _f32(noise_bg)
        noise[:, :nb.shape[1]] = nb  # :134-136 (torch.cat with a zero block)
    out = np.empty_like(inp)
    one = np.float32(1.0)
    for i in range(T):  # :141 the Python loop over opt.vid_length
        m = mask[:, i:i + 1]
        out[:, i] = inp[:, i] * m + noise * (one - m)
    return out


def mask_blend_backward(inp, mask, noise_bg, grad_out):
    """float64 gradients of mask_blend_forward w.r.t. (inp, mask, noise_bg)."""
    inp, mask, g = np.asarray(inp, np.float64), np.asarray(mask, np.float64), np.asarray(grad_out, np.float64)
    N, T, C, H, W = inp.shape
    noise = np.zeros((N, C, H, W), np.float64)
    Cn = 0
    if noise_bg is not None:
        Cn = noise_bg.shape[1]
        noise[:, :Cn] = noise_bg
    gi = g * mask[:, :, None]
    gm = (g * (inp - noise[:, None])).sum(axis=2)
    gn = (g * (1.0 - mask[:, :, None])).sum(axis=1)[:, :Cn]
    return gi, gm, gn


# ---------------------------------------------------------------------------------------------
# Flat entry (full-size parity tests): contiguous arrays in, the problem description packed in C by
# fwo_bidir_contig — independent of the product's _problem.fill_problem.
# ---------------------------------------------------------------------------------------------
def bidir_contig(src0, src1, flow0, flow1, blend0=None, blend1=None, grad_outs=None, padding_mode="border", align_corners=False,
                 want_forward=True):
    """Bidirectional warp + mask-weighted blend (nets/OpticalUnet.py:123-146) on contiguous fp32 numpy arrays.

    src0[g], src1[g] [N,C,H,W]; flow0, flow1 [N,2,H,W]; blend0, blend1 [N,H,W] (or [N,1,H,W]) or None; grad_outs[g] [N,C,H,W].
    Returns dict(out=[...], gsrc0=[...], gsrc1=[...], gflow0, gflow1, gblend0, gblend1) (absent parts are None)."""
    c32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    src0, src1 = [c32(a) for a in src0], [c32(a) for a in src1]
    flow0, flow1, blend0, blend1 = c32(flow0), c32(flow1), c32(blend0), c32(blend1)
    N, _, H, W = flow0.shape
    G = len(src0)
    Cs = (C.c_int32 * G)(*[a.shape[1] for a in src0])
    arr = lambda xs: None if xs is None else (C.c_void_p * G)(*[x.ctypes.data for x in xs])
    out = [np.empty_like(a) for a in src0] if want_forward else None
    res = dict(out=out, gsrc0=None, gsrc1=None, gflow0=None, gflow1=None, gblend0=None, gblend1=None)
    gos = None
    if grad_outs is not None:
        gos = [c32(g) for g in grad_outs]
        res["gsrc0"], res["gsrc1"] = [np.zeros_like(a) for a in src0], [np.zeros_like(a) for a in src1]
        res["gflow0"], res["gflow1"] = np.zeros_like(flow0), np.zeros_like(flow1)
        if blend0 is not None:
            res["gblend0"] = np.zeros((N, H, W), np.float32)
        if blend1 is not None:
            res["gblend1"] = np.zeros((N, H, W), np.float32)
    p = lambda a: None if a is None else a.ctypes.data
    rc = load().fwo_bidir_contig(N, H, W, G, Cs, arr(src0), arr(src1), p(flow0), p(flow1), p(blend0), p(blend1), _PAD[padding_mode],
                                 int(bool(align_corners)), arr(out), arr(gos), arr(res["gsrc0"]), arr(res["gsrc1"]), p(res["gflow0"]),
                                 p(res["gflow1"]), p(res["gblend0"]), p(res["gblend1"]))
    if rc:
        raise ValueError(f"oracle bidir_contig: code {rc}")
    return res


_PAD = {"zeros": 0, "border": 1}  # FWB_PAD_ZEROS / FWB_PAD_BORDER (include/flowwarp_b200.h)
