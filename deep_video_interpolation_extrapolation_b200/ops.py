"""PyTorch custom op over the C-ABI: fused flow warp (+gate) (+blend), forward and backward.

`flow_warp_blend` is the one op every reference call site maps onto:

    FlowWrapper.forward   utils/net_utils.py:93-114     1 direction, sign -1, zeros padding
    warp / warp_back      utils/net_utils.py:116-129    + gate mask, T frames in one launch
    OpticalUnet warps     nets/OpticalUnet.py:123-146   2 directions, border padding, blend masks

torch is used for device memory, streams and autograd plumbing only; all arithmetic runs in
libflowwarp_b200.so.  There is no CPU path: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch.autograd.function import once_differentiable

from . import _lib as L
from ._problem import fill_grads, fill_problem

Tensor = torch.Tensor

# A/B switch for measurements: True routes grad_src through the ATen-style global-atomic scatter kernel instead of
# the (deterministic) owner-gather kernel.  Not a fallback: both are CUDA kernels of this library.
ATOMIC_SRC = False

# True: when a source requires grad (and the fused backward will be used) the forward call allocates grad_src and lets
# the forward kernel zero-fill it (fwb_warp_blend_forward_zero), so the backward runs without a memset pass.  Costs the
# grad_src buffers being alive between forward and backward; False restores "allocate and zero in backward".
PREZERO_GRAD_SRC = True


def _flags(cfg) -> int:
    if cfg.deterministic:
        return L.FWB_FLAG_DETERMINISTIC
    return L.FWB_FLAG_ATOMIC_SRC if ATOMIC_SRC else L.FWB_FLAG_FUSED_BWD


@dataclass(frozen=True)
class _Cfg:
    n_dirs: int
    n_groups: int
    signs: Tuple[float, ...]
    padding_mode: int
    align_corners: bool
    deterministic: bool
    N: int
    T: int
    H: int
    W: int
    concat: bool = False  # one [N,T,sum(C),H,W] output, the groups' channels back to back (nets/VAE_S.py:141's torch.cat)


def _ptr(x: Tensor) -> int:
    return x.data_ptr()


def _strides(x: Tensor):
    return x.stride()


def _wcontig(x: Optional[Tensor]) -> Optional[Tensor]:
    if x is None:
        return None
    if x.stride(-1) != 1 and x.size(-1) != 1:
        return x.contiguous()
    if x.size(-1) == 1 and x.stride(-1) != 1:
        return x.contiguous()
    return x


def _expand_t(s: Tensor, T: int) -> Tensor:
    """[N,1,C,H,W] -> stride-0 view [N,T,C,H,W] (one source frame feeds all T flows, utils/net_utils.py:118)."""
    return s.expand(s.shape[0], T, *s.shape[2:]) if s.shape[1] != T else s


def _stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _alloc_grad_src(s: Tensor, cfg, shared: bool) -> Tensor:
    """grad buffer of source `s` ([N,T,C,H,W]).  shared: the op itself broadcast a single frame [N,1,C,H,W] over the T flows
    (utils/net_utils.py:118) -> one plane set the kernels sum into (T-stride 0).  A source the CALLER expanded to [N,T,...]
    keeps a per-frame gradient: autograd's own expand backward sums it."""
    buf = torch.empty((cfg.N, 1 if shared else cfg.T, s.shape[2], cfg.H, cfg.W), dtype=torch.float32, device=s.device)
    return buf.expand(cfg.N, cfg.T, s.shape[2], cfg.H, cfg.W) if shared else buf


class _WarpBlendFn(torch.autograd.Function):
    """tensors = flows[D] + gates[D] + blends[D] + srcs[G*D] (g-major); all already canonical 5-D/4-D."""

    @staticmethod
    def forward(ctx, cfg: _Cfg, *tensors):
        D, G = cfg.n_dirs, cfg.n_groups
        tensors = tuple(_wcontig(t) for t in tensors)
        flows, gates, blends = tensors[:D], tensors[D:2 * D], tensors[2 * D:3 * D]
        shared = [[s.shape[1] == 1 and cfg.T > 1 for s in tensors[3 * D + g * D: 3 * D + (g + 1) * D]] for g in range(G)]
        srcs = [[_expand_t(s, cfg.T) for s in tensors[3 * D + g * D: 3 * D + (g + 1) * D]] for g in range(G)]
        dev = flows[0].device
        if cfg.concat:  # the kernels take per-group strides: every group writes its channel slice of ONE buffer
            Cs = [srcs[g][0].shape[2] for g in range(G)]
            big = torch.empty((cfg.N, cfg.T, sum(Cs), cfg.H, cfg.W), dtype=torch.float32, device=dev)
            offs = [sum(Cs[:g]) for g in range(G)]
            outs = [big[:, :, offs[g]:offs[g] + Cs[g]] for g in range(G)]
        else:
            outs = [torch.empty((cfg.N, cfg.T, srcs[g][0].shape[2], cfg.H, cfg.W), dtype=torch.float32, device=dev)
                    for g in range(G)]
        lib = L.load()
        with torch.cuda.device(dev):
            p = fill_problem(N=cfg.N, T=cfg.T, H=cfg.H, W=cfg.W, flows=flows, gates=gates, blends=blends,
                             signs=cfg.signs, srcs=srcs, outs=outs, padding_mode=cfg.padding_mode,
                             align_corners=cfg.align_corners,
                             flags=_flags(cfg),
                             ptr=_ptr, strides=_strides)
            need_src = ctx.needs_input_grad[1 + 3 * D:]
            pre = None
            if PREZERO_GRAD_SRC and _flags(cfg) == L.FWB_FLAG_FUSED_BWD and any(need_src):
                pre = [[_alloc_grad_src(srcs[g][d], cfg, shared[g][d]) if need_src[g * D + d] else None for d in range(D)]
                       for g in range(G)]
                q = fill_grads(p, grad_outs=[None] * G, grad_srcs=pre, grad_flows=[None] * D, grad_gates=[None] * D,
                               grad_blends=[None] * D, ptr=_ptr, strides=_strides)
                L.check(lib.fwb_warp_blend_forward_zero(ctypes.byref(p), ctypes.byref(q), _stream_ptr(dev)),
                        "fwb_warp_blend_forward_zero")
            else:
                L.check(lib.fwb_warp_blend_forward(ctypes.byref(p), _stream_ptr(dev)), "fwb_warp_blend_forward")
        ctx.pre_gsrcs = pre
        ctx.cfg = cfg
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        return (big,) if cfg.concat else tuple(outs)

    @staticmethod
    @once_differentiable  # the gradients are computed by CUDA kernels autograd cannot see through: no silent double backward
    def backward(ctx, *grad_outs):
        cfg: _Cfg = ctx.cfg
        D, G = cfg.n_dirs, cfg.n_groups
        it = iter(ctx.saved_tensors)
        tensors = [next(it) if pr else None for pr in ctx.present]
        flows, gates, blends = tensors[:D], tensors[D:2 * D], tensors[2 * D:3 * D]
        shared = [[s.shape[1] == 1 and cfg.T > 1 for s in tensors[3 * D + g * D: 3 * D + (g + 1) * D]] for g in range(G)]
        srcs = [[_expand_t(s, cfg.T) for s in tensors[3 * D + g * D: 3 * D + (g + 1) * D]] for g in range(G)]
        need = ctx.needs_input_grad[1:]
        dev = flows[0].device
        N, T, H, W = cfg.N, cfg.T, cfg.H, cfg.W

        if cfg.concat:  # one upstream gradient for the concatenated output: channel-slice views, no copy
            if grad_outs[0] is None:
                return (None,) + (None,) * len(tensors)
            gcat = _wcontig(grad_outs[0])
            Cs = [srcs[g][0].shape[2] for g in range(G)]
            grad_outs = tuple(gcat[:, :, sum(Cs[:g]):sum(Cs[:g + 1])] for g in range(G))
        gos = [None if g is None else _wcontig(g) for g in grad_outs]
        if all(g is None for g in gos):
            return (None,) + (None,) * len(tensors)
        f32 = dict(dtype=torch.float32, device=dev)
        g_flows = [torch.empty((N, 2, T, H, W), **f32) if need[d] else None for d in range(D)]
        g_gates = [torch.empty((N, T, H, W), **f32) if (need[D + d] and gates[d] is not None) else None
                   for d in range(D)]
        g_blends = [torch.empty((N, T, H, W), **f32) if (need[2 * D + d] and blends[d] is not None) else None
                    for d in range(D)]
        pre, ctx.pre_gsrcs = ctx.pre_gsrcs, None  # zero-filled by the forward; good for ONE backward
        g_srcs: List[List[Optional[Tensor]]] = []
        for g in range(G):
            row = []
            for d in range(D):
                if need[3 * D + g * D + d] and gos[g] is not None:
                    row.append(pre[g][d] if pre is not None else _alloc_grad_src(srcs[g][d], cfg, shared[g][d]))
                else:
                    row.append(None)
            g_srcs.append(row)
        flags = _flags(cfg) | (L.FWB_FLAG_GRAD_SRC_ZEROED if pre is not None else 0)

        lib = L.load()
        with torch.cuda.device(dev):
            p = fill_problem(N=N, T=T, H=H, W=W, flows=flows, gates=gates, blends=blends, signs=cfg.signs,
                             srcs=srcs, outs=None, padding_mode=cfg.padding_mode,
                             align_corners=cfg.align_corners,
                             flags=flags,
                             ptr=_ptr, strides=_strides)
            q = fill_grads(p, grad_outs=gos, grad_srcs=g_srcs, grad_flows=g_flows, grad_gates=g_gates,
                           grad_blends=g_blends, ptr=_ptr, strides=_strides)
            nbytes = int(lib.fwb_workspace_bytes(ctypes.byref(p)))
            ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
            st = _stream_ptr(dev)
            want_src = any(x is not None for row in g_srcs for x in row)
            want_flow = any(x is not None for x in g_flows + g_gates + g_blends)
            # kernel 2 also fills the segment tables kernel 3 consumes -> always first
            if want_flow or want_src:
                L.check(lib.fwb_warp_blend_backward_flow(ctypes.byref(p), ctypes.byref(q), ws.data_ptr(), nbytes, st),
                        "fwb_warp_blend_backward_flow")
            if want_src:
                L.check(lib.fwb_warp_blend_backward_src(ctypes.byref(p), ctypes.byref(q), ws.data_ptr(), nbytes, st),
                        "fwb_warp_blend_backward_src")
        flat_src = []
        for g in range(G):
            for d in range(D):
                x = g_srcs[g][d]
                if x is not None and shared[g][d]:
                    x = x[:, :1]  # the kernel already summed over the T frames that share this source
                flat_src.append(x)
        return (None, *g_flows, *g_gates, *g_blends, *flat_src)


def _as_list(x, n, what):
    if x is None:
        return [None] * n
    if isinstance(x, Tensor):
        x = [x]
    x = list(x)
    if len(x) != n:
        raise ValueError(f"{what}: expected {n} entries, got {len(x)}")
    return x


def _canon_mask(m: Optional[Tensor], N, T, H, W, what) -> Optional[Tensor]:
    if m is None:
        return None
    if m.dim() == 3:  # [N,H,W]
        m = m.unsqueeze(1)
    if m.dim() == 5 and m.shape[2] == 1:  # [N,T,1,H,W]
        m = m.squeeze(2)
    if m.dim() != 4:
        raise ValueError(f"{what}: expected [N,T,H,W] / [N,1,H,W] / [N,H,W], got {tuple(m.shape)}")
    if m.shape[0] != N or m.shape[2] != H or m.shape[3] != W or m.shape[1] not in (1, T):
        raise ValueError(f"{what}: shape {tuple(m.shape)} does not match N={N} T={T} H={H} W={W}")
    if m.shape[1] != T:
        m = m.expand(N, T, H, W)
    return m


def flow_warp_blend(
    srcs: Sequence[Union[Tensor, Sequence[Tensor]]],
    flows: Union[Tensor, Sequence[Tensor]],
    gates: Union[None, Tensor, Sequence[Optional[Tensor]]] = None,
    blends: Union[None, Tensor, Sequence[Optional[Tensor]]] = None,
    signs: Union[None, float, Sequence[float]] = None,
    padding_mode: str = "zeros",
    align_corners: bool = False,
    deterministic: bool = False,
    concat: bool = False,
) -> List[Tensor]:
    """out[g] = sum_d blend_d * grid_sample(srcs[g][d], base + sign_d * flow_d * gate_d).

    srcs    one entry per channel group (e.g. RGB, seg) sharing the flows/masks; each entry is a
            tensor (1 direction) or a sequence with one tensor per direction; [N,C,H,W] or [N,T,C,H,W]
    flows   one per direction; [N,2,H,W] or [N,2,T,H,W]; normalised units, channel 0 horizontal
    gates   per direction or None; flow-gating mask (utils/net_utils.py:118)
    blends  per direction or None; blend weight of that direction's warp (nets/OpticalUnet.py:141-146)
    signs   per direction, -1 (`base - flow`, default) or +1 (`base + flow`)
    concat  write all groups into ONE tensor, channels back to back (the `torch.cat([output, output_seg], dim=2)` of
            nets/VAE_S.py:141 without the copy); the returned list then has that single tensor
    Returns one [N,C,H,W] (all inputs 4-D) or [N,T,C,H,W] tensor per group.
    """
    if isinstance(flows, Tensor):
        flows = [flows]
    flows = list(flows)
    D = len(flows)
    if D not in (1, 2):
        raise ValueError(f"flow_warp_blend: 1 or 2 directions supported, got {D}")
    if isinstance(srcs, Tensor):
        srcs = [srcs]
    groups = [[s] if isinstance(s, Tensor) else list(s) for s in srcs]
    G = len(groups)
    if not 1 <= G <= L.FWB_MAX_GROUPS:
        raise ValueError(f"flow_warp_blend: 1..{L.FWB_MAX_GROUPS} channel groups supported, got {G}")
    gates, blends = _as_list(gates, D, "gates"), _as_list(blends, D, "blends")
    if signs is None:
        signs = [-1.0] * D
    elif isinstance(signs, (int, float)):
        signs = [float(signs)] * D
    signs = tuple(float(s) for s in signs)
    if len(signs) != D or any(s not in (-1.0, 1.0) for s in signs):
        raise ValueError("signs: one of -1/+1 per direction")
    if padding_mode not in ("zeros", "border"):
        raise ValueError(f"padding_mode must be 'zeros' or 'border', got {padding_mode!r}")

    f0 = flows[0]
    every = [t for t in flows + gates + blends + [s for g in groups for s in g] if t is not None]
    for t in every:
        if not isinstance(t, Tensor):
            raise TypeError("flow_warp_blend: tensors expected")
        if not t.is_cuda:
            raise RuntimeError("flow_warp_blend: CUDA tensors required (this library has no CPU path)")
        if t.device != f0.device:
            raise RuntimeError(f"flow_warp_blend: all tensors must be on {f0.device}, got {t.device}")
        if t.dtype != torch.float32:
            raise RuntimeError(f"flow_warp_blend: float32 required, got {t.dtype}")

    five_d = any(f.dim() == 5 for f in flows) or any(s.dim() == 5 for g in groups for s in g)
    cflows = []
    for f in flows:
        if f.dim() == 4:
            f = f.unsqueeze(2)
        if f.dim() != 5 or f.shape[1] != 2:
            raise RuntimeError(f"flow must be [N,2,H,W] or [N,2,T,H,W], got {tuple(f.shape)}")
        cflows.append(f)
    N, _, T, H, W = cflows[0].shape
    if H < 1 or W < 1:
        raise RuntimeError(f"flow_warp_blend: non-empty spatial dims required, got H={H} W={W}")
    for f in cflows:
        if tuple(f.shape) != (N, 2, T, H, W):
            raise RuntimeError("flow_warp_blend: all flows must share one shape")
    cgates = [_canon_mask(m, N, T, H, W, "gate") for m in gates]
    cblends = [_canon_mask(m, N, T, H, W, "blend") for m in blends]
    csrcs = []
    for g in groups:
        if len(g) != D:
            raise ValueError(f"each source group needs {D} tensors (one per direction), got {len(g)}")
        row = []
        for s in g:
            if s.dim() == 4:
                s = s.unsqueeze(1)
            if s.dim() != 5:
                raise RuntimeError(f"source must be [N,C,H,W] or [N,T,C,H,W], got {tuple(s.shape)}")
            if s.shape[0] != N or s.shape[3] != H or s.shape[4] != W:
                raise RuntimeError(
                    f"source {tuple(s.shape)} and flow {tuple(cflows[0].shape)} disagree on batch or spatial size")
            if s.shape[1] not in (1, T):
                raise RuntimeError(f"source has {s.shape[1]} frames, flow has {T}")
            if s.shape[2] < 1:
                raise RuntimeError("source needs at least one channel")
            row.append(s)  # a single frame stays [N,1,C,H,W]; the Function expands it (T-stride 0)
        if any(s.shape[2] != row[0].shape[2] for s in row):
            raise RuntimeError("both directions of a group must have the same channel count")
        csrcs.append(row)

    if N == 0:  # nothing to launch (empty tensors have no storage to point the C-ABI at)
        outs = [s.new_empty((0, T, s.shape[2], H, W)) for s in (row[0] for row in csrcs)]
        if concat:
            outs = [torch.cat(outs, 2)]
        return outs if five_d else [o.squeeze(1) for o in outs]
    cfg = _Cfg(D, G, signs, L.FWB_PAD_BORDER if padding_mode == "border" else L.FWB_PAD_ZEROS,
               bool(align_corners), bool(deterministic), N, T, H, W, bool(concat))
    flat = [*cflows, *cgates, *cblends, *[s for row in csrcs for s in row]]
    outs = list(_WarpBlendFn.apply(cfg, *flat))
    if not five_d:
        outs = [o.squeeze(1) for o in outs]
    return outs


def _fill_dirs(p, flows, gates, blends, signs):
    for d, f in enumerate(flows):
        D = p.dir[d]
        D.flow = f.data_ptr()
        D.flow_sn, D.flow_sc, D.flow_st, D.flow_sh = f.stride()[:4]
        if gates[d] is not None:
            D.gate = gates[d].data_ptr()
            D.gate_sn, D.gate_st, D.gate_sh = gates[d].stride()[:3]
        if blends[d] is not None:
            D.blend = blends[d].data_ptr()
            D.blend_sn, D.blend_st, D.blend_sh = blends[d].stride()[:3]
        D.sign = float(signs[d])


class _LabelWarpFn(torch.autograd.Function):
    """tensors = flows[D] + gates[D] + blends[D] (canonical 5-D / 4-D); labels[D] uint8 [N,T,H,W] ride in cfg-side args."""

    @staticmethod
    def forward(ctx, cfg: _Cfg, K: int, labels, *tensors):
        D = cfg.n_dirs
        tensors = tuple(_wcontig(t) for t in tensors)
        flows, gates, blends = tensors[:D], tensors[D:2 * D], tensors[2 * D:3 * D]
        dev = flows[0].device
        out = torch.empty((cfg.N, cfg.T, K, cfg.H, cfg.W), dtype=torch.float32, device=dev)
        lib = L.load()
        with torch.cuda.device(dev):
            p = _LabelWarpFn._problem(cfg, K, labels, flows, gates, blends)
            p.out = out.data_ptr()
            p.out_sn, p.out_st, p.out_sc, p.out_sh = out.stride()[:4]
            L.check(lib.fwb_label_warp_blend_forward(ctypes.byref(p), _stream_ptr(dev)), "fwb_label_warp_blend_forward")
        ctx.cfg, ctx.K, ctx.labels = cfg, K, labels
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        return out

    @staticmethod
    def _problem(cfg, K, labels, flows, gates, blends):
        p = L.fwb_label_problem()
        p.N, p.T, p.H, p.W, p.n_dirs, p.K = cfg.N, cfg.T, cfg.H, cfg.W, cfg.n_dirs, K
        p.padding_mode, p.align_corners = cfg.padding_mode, int(cfg.align_corners)
        _fill_dirs(p, flows, gates, blends, cfg.signs)
        for d, lab in enumerate(labels):
            p.labels[d] = lab.data_ptr()
            p.lab_sn[d], p.lab_st[d], p.lab_sh[d] = lab.stride()[:3]
        return p

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        cfg: _Cfg = ctx.cfg
        D = cfg.n_dirs
        it = iter(ctx.saved_tensors)
        tensors = [next(it) if pr else None for pr in ctx.present]
        flows, gates, blends = tensors[:D], tensors[D:2 * D], tensors[2 * D:3 * D]
        need = ctx.needs_input_grad[3:]
        dev = flows[0].device
        N, T, H, W = cfg.N, cfg.T, cfg.H, cfg.W
        f32 = dict(dtype=torch.float32, device=dev)
        g_flows = [torch.empty((N, 2, T, H, W), **f32) if need[d] else None for d in range(D)]
        g_gates = [torch.empty((N, T, H, W), **f32) if (need[D + d] and gates[d] is not None) else None for d in range(D)]
        g_blends = [torch.empty((N, T, H, W), **f32) if (need[2 * D + d] and blends[d] is not None) else None for d in range(D)]
        if any(x is not None for x in g_flows + g_gates + g_blends):
            go = _wcontig(go)
            lib = L.load()
            with torch.cuda.device(dev):
                p = _LabelWarpFn._problem(cfg, ctx.K, ctx.labels, flows, gates, blends)
                p.grad_out = go.data_ptr()
                p.go_sn, p.go_st, p.go_sc, p.go_sh = go.stride()[:4]
                for d in range(D):
                    if g_flows[d] is not None:
                        p.grad_flow[d] = g_flows[d].data_ptr()
                        p.gf_sn[d], p.gf_sc[d], p.gf_st[d], p.gf_sh[d] = g_flows[d].stride()[:4]
                    if g_gates[d] is not None:
                        p.grad_gate[d] = g_gates[d].data_ptr()
                        p.gg_sn[d], p.gg_st[d], p.gg_sh[d] = g_gates[d].stride()[:3]
                    if g_blends[d] is not None:
                        p.grad_blend[d] = g_blends[d].data_ptr()
                        p.gb_sn[d], p.gb_st[d], p.gb_sh[d] = g_blends[d].stride()[:3]
                p.accumulate = 0
                L.check(lib.fwb_label_warp_blend_backward(ctypes.byref(p), _stream_ptr(dev)), "fwb_label_warp_blend_backward")
        return (None, None, None, *g_flows, *g_gates, *g_blends)


def label_warp_blend(
    labels: Union[Tensor, Sequence[Tensor]], num_classes: int, flows: Union[Tensor, Sequence[Tensor]],
    gates: Union[None, Tensor, Sequence[Optional[Tensor]]] = None,
    blends: Union[None, Tensor, Sequence[Optional[Tensor]]] = None,
    signs: Union[None, float, Sequence[float]] = None, padding_mode: str = "zeros", align_corners: bool = False,
) -> Tensor:
    """flow_warp_blend for a segmentation map given as uint8 LABELS: returns what
    `flow_warp_blend([one_hot(labels_d) for d], flows, ...)` returns ([N,K,H,W] or [N,T,K,H,W], float32), bit for bit,
    reading 1 byte per tap instead of 4*K (the reference one-hots in folder.py:193-200 and warps the
    K float planes, nets/VAE_S.py:135).  labels: one uint8 tensor per direction, [N,H,W] / [N,1,H,W] / [N,T,H,W];
    labels >= num_classes belong to no class.  Gradients flow to flows, gates and blends (labels have none).
    """
    if isinstance(flows, Tensor):
        flows = [flows]
    flows = list(flows)
    D = len(flows)
    if D not in (1, 2):
        raise ValueError(f"label_warp_blend: 1 or 2 directions supported, got {D}")
    labels = [labels] if isinstance(labels, Tensor) else list(labels)
    if len(labels) != D:
        raise ValueError(f"label_warp_blend: one label map per direction expected ({D}), got {len(labels)}")
    if not 1 <= int(num_classes) <= 256:
        raise ValueError("num_classes must be in 1..256")
    gates, blends = _as_list(gates, D, "gates"), _as_list(blends, D, "blends")
    if signs is None:
        signs = [-1.0] * D
    elif isinstance(signs, (int, float)):
        signs = [float(signs)] * D
    signs = tuple(float(s) for s in signs)
    if len(signs) != D or any(s not in (-1.0, 1.0) for s in signs):
        raise ValueError("signs: one of -1/+1 per direction")
    if padding_mode not in ("zeros", "border"):
        raise ValueError(f"padding_mode must be 'zeros' or 'border', got {padding_mode!r}")
    f0 = flows[0]
    for t in [x for x in flows + gates + blends if x is not None]:
        if not t.is_cuda:
            raise RuntimeError("label_warp_blend: CUDA tensors required (this library has no CPU path)")
        if t.device != f0.device:
            raise RuntimeError(f"label_warp_blend: all tensors must be on {f0.device}, got {t.device}")
        if t.dtype != torch.float32:
            raise RuntimeError(f"label_warp_blend: float32 required, got {t.dtype}")
    five_d = any(f.dim() == 5 for f in flows)
    cflows = []
    for f in flows:
        if f.dim() == 4:
            f = f.unsqueeze(2)
        if f.dim() != 5 or f.shape[1] != 2:
            raise RuntimeError(f"flow must be [N,2,H,W] or [N,2,T,H,W], got {tuple(f.shape)}")
        cflows.append(f)
    N, _, T, H, W = cflows[0].shape
    if H < 1 or W < 1:
        raise RuntimeError(f"label_warp_blend: non-empty spatial dims required, got H={H} W={W}")
    for f in cflows:
        if tuple(f.shape) != (N, 2, T, H, W):
            raise RuntimeError("label_warp_blend: all flows must share one shape")
    clabels = []
    for lab in labels:
        if not lab.is_cuda or lab.device != f0.device:
            raise RuntimeError("label_warp_blend: labels must be CUDA tensors on the flows' device")
        if lab.dtype != torch.uint8:
            raise RuntimeError(f"label_warp_blend: uint8 labels required, got {lab.dtype}")
        if lab.dim() == 3:
            lab = lab.unsqueeze(1)
        if lab.dim() != 4 or lab.shape[0] != N or lab.shape[1] not in (1, T) or tuple(lab.shape[2:]) != (H, W):
            raise RuntimeError(f"labels must be [N,H,W], [N,1,H,W] or [N,T,H,W], got {tuple(lab.shape)}")
        if lab.stride(-1) != 1 and W > 1:
            lab = lab.contiguous()
        clabels.append(lab.expand(N, T, H, W) if lab.shape[1] != T else lab)
    cgates = [_canon_mask(m, N, T, H, W, "gate") for m in gates]
    cblends = [_canon_mask(m, N, T, H, W, "blend") for m in blends]
    if N == 0:
        out = f0.new_empty((0, T, int(num_classes), H, W))
        return out if five_d else out.squeeze(1)
    cfg = _Cfg(D, 1, signs, L.FWB_PAD_BORDER if padding_mode == "border" else L.FWB_PAD_ZEROS, bool(align_corners), True,
               N, T, H, W)
    out = _LabelWarpFn.apply(cfg, int(num_classes), tuple(clabels), *cflows, *cgates, *cblends)
    return out if five_d else out.squeeze(1)


def _fill_blend(inp: Tensor, mask: Tensor, noise: Optional[Tensor]) -> "L.fwb_blend":
    b = L.fwb_blend()
    N, T, C, H, W = inp.shape
    b.N, b.T, b.C, b.H, b.W = N, T, C, H, W
    b.Cn = 0 if noise is None else noise.shape[1]
    b.input = inp.data_ptr()
    b.in_sn, b.in_st, b.in_sc, b.in_sh = inp.stride()[:4]
    b.mask = mask.data_ptr()
    b.m_sn, b.m_st, b.m_sh = mask.stride()[:3]
    if noise is not None:
        b.noise = noise.data_ptr()
        b.nz_sn, b.nz_sc, b.nz_sh = noise.stride()[:3]
    return b


class _MaskBlendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp: Tensor, mask: Tensor, noise: Optional[Tensor]):
        inp, mask, noise = _wcontig(inp), _wcontig(mask), _wcontig(noise)
        out = torch.empty(inp.shape, dtype=torch.float32, device=inp.device)
        lib = L.load()
        with torch.cuda.device(inp.device):
            b = _fill_blend(inp, mask, noise)
            b.out = out.data_ptr()
            b.out_sn, b.out_st, b.out_sc, b.out_sh = out.stride()[:4]
            L.check(lib.fwb_mask_blend_forward(ctypes.byref(b), _stream_ptr(inp.device)), "fwb_mask_blend_forward")
        ctx.save_for_backward(inp, mask, noise)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, go: Tensor):
        inp, mask, noise = ctx.saved_tensors
        go = _wcontig(go)
        f32 = dict(dtype=torch.float32, device=inp.device)
        gi = torch.empty(inp.shape, **f32) if ctx.needs_input_grad[0] else None
        gm = torch.empty(mask.shape, **f32) if ctx.needs_input_grad[1] else None
        gn = torch.empty(noise.shape, **f32) if (noise is not None and ctx.needs_input_grad[2]) else None
        lib = L.load()
        with torch.cuda.device(inp.device):
            b = _fill_blend(inp, mask, noise)
            b.grad_out = go.data_ptr()
            b.go_sn, b.go_st, b.go_sc, b.go_sh = go.stride()[:4]
            if gi is not None:
                b.grad_input = gi.data_ptr()
                b.gi_sn, b.gi_st, b.gi_sc, b.gi_sh = gi.stride()[:4]
            if gm is not None:
                b.grad_mask = gm.data_ptr()
                b.gm_sn, b.gm_st, b.gm_sh = gm.stride()[:3]
            if gn is not None:
                b.grad_noise = gn.data_ptr()
                b.gn_sn, b.gn_sc, b.gn_sh = gn.stride()[:3]
            L.check(lib.fwb_mask_blend_backward(ctypes.byref(b), _stream_ptr(inp.device)), "fwb_mask_blend_backward")
        return gi, gm, gn


def mask_blend(input: Tensor, mask: Tensor, noise: Optional[Tensor] = None) -> Tensor:
    """out[:, i] = input[:, i] * mask[:, i:i+1] + noise * (1 - mask[:, i:i+1])   (utils/net_utils.py:141-143)

    input [N,T,C,H,W]; mask [N,T,H,W]; noise [N,Cn,H,W] with Cn <= C or None: channels without a noise plane blend
    against zero, which is what `refine` builds with `torch.cat([noise_bg, zeros(bs, 20, h, w)])` when opt.seg
    (utils/net_utils.py:134-136).  One streaming kernel forward, one backward; bit-identical to the torch expression.
    """
    every = [t for t in (input, mask, noise) if t is not None]
    for t in every:
        if not isinstance(t, Tensor):
            raise TypeError("mask_blend: tensors expected")
        if not t.is_cuda:
            raise RuntimeError("mask_blend: CUDA tensors required (this library has no CPU path)")
        if t.device != input.device:
            raise RuntimeError(f"mask_blend: all tensors must be on {input.device}, got {t.device}")
        if t.dtype != torch.float32:
            raise RuntimeError(f"mask_blend: float32 required, got {t.dtype}")
    if input.dim() != 5:
        raise RuntimeError(f"mask_blend: input must be [N,T,C,H,W], got {tuple(input.shape)}")
    N, T, C, H, W = input.shape
    if H < 1 or W < 1 or C < 1 or T < 1:
        raise RuntimeError("mask_blend: non-empty T, C and spatial dims required")
    if mask.dim() != 4 or mask.shape[0] != N or mask.shape[1] < T or tuple(mask.shape[2:]) != (H, W):
        raise RuntimeError(f"mask_blend: mask must be [N,T>={T},H,W], got {tuple(mask.shape)}")
    mask = mask[:, :T]
    if noise is not None:
        if noise.dim() != 4 or noise.shape[0] != N or noise.shape[1] > C or tuple(noise.shape[2:]) != (H, W):
            raise RuntimeError(f"mask_blend: noise must be [N,Cn<={C},H,W], got {tuple(noise.shape)}")
        if noise.shape[1] == 0:
            noise = None
    if N == 0:
        return input.new_empty(input.shape)
    return _MaskBlendFn.apply(input, mask, noise)


def sample_indices(flow: Tensor, gate: Optional[Tensor] = None, sign: float = -1.0,
                   padding_mode: str = "zeros", align_corners: bool = False):
    """Debug / parity: (x0, y0, valid_bits, ix, iy) the kernels use for `flow` [N,2,H,W] or [N,2,T,H,W].

    x0,y0 int32, valid uint8 (bit0 nw, bit1 ne, bit2 sw, bit3 se), ix,iy float32; shape [N,T,H,W].
    """
    if not flow.is_cuda:
        raise RuntimeError("sample_indices: CUDA tensors required")
    f = flow.unsqueeze(2) if flow.dim() == 4 else flow
    f = _wcontig(f)
    N, _, T, H, W = f.shape
    g = _wcontig(_canon_mask(gate, N, T, H, W, "gate"))
    dev = f.device
    dummy = torch.empty((N, 1, 1, H, W), dtype=torch.float32, device=dev).expand(N, T, 1, H, W)
    x0 = torch.empty((N, T, H, W), dtype=torch.int32, device=dev)
    y0 = torch.empty_like(x0)
    valid = torch.empty((N, T, H, W), dtype=torch.uint8, device=dev)
    ix = torch.empty((N, T, H, W), dtype=torch.float32, device=dev)
    iy = torch.empty_like(ix)
    lib = L.load()
    with torch.cuda.device(dev):
        p = fill_problem(N=N, T=T, H=H, W=W, flows=[f], gates=[g], blends=[None], signs=[sign], srcs=[[dummy]],
                         outs=None, padding_mode=L.FWB_PAD_BORDER if padding_mode == "border" else L.FWB_PAD_ZEROS,
                         align_corners=align_corners, flags=0, ptr=_ptr, strides=_strides)
        L.check(lib.fwb_sample_indices(ctypes.byref(p), 0, x0.data_ptr(), y0.data_ptr(), valid.data_ptr(),
                                       ix.data_ptr(), iy.data_ptr(), _stream_ptr(dev)), "fwb_sample_indices")
    return x0, y0, valid, ix, iy
