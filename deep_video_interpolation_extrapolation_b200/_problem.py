"""Fill the C-ABI structs (include/flowwarp_b200.h) from array-likes.

Array-agnostic on purpose: the product passes CUDA torch tensors, the test oracle passes numpy
arrays through the very same struct layout, so both sides see the identical problem description.

Canonical logical shapes (see the header):
    flow  [N,2,T,H,W]   gate/blend [N,T,H,W]   src [N,T,C,H,W] (T-stride 0 = one frame for all T)
    out / grad_out [N,T,C,H,W]
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from . import _lib as L


def _chk_w(st, what, W=None):
    # a size-1 innermost dimension may carry any stride (numpy / torch views do): it is never stepped over
    if st[-1] != 1 and W != 1:
        raise ValueError(f"{what}: innermost (W) stride must be 1, got {st[-1]}")


def fill_problem(
    *, N: int, T: int, H: int, W: int,
    flows: Sequence, gates: Sequence, blends: Sequence, signs: Sequence[float],
    srcs: Sequence[Sequence], outs: Sequence,
    padding_mode: int, align_corners: bool, flags: int,
    ptr: Callable, strides: Callable,
) -> L.fwb_problem:
    """flows[d] 5-D, gates[d]/blends[d] 4-D or None, srcs[g][d] 5-D, outs[g] 5-D or None."""
    p = L.fwb_problem()
    p.N, p.T, p.H, p.W = N, T, H, W
    p.n_dirs, p.n_groups = len(flows), len(srcs)
    p.padding_mode, p.align_corners, p.flags = padding_mode, int(bool(align_corners)), flags
    for d in range(p.n_dirs):
        D = p.dir[d]
        st = strides(flows[d])
        _chk_w(st, "flow", p.W)
        D.flow = ptr(flows[d])
        D.flow_sn, D.flow_sc, D.flow_st, D.flow_sh = st[0], st[1], st[2], st[3]
        if gates[d] is not None:
            st = strides(gates[d])
            _chk_w(st, "gate", p.W)
            D.gate = ptr(gates[d])
            D.gate_sn, D.gate_st, D.gate_sh = st[0], st[1], st[2]
        if blends[d] is not None:
            st = strides(blends[d])
            _chk_w(st, "blend", p.W)
            D.blend = ptr(blends[d])
            D.blend_sn, D.blend_st, D.blend_sh = st[0], st[1], st[2]
        D.sign = float(signs[d])
    for g in range(p.n_groups):
        R = p.grp[g]
        for d in range(p.n_dirs):
            s = srcs[g][d]
            st = strides(s)
            _chk_w(st, "src", p.W)
            R.src[d] = ptr(s)
            R.src_sn[d], R.src_st[d], R.src_sc[d], R.src_sh[d] = st[0], st[1], st[2], st[3]
        R.C = _shape(srcs[g][0])[2]
        if outs is not None and outs[g] is not None:
            st = strides(outs[g])
            _chk_w(st, "out", p.W)
            R.out = ptr(outs[g])
            R.out_sn, R.out_st, R.out_sc, R.out_sh = st[0], st[1], st[2], st[3]
    return p


def _shape(x):
    return tuple(x.shape)


def fill_grads(
    p: L.fwb_problem, *, grad_outs: Sequence, grad_srcs: Sequence[Sequence], grad_flows: Sequence,
    grad_gates: Sequence, grad_blends: Sequence, ptr: Callable, strides: Callable,
) -> L.fwb_grads:
    """grad_outs[g] 5-D or None; grad_srcs[g][d] 5-D (T-stride 0 when the source is shared) or None."""
    q = L.fwb_grads()
    for g in range(p.n_groups):
        if grad_outs[g] is not None:
            st = strides(grad_outs[g])
            _chk_w(st, "grad_out", p.W)
            q.grad_out[g] = ptr(grad_outs[g])
            q.go_sn[g], q.go_st[g], q.go_sc[g], q.go_sh[g] = st[0], st[1], st[2], st[3]
        for d in range(p.n_dirs):
            x = grad_srcs[g][d]
            if x is not None:
                st = strides(x)
                _chk_w(st, "grad_src", p.W)
                q.grad_src[g][d] = ptr(x)
                q.gs_sn[g][d], q.gs_st[g][d], q.gs_sc[g][d], q.gs_sh[g][d] = st[0], st[1], st[2], st[3]
    for d in range(p.n_dirs):
        x = grad_flows[d]
        if x is not None:
            st = strides(x)
            _chk_w(st, "grad_flow", p.W)
            q.grad_flow[d] = ptr(x)
            q.gf_sn[d], q.gf_sc[d], q.gf_st[d], q.gf_sh[d] = st[0], st[1], st[2], st[3]
        x = grad_gates[d]
        if x is not None:
            st = strides(x)
            _chk_w(st, "grad_gate", p.W)
            q.grad_gate[d] = ptr(x)
            q.gg_sn[d], q.gg_st[d], q.gg_sh[d] = st[0], st[1], st[2]
        x = grad_blends[d]
        if x is not None:
            st = strides(x)
            _chk_w(st, "grad_blend", p.W)
            q.grad_blend[d] = ptr(x)
            q.gb_sn[d], q.gb_st[d], q.gb_sh[d] = st[0], st[1], st[2]
    return q
