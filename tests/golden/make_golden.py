"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/utils/net_utils.py (FlowWrapper, warp, warp_back — utils/net_utils.py:89-129) under the
installed torch 2.11 (CPU) and records inputs' seeds + outputs + autograd gradients.  Inputs are regenerated
from tests/synth.py by seed, so only small arrays are stored.  /root/reference does not exist on the GPU box:
nothing at test time imports it — tests read these files.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
import synth  # noqa: E402

warnings.filterwarnings("ignore")
from utils.net_utils import FlowWrapper, warp, warp_back  # noqa: E402  (the reference, unmodified)

torch.manual_seed(0)
fw = FlowWrapper()
T_ = torch.from_numpy


def t(a, grad=True):
    x = T_(np.ascontiguousarray(a)).clone()
    x.requires_grad_(grad)
    return x


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print(name, {k: getattr(v, "shape", v) for k, v in kw.items()})


# 1. FlowWrapper forward + autograd, several small shapes (utils/net_utils.py:93-114)
cases = [(2, 3, 12, 20, 2.0), (1, 5, 9, 7, 1.5), (2, 4, 16, 33, 4.0), (1, 2, 1, 17, 2.0), (1, 2, 13, 1, 2.0), (1, 20, 24, 40, 3.0)]
for k, (N, C, H, W, sig) in enumerate(cases):
    x = t(synth.rgb(10 + k, N, H, W, C))
    fl = t(synth.flow(20 + k, N, H, W, sig, oob_frac=0.05))
    go = synth.grad(30 + k, (N, C, H, W))
    out = fw(x, fl)
    out.backward(T_(go))
    save(f"flowwrapper_{k}.npz", shape=np.array([N, C, H, W]), sigma=sig, seeds=np.array([10 + k, 20 + k, 30 + k]),
         out=out.detach().numpy(), grad_x=x.grad.numpy(), grad_flow=fl.grad.numpy())

# 2. warp: one frame, T gated flows (utils/net_utils.py:116-121)
N, C, T, H, W = 2, 4, 3, 10, 16
opt = types.SimpleNamespace(vid_length=T)
x = t(synth.rgb(40, N, H, W, C))
fl = t(synth.flow(41, N, H, W, 2.0, T=T, oob_frac=0.05))
m = t(synth.mask(42, N, H, W, T=T))
go = synth.grad(43, (N, T, C, H, W))
out = warp(x, fl, opt, fw, m)
out.backward(T_(go))
save("warp_0.npz", shape=np.array([N, C, T, H, W]), seeds=np.array([40, 41, 42, 43]), out=out.detach().numpy(),
     grad_x=x.grad.numpy(), grad_flow=fl.grad.numpy(), grad_mask=m.grad.numpy())

# 3. warp_back: per-frame source, sign flipped (utils/net_utils.py:124-129)
xb = t(np.stack([synth.rgb(50 + i, N, H, W, C) for i in range(T)], 1))
fl = t(synth.flow(51, N, H, W, 2.0, T=T, oob_frac=0.05))
m = t(synth.mask(52, N, H, W, T=T))
go = synth.grad(53, (N, T, C, H, W))
out = warp_back(xb, fl, opt, fw, m)
out.backward(T_(go))
save("warp_back_0.npz", shape=np.array([N, C, T, H, W]), seeds=np.array([50, 51, 52, 53]), out=out.detach().numpy(),
     grad_x=xb.grad.numpy(), grad_flow=fl.grad.numpy(), grad_mask=m.grad.numpy())

# 4. exact coordinate probe through the reference itself: a parity image along one axis with the other
#    axis of size 1 makes the bilinear output equal frac(ix) or 1-frac(ix) EXACTLY (both differences are
#    exact in fp32), which pins base grid + subtraction + unnormalisation bit for bit.
rng = np.random.default_rng(60)
for axis, size in [("x", 7), ("x", 150), ("x", 256), ("x", 1000), ("y", 9), ("y", 128), ("y", 513)]:
    M = 64  # rows (or columns) of independent random flows
    if axis == "x":
        H, W = 1, size
        img = (np.arange(W) % 2).astype(np.float32).reshape(1, 1, 1, W).repeat(M, 0)
        fl = np.zeros((M, 2, 1, W), np.float32)
        fl[:, 0] = rng.uniform(-1.2, 1.2, (M, 1, W)).astype(np.float32)
        fl[:, 1] = -1.0  # H == 1: base_y = -1, gy = 0, iy = 0 exactly
    else:
        H, W = size, 1
        img = (np.arange(H) % 2).astype(np.float32).reshape(1, 1, H, 1).repeat(M, 0)
        fl = np.zeros((M, 2, H, 1), np.float32)
        fl[:, 1] = rng.uniform(-1.2, 1.2, (M, H, 1)).astype(np.float32)
        fl[:, 0] = -1.0
    out = fw(T_(img), T_(fl)).numpy()
    save(f"coordprobe_{axis}{size}.npz", axis=axis, size=size, flow=fl, out=out)

# 5. BASELINE config 1: one 3-frame clip (int_5_len_3 layout), 128x256 RGB + 20-class seg, batch 1, forward on CPU:
#    frame1 and frame3 (and seg1, seg3) warped to the middle frame with gated flows.  Stored subsampled.
N, H, W = 1, 128, 256
opt1 = types.SimpleNamespace(vid_length=1)
f1, f3 = synth.rgb(70, N, H, W), synth.rgb(71, N, H, W)
s1, s3 = synth.seg(72, N, H, W), synth.seg(73, N, H, W)
flf, flb = synth.flow(74, N, H, W, 8.0, T=1), synth.flow(75, N, H, W, 8.0, T=1)
mf, mb = synth.mask(76, N, H, W, T=1), synth.mask(77, N, H, W, T=1)
with torch.no_grad():
    o = dict(
        rgb_f=warp(T_(f1), T_(flf), opt1, fw, T_(mf)).numpy(), seg_f=warp(T_(s1), T_(flf), opt1, fw, T_(mf)).numpy(),
        rgb_b=warp_back(T_(f3)[:, None], T_(flb), opt1, fw, T_(mb)).numpy(),
        seg_b=warp_back(T_(s3)[:, None], T_(flb), opt1, fw, T_(mb)).numpy())
save("config1_clip.npz", seeds=np.arange(70, 78), **{k: v[..., ::4, ::4] for k, v in o.items()},
     **{k + "_sum": v.astype(np.float64).sum((0, 1, 3, 4)) for k, v in o.items()})
