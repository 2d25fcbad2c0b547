"""torch restatement of the reference composition (the "stock F.grid_sample path").

TEST INFRASTRUCTURE / BASELINE ONLY — never imported by the product package.  It performs the same
sequence of torch calls the reference does, so that (a) tests on the GPU box can compare against
"the reference's own torch grid_sample path" without /root/reference being present, (b) bench.py can
time that path on the host cores (cpu_baseline, --impl reference) and on the B200 (GPU baseline).
tests/test_oracle_golden.py pins it to the unmodified reference via the committed golden vectors.

    ref_flow_wrapper   utils/net_utils.py:93-114
    ref_warp           utils/net_utils.py:116-121
    ref_warp_back      utils/net_utils.py:124-129
    ref_bidirectional  nets/OpticalUnet.py:7-15,123-146 (with the evident fix at :138)
    ref_warp_blend     the sum of the two weighted warps (the synthesized frame of the north star)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _base_grid(N, H, W, device):
    # utils/net_utils.py:99-107 — built on the CPU with linspace/ger, then moved to x's device
    xs = torch.linspace(-1, 1, W) if W > 1 else torch.tensor([-1.0])
    ys = torch.linspace(-1, 1, H) if H > 1 else torch.tensor([-1.0])
    g = torch.empty(N, H, W, 2)
    g[..., 0] = torch.outer(torch.ones(H), xs)
    g[..., 1] = torch.outer(ys, torch.ones(W))
    return g.to(device)


def ref_flow_wrapper(x, flow, align_corners=False, padding_mode="zeros"):
    N, _, H, W = x.shape
    grid = _base_grid(N, H, W, x.device) - flow.permute(0, 2, 3, 1)  # :109-111
    return F.grid_sample(x, grid, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)  # :113


def ref_warp(frame, flow, T, mask, align_corners=False):
    # utils/net_utils.py:116-121
    return torch.stack([ref_flow_wrapper(frame, flow[:, :, i] * mask[:, i:i + 1], align_corners) for i in range(T)], 1)


def ref_warp_back(frame, flowback, T, mask, align_corners=False):
    # utils/net_utils.py:124-129
    return torch.stack(
        [ref_flow_wrapper(frame[:, i], -flowback[:, :, i] * mask[:, i:i + 1], align_corners) for i in range(T)], 1)


def ref_bidirectional(frame0, frame1, for_flow, for_mask_raw, back_flow, back_mask_raw, align_corners=False):
    # nets/OpticalUnet.py:123-146; masks are the raw tanh outputs
    N, _, H, W = frame0.shape
    base = _base_grid(N, H, W, frame0.device)
    gf = torch.stack([base[..., 0] - for_flow[:, 0], base[..., 1] - for_flow[:, 1]], 3)
    gb = torch.stack([base[..., 0] + back_flow[:, 0], base[..., 1] + back_flow[:, 1]], 3)
    fo = F.grid_sample(frame0, gf, mode="bilinear", padding_mode="border", align_corners=align_corners)
    bo = F.grid_sample(frame1, gb, mode="bilinear", padding_mode="border", align_corners=align_corners)
    C = frame0.shape[1]
    mf = (0.5 * (1.0 + for_mask_raw)).repeat(1, C, 1, 1)
    mb = (0.5 * (1.0 + back_mask_raw)).repeat(1, C, 1, 1)
    return mf * fo, mf, mb * bo, mb


def ref_warp_blend(frames0, frames1, for_flow, back_flow, for_mask, back_mask, padding_mode="border",
                   align_corners=False):
    """Synthesized frame per channel group: for_mask*warp(f0, base-for_flow) + back_mask*warp(f1, base+back_flow).

    Composed the way the reference composes it: one grid_sample call per group and direction (RGB and seg
    are warped in separate calls that rebuild the grid, nets/VAE_S.py:134-135), masks in [0,1]."""
    outs = []
    for a, b in zip(frames0, frames1):
        N, _, H, W = a.shape
        base = _base_grid(N, H, W, a.device)
        gf = base - for_flow.permute(0, 2, 3, 1)
        gb = base + back_flow.permute(0, 2, 3, 1)
        wa = F.grid_sample(a, gf, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)
        wb = F.grid_sample(b, gb, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)
        outs.append(for_mask * wa + back_mask * wb)
    return outs


# --------------------------------------------------------------------------------------------- flow-regularisation losses
# The reference's losses.py is deleted; TrainingLoss._flowgradloss (pyc line 413) and TrainingLoss._flowconsist (pyc line 481)
# are restated here from the bytecode in __pycache__/losses.cpython-36.pyc, on top of utils/net_utils.py:243-248.
def ref_gradientx(img):
    return img[:, :, :, :-1] - img[:, :, :, 1:]


def ref_gradienty(img):
    return img[:, :, :-1, :] - img[:, :, 1:, :]


def ref_flowgradloss_frame(flow, image):
    flow = flow * 128
    image = image * 256
    fx, fy = ref_gradientx(flow), ref_gradienty(flow)
    wx = torch.exp(-torch.mean(torch.abs(ref_gradientx(image)), 1, keepdim=True))
    wy = torch.exp(-torch.mean(torch.abs(ref_gradienty(image)), 1, keepdim=True))
    return torch.mean(torch.abs(fx * wx)) + torch.mean(torch.abs(fy * wy))


def ref_flowgradloss(flow, image, t):
    """flow [N,2,T,H,W], image [N,T,C,H,W] (TrainingLoss.flowgradloss: sum over the frames / t)."""
    s = 0.0
    for i in range(t):
        s = s + ref_flowgradloss_frame(flow[:, :, i], image[:, i])
    return s / t


def ref_flowconsist(flow, flowback, mask_fw, mask_bw, t, align_corners=False):
    """TrainingLoss.flowconsist: sum over the frames of prev + next (masks only when both are given)."""
    s = 0.0
    for i in range(t):
        f, b = flow[:, :, i], flowback[:, :, i]
        prev = torch.abs(ref_flow_wrapper(f, -b, align_corners) - b)
        nxt = torch.abs(ref_flow_wrapper(b, f, align_corners) - f)
        if mask_fw is not None and mask_bw is not None:
            prev, nxt = mask_bw[:, i:i + 1] * prev, mask_fw[:, i:i + 1] * nxt
        s = s + prev.mean() + nxt.mean()
    return s
