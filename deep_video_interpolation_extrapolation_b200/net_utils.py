"""Drop-in for the warp helpers of the reference's utils/net_utils.py, on the B200-native op.

Same names, argument order and tensor contract as the reference so a maintainer can write
`from deep_video_interpolation_extrapolation_b200.net_utils import FlowWrapper, warp, warp_back`
in place of the originals:

    FlowWrapper            utils/net_utils.py:89-114   parameter-free nn.Module, forward(x, flow)
    warp                   utils/net_utils.py:116-121  -> [N,T,C,H,W], flow gated by mask
    warp_back              utils/net_utils.py:124-129  per-frame source, flow sign flipped
    refine                 utils/net_utils.py:131-147  same signature; the per-frame blend is one streaming kernel
    blend_with_noise       utils/net_utils.py:141-143  input*mask + noise*(1-mask)  (the `refine` pre-blend)
    warp_cat               nets/VAE_S.py:134-135,141   warp of several channel groups into ONE concatenated tensor
    bidirectional_warp     nets/OpticalUnet.py:123-146 forward/backward warps, border padding, mask weighting
    warp_blend             the same two warps fused with their mask-weighted sum (the synthesized frame)

Differences that are deliberate and documented:
  * T frames and several channel groups (RGB + seg) go through ONE kernel launch instead of a Python
    loop + `torch.cat` (`warp_multi`, and `warp`/`warp_back` themselves).
  * `align_corners` is explicit.  The reference passes none (utils/net_utils.py:113): under the torch
    that runs it today that means False (the default here); its pinned torch 1.0.1 behaved as True.
  * CUDA only.  A CPU tensor raises instead of silently taking another path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .ops import flow_warp_blend, label_warp_blend, mask_blend

Tensor = torch.Tensor


class FlowWrapper(nn.Module):
    """utils/net_utils.py:89-114.  No parameters, no buffers (old checkpoints load unchanged)."""

    def __init__(self, align_corners: bool = False, padding_mode: str = "zeros", deterministic: bool = False):
        super().__init__()
        self.align_corners = align_corners
        self.padding_mode = padding_mode
        self.deterministic = deterministic

    def forward(self, x: Tensor, flow: Tensor) -> Tensor:
        # x: [N,C,H,W]; flow: [N,2,H,W] (any strides).  grid = base_grid - flow  (utils/net_utils.py:111)
        if x.dim() != 4 or flow.dim() != 4:
            raise RuntimeError(f"FlowWrapper: expected 4-D x and flow, got {tuple(x.shape)} and {tuple(flow.shape)}")
        return flow_warp_blend([x], [flow], signs=-1.0, padding_mode=self.padding_mode,
                               align_corners=self.align_corners, deterministic=self.deterministic)[0]


def _opts(flowwarpper) -> dict:
    if isinstance(flowwarpper, FlowWrapper):
        return dict(padding_mode=flowwarpper.padding_mode, align_corners=flowwarpper.align_corners,
                    deterministic=flowwarpper.deterministic)
    return dict(padding_mode="zeros", align_corners=False, deterministic=False)


def _check_T(opt, flow: Tensor) -> int:
    T = int(opt.vid_length)
    if flow.dim() != 5 or flow.shape[2] < T:
        raise RuntimeError(f"flow must be [N,2,T>={T},H,W], got {tuple(flow.shape)}")
    return T


def warp(frame: Tensor, flow: Tensor, opt, flowwarpper, mask: Tensor) -> Tensor:
    """utils/net_utils.py:116-121: out[:,i] = FlowWrapper(frame, flow[:,:,i] * mask[:,i:i+1]) for i < opt.vid_length.

    frame [N,C,H,W], flow [N,2,T,H,W], mask [N,T,H,W] -> [N,T,C,H,W]; one launch for all T frames.
    """
    T = _check_T(opt, flow)
    return flow_warp_blend([frame], [flow[:, :, :T]], gates=[mask[:, :T]], signs=-1.0, **_opts(flowwarpper))[0]


def warp_back(frame: Tensor, flowback: Tensor, opt, flowwarpper, mask: Tensor) -> Tensor:
    """utils/net_utils.py:124-129: out[:,i] = FlowWrapper(frame[:,i], -flowback[:,:,i] * mask[:,i:i+1]).

    frame [N,T,C,H,W] (per-frame source) -> [N,T,C,H,W].  `base - (-f*m)` == `base + f*m` bit for bit.
    """
    T = _check_T(opt, flowback)
    return flow_warp_blend([frame[:, :T]], [flowback[:, :, :T]], gates=[mask[:, :T]], signs=+1.0,
                           **_opts(flowwarpper))[0]


def warp_multi(frames: Sequence[Tensor], flow: Tensor, opt, flowwarpper, mask: Tensor) -> List[Tensor]:
    """`warp` for several channel groups that share flow and mask (nets/VAE_S.py:134-135 warps RGB and
    the 20-channel seg map with identical flow/mask in two calls): coordinates are computed once."""
    T = _check_T(opt, flow)
    return flow_warp_blend(list(frames), [flow[:, :, :T]], gates=[mask[:, :T]], signs=-1.0, **_opts(flowwarpper))


def warp_cat(frames: Sequence[Tensor], flow: Tensor, opt, flowwarpper, mask: Tensor) -> Tensor:
    """`torch.cat([warp(f, flow, opt, floww, mask) for f in frames], dim=2)` (nets/VAE_S.py:134-135,141: the RGB and seg warps
    concatenated for the refine net) as ONE launch writing one [N,T,sum(C),H,W] tensor: no per-group outputs, no cat copy."""
    T = _check_T(opt, flow)
    return flow_warp_blend(list(frames), [flow[:, :, :T]], gates=[mask[:, :T]], signs=-1.0, concat=True, **_opts(flowwarpper))[0]


def refine(input: Tensor, flow: Tensor, mask: Tensor, refine_net, opt, noise_bg: Tensor) -> Tensor:
    """Drop-in for utils/net_utils.py:131-147 (same signature, same result).

    The reference blends every frame with four pointwise kernels inside a Python loop (`input[:, i] * mask[:, i:i+1] +
    noise * (1 - mask[:, i:i+1])`, after `cat([noise_bg, zeros(bs, 20, h, w)])` when opt.seg); here all T frames are blended
    by one streaming kernel (channels beyond noise_bg's blend against zero: bit-identical to the cat with zeros), then
    `refine_net` runs per frame exactly as in the reference (:141-144) and the results are concatenated (:146)."""
    T = int(opt.vid_length)
    blended = mask_blend(input[:, :T], mask[:, :T], noise_bg)
    out = [torch.unsqueeze(refine_net(blended[:, i], flow[:, :, i, :, :]), 1) for i in range(T)]
    return torch.cat(out, 1)


def blend_with_noise(input: Tensor, mask: Tensor, noise: Tensor) -> Tensor:
    """utils/net_utils.py:141-143 (the blend in `refine`): input[:,i]*mask[:,i:i+1] + noise*(1-mask[:,i:i+1]) for
    every frame i, as ONE streaming kernel (`ops.mask_blend`: the four pointwise torch kernels per frame and the
    Python loop are gone).  `noise` may carry fewer channels than `input` (the RGB noise of `refine`): the remaining
    channels blend against zero, which replaces the `cat([noise_bg, zeros(bs,20,h,w)])` of :134-136."""
    return mask_blend(input, mask, noise)


def bidirectional_warp(
    frame0: Tensor, frame1: Tensor, for_flow: Tensor, for_mask: Tensor, back_flow: Tensor, back_mask: Tensor,
    align_corners: bool = False, deterministic: bool = False,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """nets/OpticalUnet.py:123-146 with the evident fix at :138 (`back_coor_x/y`).

    for_mask/back_mask are the raw tanh outputs [N,1,H,W]; returns
    (for_output, for_mask3, back_output, back_mask3) exactly as the reference tuple at :148 orders them
    (masks already mapped to 0.5*(1+m) and repeated to the frame's channel count).
    """
    mf = 0.5 * (1.0 + for_mask)
    mb = 0.5 * (1.0 + back_mask)
    C = frame0.shape[1]
    kw = dict(padding_mode="border", align_corners=align_corners, deterministic=deterministic)
    for_out = flow_warp_blend([frame0], [for_flow], blends=[mf], signs=-1.0, **kw)[0]
    back_out = flow_warp_blend([frame1], [back_flow], blends=[mb], signs=+1.0, **kw)[0]
    return for_out, mf.repeat(1, C, 1, 1), back_out, mb.repeat(1, C, 1, 1)


def warp_blend(
    frames0: Sequence[Tensor], frames1: Sequence[Tensor], for_flow: Tensor, back_flow: Tensor,
    for_mask: Tensor, back_mask: Tensor, padding_mode: str = "border", align_corners: bool = False,
    deterministic: bool = False,
) -> List[Tensor]:
    """The headline fused op: synthesized frame = for_mask * warp(frame0, base - for_flow)
    + back_mask * warp(frame1, base + back_flow), for every channel group (RGB, seg, ...) in one launch.

    frames0[g], frames1[g]: [N,C_g,H,W]; flows [N,2,H,W]; masks [N,1,H,W] (already in [0,1]).
    Semantics: nets/OpticalUnet.py:123-146 followed by the sum of the two weighted warps.
    """
    if len(frames0) != len(frames1):
        raise ValueError("warp_blend: frames0 and frames1 need the same number of channel groups")
    return flow_warp_blend([(a, b) for a, b in zip(frames0, frames1)], [for_flow, back_flow],
                           blends=[for_mask, back_mask], signs=[-1.0, +1.0], padding_mode=padding_mode,
                           align_corners=align_corners, deterministic=deterministic)


def warp_blend_labels(
    frames0: Sequence[Tensor], frames1: Sequence[Tensor], labels0: Tensor, labels1: Tensor, num_classes: int,
    for_flow: Tensor, back_flow: Tensor, for_mask: Tensor, back_mask: Tensor, padding_mode: str = "border",
    align_corners: bool = False, deterministic: bool = False,
) -> List[Tensor]:
    """`warp_blend` with the segmentation maps given as uint8 label maps [N,H,W] instead of one-hot float tensors
    (the compact format of SURVEY §8f row 4): returns [*blended float groups, blended seg [N,K,H,W]], the last entry
    bit-identical to `warp_blend([one_hot(labels0)], [one_hot(labels1)], ...)`.  The float groups (RGB) go through the
    dense kernels, the label maps through the label kernels; autograd sums their flow / mask gradients."""
    outs = warp_blend(frames0, frames1, for_flow, back_flow, for_mask, back_mask, padding_mode=padding_mode,
                      align_corners=align_corners, deterministic=deterministic) if len(frames0) else []
    seg = label_warp_blend([labels0, labels1], num_classes, [for_flow, back_flow], blends=[for_mask, back_mask],
                           signs=[-1.0, +1.0], padding_mode=padding_mode, align_corners=align_corners)
    return [*outs, seg]
