"""B200-native optical-flow backward warp + mask-weighted blend.

A drop-in for ONE hot path of lzhangbj/deep_video_interpolation_extrapolation
(utils/net_utils.py:89-129, nets/OpticalUnet.py:123-146): hand-written sm_100a kernels behind a
C-ABI (include/flowwarp_b200.h), exposed as a PyTorch autograd op.  No CPU / torch fallback.
"""
from .net_utils import (FlowWrapper, bidirectional_warp, blend_with_noise, refine, warp, warp_back, warp_blend,
                        warp_blend_labels, warp_cat, warp_multi)
from .host_pipeline import HostWarpBlend, warp_blend_host
from .ops import flow_warp_blend, label_warp_blend, mask_blend, sample_indices
from .losses import (flow_consistency_loss, flow_gradient_loss, flowconsist, flowgradloss, gradientx, gradienty, masked_l1_mean)

__all__ = [
    "FlowWrapper", "warp", "warp_back", "warp_multi", "warp_cat", "warp_blend", "bidirectional_warp", "blend_with_noise", "refine",
    "flow_warp_blend", "label_warp_blend", "warp_blend_labels", "mask_blend", "sample_indices", "HostWarpBlend", "warp_blend_host",
    "gradientx", "gradienty", "flow_gradient_loss", "flowgradloss", "flow_consistency_loss", "flowconsist", "masked_l1_mean",
]
__version__ = "1.0.0"
