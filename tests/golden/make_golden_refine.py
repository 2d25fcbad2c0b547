"""Generate tests/golden/refine_*.npz from the UNMODIFIED reference `refine` (utils/net_utils.py:131-150).

    python tests/golden/make_golden_refine.py      (build container only: imports /root/reference)

`refine` pushes `input[:, i] * mask[:, i:i+1] + noise * (1 - mask[:, i:i+1])` through `refine_net` per frame; with an
identity `refine_net` its return value IS the blend, noise = cat([noise_bg, zeros(bs, 20, h, w)]) for opt.seg (:134-136).
Inputs are regenerated from tests/synth.py by seed; outputs and autograd gradients are stored.
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
from utils.net_utils import refine  # noqa: E402  (the reference, unmodified)

identity = lambda x, flow: x  # noqa: E731
for k, (N, T, H, W, seg) in enumerate([(2, 3, 10, 16, True), (1, 2, 7, 9, False)]):
    C = 23 if seg else 3
    opt = types.SimpleNamespace(vid_length=T, seg=seg)
    inp = torch.from_numpy(synth.grad(80 + k, (N, T, C, H, W))).requires_grad_()
    mask = torch.from_numpy(synth.mask(81 + k, N, H, W, T=T)).requires_grad_()
    noise = torch.from_numpy(synth.rgb(82 + k, N, H, W, 3)).requires_grad_()
    flow = torch.zeros(N, 2, T, H, W)
    go = synth.grad(83 + k, (N, T, C, H, W))
    out = refine(inp, flow, mask, identity, opt, noise)
    out.backward(torch.from_numpy(go))
    name = f"refine_{k}.npz"
    np.savez_compressed(os.path.join(HERE, name), shape=np.array([N, T, C, H, W]), seeds=np.array([80 + k, 81 + k, 82 + k, 83 + k]),
                        out=out.detach().numpy(), grad_input=inp.grad.numpy(), grad_mask=mask.grad.numpy(),
                        grad_noise=noise.grad.numpy())
    print(name, tuple(out.shape))
