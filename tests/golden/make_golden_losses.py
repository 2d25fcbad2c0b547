"""Generate tests/golden/losses_*.npz: the flow-regularisation losses of the reference's TrainingLoss.

    python tests/golden/make_golden_losses.py      (build container only: imports /root/reference)

`losses.py` is deleted from the reference, its bytecode survives (__pycache__/losses.cpython-36.pyc).  The two functions below
restate TrainingLoss._flowgradloss (pyc line 413) and TrainingLoss._flowconsist (pyc line 481) in the order of that bytecode,
and every building block they call is the UNMODIFIED reference: `gradientx`, `gradienty` (utils/net_utils.py:243-248) and
`FlowWrapper` (utils/net_utils.py:89-114; the trainer hands it to the loss as self.flowwarp, runners/VAEer.py:53).
The frame loops are TrainingLoss.flowgradloss (sum over t, / t) and TrainingLoss.flowconsist (sum over t).
Inputs are regenerated from tests/synth.py by seed; losses and autograd gradients are stored.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
import synth  # noqa: E402

warnings.filterwarnings("ignore")
from utils.net_utils import FlowWrapper, gradientx, gradienty  # noqa: E402  (the reference, unmodified)


def _flowgradloss(flow, image):  # pyc line 413
    flow = flow * 128
    image = image * 256
    flowgradx = gradientx(flow)
    flowgrady = gradienty(flow)
    imggradx = gradientx(image)
    imggrady = gradienty(image)
    weightx = torch.exp(-torch.mean(torch.abs(imggradx), 1, keepdim=True))
    weighty = torch.exp(-torch.mean(torch.abs(imggrady), 1, keepdim=True))
    lossx = flowgradx * weightx
    lossy = flowgrady * weighty
    return torch.mean(torch.abs(lossx)) + torch.mean(torch.abs(lossy))


def flowgradloss(flow, image, t):
    flow_gradient_loss = 0.0
    for ii in range(t):
        flow_gradient_loss += _flowgradloss(flow[:, :, ii, :, :], image[:, ii, :, :, :])
    return flow_gradient_loss / t


def _flowconsist(flowwarp, flow, flowback, mask_fw=None, mask_bw=None):  # pyc line 481
    if mask_fw is not None:
        prevloss = (mask_bw * torch.abs(flowwarp(flow, -flowback) - flowback)).mean()
        nextloss = (mask_fw * torch.abs(flowwarp(flowback, flow) - flow)).mean()
    else:
        prevloss = torch.abs(flowwarp(flow, -flowback) - flowback).mean()
        nextloss = torch.abs(flowwarp(flowback, flow) - flow).mean()
    return prevloss + nextloss


def flowconsist(flowwarp, flow, flowback, mask_fw, mask_bw, t):
    flowcon = 0.0
    for ii in range(t):
        if mask_bw is not None:
            flowcon += _flowconsist(flowwarp, flow[:, :, ii, :, :], flowback[:, :, ii, :, :], mask_fw=mask_fw[:, ii:ii + 1, ...],
                                    mask_bw=mask_bw[:, ii:ii + 1, ...])
        else:
            flowcon += _flowconsist(flowwarp, flow[:, :, ii, :, :], flowback[:, :, ii, :, :])
    return flowcon


if __name__ == "__main__":
    fw = FlowWrapper()
    for k, (N, T, H, W, C, masks) in enumerate([(2, 3, 24, 40, 3, True), (1, 2, 9, 13, 3, False), (2, 1, 16, 32, 1, True)]):
        s = [90 + 10 * k + q for q in range(6)]
        flow = torch.from_numpy(synth.flow(s[0], N, H, W, 3.0, T=T)).requires_grad_()
        flowback = torch.from_numpy(synth.flow(s[1], N, H, W, 3.0, T=T)).requires_grad_()
        # a smooth image in (-1, 1) (neighbour differences ~ 1/256: the edge weights are neither 0 nor 1)
        image = torch.from_numpy(np.stack([2 * synth.mask(s[2] + 100 * c, N, H, W, T=T) - 1 for c in range(C)], 2))
        m_fw = torch.from_numpy(synth.mask(s[3], N, H, W, T=T)).requires_grad_() if masks else None
        m_bw = torch.from_numpy(synth.mask(s[4], N, H, W, T=T)).requires_grad_() if masks else None
        lg = flowgradloss(flow, image, T)
        lg.backward()
        g_grad = flow.grad.clone()
        flow.grad = None
        lc = flowconsist(fw, flow, flowback, m_fw, m_bw, T)
        lc.backward()
        name = f"losses_{k}.npz"
        out = dict(shape=np.array([N, T, C, H, W]), seeds=np.array(s), masks=np.array(int(masks)), flowgrad=lg.detach().numpy(),
                   flowgrad_gflow=g_grad.numpy(), flowcon=lc.detach().numpy(), flowcon_gflow=flow.grad.numpy(),
                   flowcon_gflowback=flowback.grad.numpy())
        if masks:
            out.update(flowcon_gmfw=m_fw.grad.numpy(), flowcon_gmbw=m_bw.grad.numpy())
        # gradientx / gradienty themselves on one frame
        out.update(gx=gradientx(image[:, 0]).numpy(), gy=gradienty(image[:, 0]).numpy())
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, float(lg), float(lc))
