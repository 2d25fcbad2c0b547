"""The CPU oracle (oracle/flowwarp_oracle.c) and the torch restatement (oracle/torch_ref.py) against the golden
vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py; utils/net_utils.py:89-129).

Tolerances (norm-relative): forward max|a-b| <= 1e-6*max|ref|, backward <= 1e-5*max|ref|.
The coordinate probe is bit-exact.
"""
import glob
import os

import numpy as np
import pytest
import torch

import synth
from oracle import torch_ref

FWD_TOL, BWD_TOL = 1e-6, 1e-5


def relerr(a, ref):
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def _cases(pattern):
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(here, pattern)))


@pytest.mark.parametrize("name", _cases("flowwrapper_*.npz"))
def test_flowwrapper_golden(oracle, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    N, C, H, W = z["shape"]
    s = z["seeds"]
    x, fl, go = synth.rgb(s[0], N, H, W, C), synth.flow(s[1], N, H, W, float(z["sigma"]), oob_frac=0.05), synth.grad(s[2], (N, C, H, W))
    out = oracle.forward([x], [fl])[0][:, 0]
    assert relerr(out, z["out"]) <= FWD_TOL
    g = oracle.backward([x], [fl], [go])
    assert relerr(g["grad_srcs"][0][0][:, 0], z["grad_x"]) <= BWD_TOL
    assert relerr(g["grad_flows"][0][:, :, 0], z["grad_flow"]) <= BWD_TOL
    # torch restatement: same calls as the reference -> identical bits on the same torch build
    xt, ft = torch.from_numpy(x).requires_grad_(), torch.from_numpy(fl).requires_grad_()
    o2 = torch_ref.ref_flow_wrapper(xt, ft)
    o2.backward(torch.from_numpy(go))
    assert relerr(o2.detach().numpy(), z["out"]) <= 1e-7
    assert relerr(xt.grad.numpy(), z["grad_x"]) <= BWD_TOL
    assert relerr(ft.grad.numpy(), z["grad_flow"]) <= BWD_TOL


def test_warp_golden(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "warp_0.npz"))
    N, C, T, H, W = z["shape"]
    s = z["seeds"]
    x, fl, m = synth.rgb(s[0], N, H, W, C), synth.flow(s[1], N, H, W, 2.0, T=T, oob_frac=0.05), synth.mask(s[2], N, H, W, T=T)
    go = synth.grad(s[3], (N, T, C, H, W))
    out = oracle.forward([x], [fl], gates=[m])[0]
    assert relerr(out, z["out"]) <= FWD_TOL
    g = oracle.backward([x], [fl], [go], gates=[m])
    assert relerr(g["grad_srcs"][0][0][:, 0], z["grad_x"]) <= BWD_TOL  # summed over the T frames sharing x
    assert relerr(g["grad_flows"][0], z["grad_flow"]) <= BWD_TOL
    assert relerr(g["grad_gates"][0], z["grad_mask"]) <= BWD_TOL
    o2 = torch_ref.ref_warp(torch.from_numpy(x), torch.from_numpy(fl), T, torch.from_numpy(m))
    assert relerr(o2.numpy(), z["out"]) <= 1e-7


def test_warp_back_golden(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "warp_back_0.npz"))
    N, C, T, H, W = z["shape"]
    s = z["seeds"]
    xb = np.stack([synth.rgb(s[0] + i, N, H, W, C) for i in range(T)], 1)
    fl, m = synth.flow(s[1], N, H, W, 2.0, T=T, oob_frac=0.05), synth.mask(s[2], N, H, W, T=T)
    go = synth.grad(s[3], (N, T, C, H, W))
    out = oracle.forward([xb], [fl], gates=[m], signs=+1.0)[0]
    assert relerr(out, z["out"]) <= FWD_TOL
    g = oracle.backward([xb], [fl], [go], gates=[m], signs=+1.0)
    assert relerr(g["grad_srcs"][0][0], z["grad_x"]) <= BWD_TOL
    assert relerr(g["grad_flows"][0], z["grad_flow"]) <= BWD_TOL
    assert relerr(g["grad_gates"][0], z["grad_mask"]) <= BWD_TOL
    o2 = torch_ref.ref_warp_back(torch.from_numpy(xb), torch.from_numpy(fl), T, torch.from_numpy(m))
    assert relerr(o2.numpy(), z["out"]) <= 1e-7


@pytest.mark.parametrize("name", _cases("coordprobe_*.npz"))
def test_coordinates_bit_exact_vs_reference(oracle, golden_dir, name):
    """Parity image with the other axis of size 1: the reference's bilinear output is frac(ix) or 1-frac(ix)
    exactly, so equality of the outputs pins linspace base + subtraction + unnormalisation bit for bit."""
    z = np.load(os.path.join(golden_dir, name))
    fl, ref = z["flow"], z["out"]
    M, _, H, W = fl.shape
    if str(z["axis"]) == "x":
        img = (np.arange(W) % 2).astype(np.float32).reshape(1, 1, 1, W).repeat(M, 0)
    else:
        img = (np.arange(H) % 2).astype(np.float32).reshape(1, 1, H, 1).repeat(M, 0)
    out = oracle.forward([img], [fl])[0][:, 0]
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), f"{(out != ref).sum()} of {out.size} differ"


def test_config1_clip_golden(oracle, golden_dir):
    """BASELINE config 1: one 3-frame 128x256 clip, RGB + 20-class seg, batch 1, forward (CPU)."""
    z = np.load(os.path.join(golden_dir, "config1_clip.npz"))
    s = z["seeds"]
    N, H, W = 1, 128, 256
    f1, f3, s1, s3 = synth.rgb(s[0], N, H, W), synth.rgb(s[1], N, H, W), synth.seg(s[2], N, H, W), synth.seg(s[3], N, H, W)
    flf, flb = synth.flow(s[4], N, H, W, 8.0, T=1), synth.flow(s[5], N, H, W, 8.0, T=1)
    mf, mb = synth.mask(s[6], N, H, W, T=1), synth.mask(s[7], N, H, W, T=1)
    rgb_f, seg_f = oracle.forward([f1, s1], [flf], gates=[mf])  # both groups in one call
    rgb_b, seg_b = oracle.forward([f3[:, None], s3[:, None]], [flb], gates=[mb], signs=+1.0)
    for name, o in dict(rgb_f=rgb_f, seg_f=seg_f, rgb_b=rgb_b, seg_b=seg_b).items():
        assert relerr(o[..., ::4, ::4], z[name]) <= FWD_TOL, name
        ssum = o.astype(np.float64).sum((0, 1, 3, 4))
        assert np.abs(ssum - z[name + "_sum"]).max() <= 1e-6 * np.abs(o).sum() / o.shape[2], name


@pytest.mark.parametrize("name", ["refine_0.npz", "refine_1.npz"])
def test_refine_blend_golden(oracle, golden_dir, name):
    """oracle.mask_blend_* against the unmodified reference `refine` (identity refine_net), utils/net_utils.py:131-150."""
    z = np.load(os.path.join(golden_dir, name))
    N, T, C, H, W = z["shape"]
    s = z["seeds"]
    inp, mask, noise = synth.grad(s[0], (N, T, C, H, W)), synth.mask(s[1], N, H, W, T=T), synth.rgb(s[2], N, H, W, 3)
    go = synth.grad(s[3], (N, T, C, H, W))
    assert np.array_equal(oracle.mask_blend_forward(inp, mask, noise), z["out"])  # bit-exact
    gi, gm, gn = oracle.mask_blend_backward(inp, mask, noise, go)
    assert relerr(gi, z["grad_input"]) <= BWD_TOL
    assert relerr(gm, z["grad_mask"]) <= BWD_TOL
    assert relerr(gn, z["grad_noise"]) <= BWD_TOL


@pytest.mark.parametrize("pad,align", [("border", False), ("zeros", True)])
def test_flat_entry_matches_struct_entry_and_torch(oracle, pad, align):
    """fwo_bidir_contig (problem packed in C from shapes: the full-size GPU tests' checker) == the struct-packed oracle entry
    bit for bit, and == the torch restatement of nets/OpticalUnet.py:123-146 within the parity bars."""
    N, H, W = 2, 40, 56
    f0 = [synth.rgb(0, N, H, W, 3), synth.seg(1, N, H, W, 5)]
    f1 = [synth.rgb(10, N, H, W, 3), synth.seg(11, N, H, W, 5)]
    ff, fb = synth.flow(3, N, H, W, 6.0), synth.flow(4, N, H, W, 6.0)
    mf, mb = synth.mask(2, N, H, W), synth.mask(12, N, H, W)
    gos = [synth.grad(5 + i, a.shape) for i, a in enumerate(f0)]
    flat = oracle.bidir_contig(f0, f1, ff, fb, mf[:, 0], mb[:, 0], grad_outs=gos, padding_mode=pad, align_corners=align)
    ref = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    for g in range(2):
        assert np.array_equal(flat["out"][g], ref[g][:, 0])
        assert np.array_equal(flat["gsrc0"][g], rg["grad_srcs"][g][0][:, 0])
        assert np.array_equal(flat["gsrc1"][g], rg["grad_srcs"][g][1][:, 0])
    assert np.array_equal(flat["gflow0"], rg["grad_flows"][0][:, :, 0]) and np.array_equal(flat["gflow1"], rg["grad_flows"][1][:, :, 0])
    assert np.array_equal(flat["gblend0"], rg["grad_blends"][0][:, 0]) and np.array_equal(flat["gblend1"], rg["grad_blends"][1][:, 0])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).requires_grad_()
    a0, a1 = [t(a) for a in f0], [t(a) for a in f1]
    tff, tfb, tmf, tmb = t(ff), t(fb), t(mf), t(mb)
    outs = torch_ref.ref_warp_blend(a0, a1, tff, tfb, tmf, tmb, padding_mode=pad, align_corners=align)
    torch.autograd.backward(outs, [torch.from_numpy(g) for g in gos])
    for g in range(2):
        assert relerr(flat["out"][g], outs[g].detach().numpy()) <= 1e-6
        assert relerr(flat["gsrc0"][g], a0[g].grad.numpy()) <= 1e-5 and relerr(flat["gsrc1"][g], a1[g].grad.numpy()) <= 1e-5
    assert relerr(flat["gflow0"], tff.grad.numpy()) <= 1e-5 and relerr(flat["gblend1"], tmb.grad.numpy()[:, 0]) <= 1e-5


# ---------------------------------------------------------------- flow-regularisation losses
def _loss_inputs(z):
    N, T, C, H, W = z["shape"]
    s = z["seeds"]
    flow = synth.flow(s[0], N, H, W, 3.0, T=T)
    flowback = synth.flow(s[1], N, H, W, 3.0, T=T)
    image = np.stack([2 * synth.mask(s[2] + 100 * c, N, H, W, T=T) - 1 for c in range(C)], 2)
    masks = bool(z["masks"])
    m_fw = synth.mask(s[3], N, H, W, T=T) if masks else None
    m_bw = synth.mask(s[4], N, H, W, T=T) if masks else None
    return int(T), flow, flowback, image, m_fw, m_bw


@pytest.mark.parametrize("name", ["losses_0.npz", "losses_1.npz", "losses_2.npz"])
def test_flow_losses_restatement_vs_reference_golden(golden_dir, name):
    """oracle/torch_ref.py's restatement of TrainingLoss.flowgradloss / flowconsist against the goldens made with the reference's
    own FlowWrapper / gradientx / gradienty (tests/golden/make_golden_losses.py)."""
    z = np.load(os.path.join(golden_dir, name))
    T, flow, flowback, image, m_fw, m_bw = _loss_inputs(z)
    tf, tb = torch.from_numpy(flow).requires_grad_(), torch.from_numpy(flowback).requires_grad_()
    ti = torch.from_numpy(image)
    assert np.array_equal(torch_ref.ref_gradientx(ti[:, 0]).numpy(), z["gx"]) and np.array_equal(torch_ref.ref_gradienty(ti[:, 0]).numpy(), z["gy"])
    lg = torch_ref.ref_flowgradloss(tf, ti, T)
    lg.backward()
    assert abs(float(lg) - float(z["flowgrad"])) <= 1e-6 * abs(float(z["flowgrad"]))
    assert np.abs(tf.grad.numpy() - z["flowgrad_gflow"]).max() <= 1e-6 * np.abs(z["flowgrad_gflow"]).max()
    tf.grad = None
    tm = [None if m is None else torch.from_numpy(m).requires_grad_() for m in (m_fw, m_bw)]
    lc = torch_ref.ref_flowconsist(tf, tb, tm[0], tm[1], T)
    lc.backward()
    assert abs(float(lc) - float(z["flowcon"])) <= 1e-6 * abs(float(z["flowcon"]))
    assert np.abs(tf.grad.numpy() - z["flowcon_gflow"]).max() <= 1e-5 * np.abs(z["flowcon_gflow"]).max()
    assert np.abs(tb.grad.numpy() - z["flowcon_gflowback"]).max() <= 1e-5 * np.abs(z["flowcon_gflowback"]).max()
