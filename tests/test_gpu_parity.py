"""GPU parity: the CUDA library (through the C-ABI / autograd op) against the CPU oracle, the committed golden
vectors of the unmodified reference, and the stock torch grid_sample composition on the same device.

Bars: sample indices + validity bits bit-exact; forward max|a-b| <= 1e-6*max|ref|;
backward <= 1e-5*max|ref|; deterministic mode bit-exact run to run.
"""
import glob
import os
import types

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
FWD_TOL, BWD_TOL = 1e-6, 1e-5


def relerr(a, ref):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def cu(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().requires_grad_(grad)


@pytest.fixture(scope="module")
def pkg():
    import deep_video_interpolation_extrapolation_b200 as P
    from deep_video_interpolation_extrapolation_b200 import _lib
    _lib.load()  # fail loudly if the CUDA library is missing
    return P


def _reload_env():
    from deep_video_interpolation_extrapolation_b200 import _lib
    _lib.load().fwb_reload_env()


@pytest.fixture(autouse=True)
def _fresh_env():
    """The library reads its A/B environment knobs once per process: every test starts from the current environment."""
    _reload_env()
    yield


def _setenv(monkeypatch, **env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _reload_env()


def _golden(pattern):
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(here, pattern)))


# ---------------------------------------------------------------- golden vectors of the reference
@pytest.mark.parametrize("det", [False, True])
@pytest.mark.parametrize("name", _golden("flowwrapper_*.npz"))
def test_flowwrapper_vs_reference_golden(pkg, golden_dir, name, det):
    z = np.load(os.path.join(golden_dir, name))
    N, C, H, W = z["shape"]
    s = z["seeds"]
    x = cu(synth.rgb(s[0], N, H, W, C), True)
    fl = cu(synth.flow(s[1], N, H, W, float(z["sigma"]), oob_frac=0.05), True)
    go = cu(synth.grad(s[2], (N, C, H, W)))
    out = pkg.FlowWrapper(deterministic=det)(x, fl)
    out.backward(go)
    assert relerr(out, z["out"]) <= FWD_TOL
    assert relerr(x.grad, z["grad_x"]) <= BWD_TOL
    assert relerr(fl.grad, z["grad_flow"]) <= BWD_TOL


@pytest.mark.parametrize("det", [False, True])
def test_warp_and_warp_back_vs_reference_golden(pkg, golden_dir, det):
    fw = pkg.FlowWrapper(deterministic=det)
    z = np.load(os.path.join(golden_dir, "warp_0.npz"))
    N, C, T, H, W = z["shape"]
    s = z["seeds"]
    opt = types.SimpleNamespace(vid_length=int(T))
    x, fl = cu(synth.rgb(s[0], N, H, W, C), True), cu(synth.flow(s[1], N, H, W, 2.0, T=T, oob_frac=0.05), True)
    m = cu(synth.mask(s[2], N, H, W, T=T), True)
    out = pkg.warp(x, fl, opt, fw, m)
    out.backward(cu(synth.grad(s[3], (N, T, C, H, W))))
    assert out.shape == (N, T, C, H, W)
    assert relerr(out, z["out"]) <= FWD_TOL
    assert relerr(x.grad, z["grad_x"]) <= BWD_TOL
    assert relerr(fl.grad, z["grad_flow"]) <= BWD_TOL
    assert relerr(m.grad, z["grad_mask"]) <= BWD_TOL

    z = np.load(os.path.join(golden_dir, "warp_back_0.npz"))
    s = z["seeds"]
    xb = cu(np.stack([synth.rgb(s[0] + i, N, H, W, C) for i in range(T)], 1), True)
    fl, m = cu(synth.flow(s[1], N, H, W, 2.0, T=T, oob_frac=0.05), True), cu(synth.mask(s[2], N, H, W, T=T), True)
    out = pkg.warp_back(xb, fl, opt, fw, m)
    out.backward(cu(synth.grad(s[3], (N, T, C, H, W))))
    assert relerr(out, z["out"]) <= FWD_TOL
    assert relerr(xb.grad, z["grad_x"]) <= BWD_TOL
    assert relerr(fl.grad, z["grad_flow"]) <= BWD_TOL
    assert relerr(m.grad, z["grad_mask"]) <= BWD_TOL


@pytest.mark.parametrize("name", _golden("coordprobe_*.npz"))
def test_coordinate_probe_bit_exact_vs_reference(pkg, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    fl, ref = z["flow"], z["out"]
    M, _, H, W = fl.shape
    if str(z["axis"]) == "x":
        img = (np.arange(W) % 2).astype(np.float32).reshape(1, 1, 1, W).repeat(M, 0)
    else:
        img = (np.arange(H) % 2).astype(np.float32).reshape(1, 1, H, 1).repeat(M, 0)
    out = pkg.FlowWrapper()(cu(img), cu(fl)).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), f"{(out != ref).sum()} of {out.size} differ"


def test_config1_clip_vs_reference_golden(pkg, golden_dir):
    z = np.load(os.path.join(golden_dir, "config1_clip.npz"))
    s = z["seeds"]
    N, H, W = 1, 128, 256
    opt = types.SimpleNamespace(vid_length=1)
    fw = pkg.FlowWrapper()
    f1, f3, s1, s3 = (cu(a) for a in (synth.rgb(s[0], N, H, W), synth.rgb(s[1], N, H, W), synth.seg(s[2], N, H, W), synth.seg(s[3], N, H, W)))
    flf, flb = cu(synth.flow(s[4], N, H, W, 8.0, T=1)), cu(synth.flow(s[5], N, H, W, 8.0, T=1))
    mf, mb = cu(synth.mask(s[6], N, H, W, T=1)), cu(synth.mask(s[7], N, H, W, T=1))
    rgb_f, seg_f = pkg.warp_multi([f1, s1], flf, opt, fw, mf)
    rgb_b = pkg.warp_back(f3[:, None], flb, opt, fw, mb)
    seg_b = pkg.warp_back(s3[:, None], flb, opt, fw, mb)
    for name, o in dict(rgb_f=rgb_f, seg_f=seg_f, rgb_b=rgb_b, seg_b=seg_b).items():
        assert relerr(o[..., ::4, ::4], z[name]) <= FWD_TOL, name


# ---------------------------------------------------------------- indices / validity: bit-exact vs oracle
@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("pad", ["zeros", "border"])
@pytest.mark.parametrize("shape", [(2, 37, 53), (1, 128, 256), (1, 1, 40), (1, 33, 1), (3, 64, 150)])
def test_indices_and_validity_bit_exact(pkg, oracle, shape, pad, align):
    N, H, W = shape
    fl = synth.flow(7, N, H, W, 6.0, oob_frac=0.05)
    g = synth.mask(8, N, H, W)
    for sign in (-1.0, 1.0):
        ref = oracle.sample_indices(fl, gate=g, sign=sign, padding_mode=pad, align_corners=align)
        got = pkg.sample_indices(cu(fl), gate=cu(g), sign=sign, padding_mode=pad, align_corners=align)
        for r, o, what in zip(ref, got, ("x0", "y0", "valid", "ix", "iy")):
            o = o.cpu().numpy()
            if r.dtype == np.float32:
                assert np.array_equal(r.view(np.uint32), o.view(np.uint32)), what
            else:
                assert np.array_equal(r, o), what


def test_indices_match_torch_cuda_grid_sample(pkg):
    """The same parity-image probe as the golden one, against torch's own CUDA grid_sample on this device:
    pins the kernel's coordinate arithmetic to ATen's CUDA binary bit for bit."""
    from oracle import torch_ref
    rng = np.random.default_rng(3)
    for W in (7, 150, 256, 1000, 2047):
        M = 64
        img = cu((np.arange(W) % 2).astype(np.float32).reshape(1, 1, 1, W).repeat(M, 0))
        fl = np.zeros((M, 2, 1, W), np.float32)
        fl[:, 0] = rng.uniform(-1.2, 1.2, (M, 1, W)).astype(np.float32)
        fl[:, 1] = -1.0
        fl = cu(fl)
        for align in (False, True):
            # align_corners=True + zeros + bilinear dispatches to cuDNN's closed-source sampler
            # (torch:include/ATen/native/GridSamplerUtils.h:93-110); the contract is ATen's arithmetic
            with torch.backends.cudnn.flags(enabled=False):
                ref = torch_ref.ref_flow_wrapper(img, fl, align_corners=align)
            out = pkg.FlowWrapper(align_corners=align)(img, fl)
            assert torch.equal(ref, out), (W, align, int((ref != out).sum()))


# ---------------------------------------------------------------- fused op vs oracle and vs stock torch
def _blend_inputs(N, H, W, Cs=(3, 20), sigma=8.0):
    f0 = [synth.rgb(0, N, H, W, Cs[0])] + [synth.seg(1, N, H, W, c) for c in Cs[1:]]
    f1 = [synth.rgb(10, N, H, W, Cs[0])] + [synth.seg(11, N, H, W, c) for c in Cs[1:]]
    ff, fb = synth.flow(3, N, H, W, sigma), synth.flow(4, N, H, W, sigma)
    mf, mb = synth.mask(2, N, H, W), synth.mask(12, N, H, W)
    gos = [synth.grad(5 + i, a.shape) for i, a in enumerate(f0)]
    return f0, f1, ff, fb, mf, mb, gos


@pytest.mark.parametrize("det", [False, True])
@pytest.mark.parametrize("pad,align", [("border", False), ("zeros", False), ("border", True), ("zeros", True)])
def test_warp_blend_vs_oracle(pkg, oracle, pad, align, det):
    N, H, W = 2, 48, 80
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    ref = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
    tff, tfb, tmf, tmb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
    outs = pkg.warp_blend(t0, t1, tff, tfb, tmf, tmb, padding_mode=pad, align_corners=align, deterministic=det)
    torch.autograd.backward(outs, [cu(g) for g in gos])
    for g in range(len(f0)):
        assert relerr(outs[g], ref[g][:, 0]) <= FWD_TOL
        assert relerr(t0[g].grad, rg["grad_srcs"][g][0][:, 0]) <= BWD_TOL
        assert relerr(t1[g].grad, rg["grad_srcs"][g][1][:, 0]) <= BWD_TOL
    assert relerr(tff.grad, rg["grad_flows"][0][:, :, 0]) <= BWD_TOL
    assert relerr(tfb.grad, rg["grad_flows"][1][:, :, 0]) <= BWD_TOL
    assert relerr(tmf.grad, rg["grad_blends"][0]) <= BWD_TOL
    assert relerr(tmb.grad, rg["grad_blends"][1]) <= BWD_TOL


@pytest.mark.parametrize("pad,align", [("border", False), ("zeros", True)])
def test_warp_blend_vs_stock_torch_cuda(pkg, pad, align):
    """Against the reference's own composition of torch ops, run on the same GPU (ATen grid_sampler_2d)."""
    from oracle import torch_ref
    N, H, W = 2, 64, 96
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    a0, a1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
    aff, afb, amf, amb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
    ref = torch_ref.ref_warp_blend(a0, a1, aff, afb, amf, amb, padding_mode=pad, align_corners=align)
    torch.autograd.backward(ref, [cu(g) for g in gos])
    t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
    tff, tfb, tmf, tmb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
    outs = pkg.warp_blend(t0, t1, tff, tfb, tmf, tmb, padding_mode=pad, align_corners=align)
    torch.autograd.backward(outs, [cu(g) for g in gos])
    for g in range(len(f0)):
        assert relerr(outs[g], ref[g]) <= FWD_TOL
        assert relerr(t0[g].grad, a0[g].grad) <= BWD_TOL
        assert relerr(t1[g].grad, a1[g].grad) <= BWD_TOL
    for a, b in ((tff, aff), (tfb, afb), (tmf, amf), (tmb, amb)):
        assert relerr(a.grad, b.grad) <= BWD_TOL


def test_bidirectional_warp_matches_opticalunet_restatement(pkg):
    from oracle import torch_ref
    N, H, W = 2, 40, 72
    f0, f1 = cu(synth.rgb(0, N, H, W)), cu(synth.rgb(1, N, H, W))
    ff, fb = cu(np.tanh(synth.flow(3, N, H, W, 6.0) * 4)), cu(np.tanh(synth.flow(4, N, H, W, 6.0) * 4))
    mf, mb = cu(np.tanh(synth.grad(5, (N, 1, H, W)))), cu(np.tanh(synth.grad(6, (N, 1, H, W))))
    ref = torch_ref.ref_bidirectional(f0, f1, ff, mf, fb, mb)
    got = pkg.bidirectional_warp(f0, f1, ff, mf, fb, mb)
    for r, g in zip(ref, got):
        assert relerr(g, r) <= FWD_TOL


# ---------------------------------------------------------------- known answers (SURVEY.md §8c ii)
def test_known_answers(pkg):
    N, C, H, W = 1, 2, 9, 11
    x = cu(synth.rgb(0, N, H, W, C))
    zero = torch.zeros(N, 2, H, W, device="cuda")
    # zero flow + align_corners=True is the identity
    assert relerr(pkg.FlowWrapper(align_corners=True)(x, zero), x) <= 1e-6
    # integer-pixel flow 2k/(W-1) shifts by k pixels with zero fill (positive flow samples from the left)
    k = 3
    fl = zero.clone()
    fl[:, 0] = 2.0 * k / (W - 1)
    out = pkg.FlowWrapper(align_corners=True)(x, fl)
    assert relerr(out[..., k:], x[..., : W - k]) <= 1e-5
    assert float(out[..., : k - 1].abs().max()) <= 1e-5
    # everything out of bounds: zeros padding -> 0, border padding -> edge replicate
    fl = zero.clone()
    fl[:, 0] = 5.0
    assert float(pkg.FlowWrapper()(x, fl).abs().max()) == 0.0
    out = pkg.FlowWrapper(padding_mode="border", align_corners=True)(x, fl)
    assert torch.equal(out, x[..., :1].expand_as(x))
    # half-pixel flow averages two neighbours
    fl = zero.clone()
    fl[:, 0] = 1.0 / (W - 1)
    out = pkg.FlowWrapper(align_corners=True)(x, fl)
    assert relerr(out[..., 1:], 0.5 * (x[..., 1:] + x[..., :-1])) <= 1e-5


# ---------------------------------------------------------------- edge cases
def test_strided_views_and_channel_counts(pkg, oracle):
    N, T, H, W = 2, 3, 20, 36
    big = synth.flow(1, N, H, W, 3.0, T=T + 2)
    m = synth.mask(2, N, H, W, T=T + 2)
    for C in (1, 3, 20, 23):
        x = synth.rgb(C, N, H, W, C)
        opt = types.SimpleNamespace(vid_length=T)
        out = pkg.warp(cu(x), cu(big), opt, pkg.FlowWrapper(), cu(m))  # flow[:, :, :T] / mask[:, :T] are views
        ref = oracle.forward([x], [big[:, :, :T]], gates=[m[:, :T]])[0]
        assert relerr(out, ref) <= FWD_TOL
    # non-unit W stride falls back to a contiguous copy
    x = synth.rgb(9, N, H, W, 3)
    flt = cu(np.ascontiguousarray(synth.flow(3, N, H, W, 3.0).transpose(0, 1, 3, 2))).transpose(2, 3)
    out = pkg.FlowWrapper()(cu(x), flt)
    ref = oracle.forward([x], [flt.cpu().numpy()])[0][:, 0]
    assert relerr(out, ref) <= FWD_TOL


def test_large_displacement_and_adversarial(pkg, oracle):
    N, H, W = 1, 96, 160
    x, go = synth.rgb(0, N, H, W, 5), synth.grad(1, (N, 5, H, W))
    for fl in (synth.flow(2, N, H, W, 32.0), synth.adversarial_flow(3, N, H, W)):
        for det in (False, True):
            xt, ft = cu(x, True), cu(fl, True)
            out = pkg.FlowWrapper(deterministic=det)(xt, ft)
            out.backward(cu(go))
            ref = oracle.forward([x], [fl])[0][:, 0]
            rg = oracle.backward([x], [fl], [go])
            assert relerr(out, ref) <= FWD_TOL
            assert relerr(xt.grad, rg["grad_srcs"][0][0][:, 0]) <= BWD_TOL
            assert relerr(ft.grad, rg["grad_flows"][0][:, :, 0]) <= BWD_TOL


def test_deterministic_mode_bit_exact_run_to_run(pkg):
    N, H, W = 4, 128, 256
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    runs = []
    for _ in range(3):
        t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
        outs = pkg.warp_blend(t0, t1, cu(ff), cu(fb), cu(mf), cu(mb), deterministic=True)
        torch.autograd.backward(outs, [cu(g) for g in gos])
        runs.append([t.grad.clone() for t in t0 + t1])
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b)


# ---------------------------------------------------------------- every kernel variant stays parity-checked
VARIANTS = ["", "nocl", "sorted", "notex", "notile", "notex,notile", "generic,nofuse", "nofuse"]


@pytest.mark.parametrize("det", [False, True])
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("pad,align", [("border", False), ("zeros", True)])
def test_kernel_variants_vs_oracle(pkg, oracle, monkeypatch, pad, align, variant, det):
    """FWB_KERNELS selects among the CUDA kernels of the library (csrc/flowwarp_b200.cu `knobs`): the generic
    gather kernels, the shared-memory tile kernels, the fused backward and the owner-gather kernel 3.  All of them must
    meet the same bars."""
    _setenv(monkeypatch, FWB_KERNELS=variant)
    N, H, W = 2, 72, 128
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    ref = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
    tff, tfb, tmf, tmb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
    outs = pkg.warp_blend(t0, t1, tff, tfb, tmf, tmb, padding_mode=pad, align_corners=align, deterministic=det)
    torch.autograd.backward(outs, [cu(g) for g in gos])
    for g in range(len(f0)):
        assert relerr(outs[g], ref[g][:, 0]) <= FWD_TOL
        assert relerr(t0[g].grad, rg["grad_srcs"][g][0][:, 0]) <= BWD_TOL
        assert relerr(t1[g].grad, rg["grad_srcs"][g][1][:, 0]) <= BWD_TOL
    assert relerr(tff.grad, rg["grad_flows"][0][:, :, 0]) <= BWD_TOL
    assert relerr(tfb.grad, rg["grad_flows"][1][:, :, 0]) <= BWD_TOL
    assert relerr(tmf.grad, rg["grad_blends"][0]) <= BWD_TOL
    assert relerr(tmb.grad, rg["grad_blends"][1]) <= BWD_TOL


@pytest.mark.parametrize("det", [False, True])
@pytest.mark.parametrize("variant", VARIANTS)
def test_kernel_variants_warp_T_frames_shared_source(pkg, oracle, monkeypatch, variant, det):
    """warp(): one source frame feeds T gated flows; its gradient is summed over the T frames inside the kernels."""
    _setenv(monkeypatch, FWB_KERNELS=variant)
    N, T, H, W = 2, 3, 40, 64
    x = synth.seg(1, N, H, W, 6)
    fl, m = synth.flow(2, N, H, W, 5.0, T=T), synth.mask(3, N, H, W, T=T)
    go = synth.grad(4, (N, T, 6, H, W))
    xt, ft, mt = cu(x, True), cu(fl, True), cu(m, True)
    opt = types.SimpleNamespace(vid_length=T)
    out = pkg.warp(xt, ft, opt, pkg.FlowWrapper(deterministic=det), mt)
    out.backward(cu(go))
    ref = oracle.forward([x], [fl], gates=[m])[0]
    rg = oracle.backward([x], [fl], [go], gates=[m])
    assert relerr(out, ref) <= FWD_TOL
    assert relerr(xt.grad, rg["grad_srcs"][0][0].sum(axis=1)) <= BWD_TOL
    assert relerr(ft.grad, rg["grad_flows"][0]) <= BWD_TOL
    assert relerr(mt.grad, rg["grad_gates"][0]) <= BWD_TOL


@pytest.mark.parametrize("variant", ["", "nofuse"])
def test_kernel_variants_deterministic_bit_exact(pkg, monkeypatch, variant):
    _setenv(monkeypatch, FWB_KERNELS=variant)
    N, H, W = 2, 96, 160
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    runs = []
    for _ in range(3):
        t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
        outs = pkg.warp_blend(t0, t1, cu(ff), cu(fb), cu(mf), cu(mb), deterministic=True)
        torch.autograd.backward(outs, [cu(g) for g in gos])
        runs.append([t.grad.clone() for t in t0 + t1])
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b)


@pytest.mark.parametrize("env", [{}, {"FWB_KERNELS": "notex"}, {"FWB_TILE_BWD_PPT": "1"}, {"FWB_TILE_BWDX_KB": "4"},
                                 {"FWB_KERNELS": "notex", "FWB_TILE_FWD_KB": "8", "FWB_TILE_BWD_KB": "12"}])
@pytest.mark.parametrize("pad", ["border", "zeros"])
@pytest.mark.parametrize("shape", [(2, 37, 52), (1, 16, 32), (1, 130, 260)])
def test_tile_kernels_ragged_shapes_and_fallbacks(pkg, oracle, monkeypatch, shape, pad, env):
    """The default (texture gather + shared-memory scatter) and the shared-memory tile kernels on shapes that are not
    multiples of the tile, with 5 % of the pixels thrown out of the image (slow pixels / tap-less pixels), in the 32x8
    backward variant, and with a shared memory budget so small that every tile takes the in-kernel generic path."""
    _setenv(monkeypatch, **env)
    N, H, W = shape
    f0 = [synth.rgb(0, N, H, W), synth.seg(1, N, H, W, 5)]
    f1 = [synth.rgb(10, N, H, W), synth.seg(11, N, H, W, 5)]
    ff, fb = synth.flow(3, N, H, W, 6.0, oob_frac=0.05), synth.flow(4, N, H, W, 6.0, oob_frac=0.05)
    mf, mb = synth.mask(2, N, H, W), synth.mask(12, N, H, W)
    gos = [synth.grad(5 + i, a.shape) for i, a in enumerate(f0)]
    ref = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode=pad)
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode=pad)
    t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
    tff, tfb, tmf, tmb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
    outs = pkg.warp_blend(t0, t1, tff, tfb, tmf, tmb, padding_mode=pad)
    torch.autograd.backward(outs, [cu(g) for g in gos])
    for g in range(len(f0)):
        assert relerr(outs[g], ref[g][:, 0]) <= FWD_TOL
        assert relerr(t0[g].grad, rg["grad_srcs"][g][0][:, 0]) <= BWD_TOL
        assert relerr(t1[g].grad, rg["grad_srcs"][g][1][:, 0]) <= BWD_TOL
    assert relerr(tff.grad, rg["grad_flows"][0][:, :, 0]) <= BWD_TOL
    assert relerr(tfb.grad, rg["grad_flows"][1][:, :, 0]) <= BWD_TOL
    assert relerr(tmf.grad, rg["grad_blends"][0]) <= BWD_TOL
    assert relerr(tmb.grad, rg["grad_blends"][1]) <= BWD_TOL


def test_fused_backward_propagates_non_finite_grad_out(pkg):
    """The fused backward accumulates in fixed point; a channel pair whose grad_out holds inf/NaN must take the
    exact float path so that the non-finite value reaches grad_src like it does in the reference."""
    N, H, W = 1, 32, 64
    x = cu(synth.rgb(0, N, H, W, 4), True)
    fl = cu(synth.flow(1, N, H, W, 2.0, oob_frac=0.0), True)
    go = torch.ones(N, 4, H, W, device="cuda")
    go[0, 1, 10, 20] = float("inf")
    out = pkg.FlowWrapper()(x, fl)
    out.backward(go)
    g = x.grad
    assert torch.isinf(g[0, 1]).any()
    assert torch.isfinite(g[0, 0]).all() and torch.isfinite(g[0, 2]).all() and torch.isfinite(g[0, 3]).all()


def test_empty_batch_and_errors(pkg):
    x = torch.zeros(0, 3, 8, 8, device="cuda")
    assert pkg.FlowWrapper()(x, torch.zeros(0, 2, 8, 8, device="cuda")).shape == (0, 3, 8, 8)
    x = torch.zeros(2, 3, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        pkg.FlowWrapper()(x, torch.zeros(3, 2, 8, 8, device="cuda"))  # batch mismatch
    with pytest.raises(RuntimeError):
        pkg.FlowWrapper()(x, torch.zeros(2, 3, 8, 8, device="cuda"))  # flow needs 2 channels
    with pytest.raises(RuntimeError):
        pkg.FlowWrapper()(x.cpu(), torch.zeros(2, 2, 8, 8))  # no CPU path
    with pytest.raises(RuntimeError):
        pkg.FlowWrapper()(x.double(), torch.zeros(2, 2, 8, 8, device="cuda").double())


# ---------------------------------------------------------------- full-size, size-independent properties
def test_full_size_properties_config2(pkg):
    """BASELINE config 2 shape (16 x 23 x 256 x 512): properties that need no oracle."""
    N, H, W = 16, 256, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    f0 = [torch.rand(N, 3, H, W, device="cuda", generator=g) * 2 - 1, torch.rand(N, 20, H, W, device="cuda", generator=g)]
    f1 = [torch.rand(N, 3, H, W, device="cuda", generator=g) * 2 - 1, torch.rand(N, 20, H, W, device="cuda", generator=g)]
    ff, fb = cu(synth.flow(3, N, H, W, 8.0)), cu(synth.flow(4, N, H, W, 8.0))
    mf, mb = cu(synth.mask(2, N, H, W)), cu(synth.mask(12, N, H, W))
    out = pkg.warp_blend(f0, f1, ff, fb, mf, mb)
    # linearity in the sources
    out2 = pkg.warp_blend([2 * a for a in f0], [2 * a for a in f1], ff, fb, mf, mb)
    for a, b in zip(out, out2):
        assert relerr(b, 2 * a) <= 1e-6
    # constant images + border padding + masks summing to 1 reproduce the constant
    ones0 = [torch.ones_like(a) for a in f0]
    o = pkg.warp_blend(ones0, ones0, ff, fb, mf, 1 - mf)
    for a in o:
        assert float((a - 1).abs().max()) <= 1e-6
    # adjointness: <warp(x), g> == <x, warp^T(g)>  (the src-gradient kernel is the transpose of the forward)
    x0 = [a.clone().requires_grad_() for a in f0]
    x1 = [a.clone().requires_grad_() for a in f1]
    outs = pkg.warp_blend(x0, x1, ff, fb, mf, mb)
    gos = [torch.randn(a.shape, device="cuda", generator=g) for a in outs]
    torch.autograd.backward(outs, gos)
    lhs = sum(float((o.double() * q.double()).sum()) for o, q in zip(outs, gos))
    rhs = sum(float((x.double() * x.grad.double()).sum()) for x in x0 + x1)
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    # deterministic and atomic paths agree
    y0 = [a.clone().requires_grad_() for a in f0]
    y1 = [a.clone().requires_grad_() for a in f1]
    torch.autograd.backward(pkg.warp_blend(y0, y1, ff, fb, mf, mb, deterministic=True), gos)
    for a, b in zip(x0 + x1, y0 + y1):
        assert relerr(a.grad, b.grad) <= BWD_TOL


# ---------------------------------------------------------------- host-buffer pipeline (the e2e path of bench.py)
@pytest.mark.parametrize("chunk", [1, 2, 5])
def test_host_pipeline_matches_oracle_and_device_op(pkg, oracle, chunk):
    """HostWarpBlend (pinned host in, pinned host out, batch-chunked on three streams) returns what the oracle
    computes; ragged last chunk (N=5, chunk=2) and a chunk larger than the batch included."""
    N, H, W = 5, 40, 64
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    ref = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode="border")
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode="border")
    h = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    pipe = pkg.HostWarpBlend("cuda:0", chunk=chunk)
    for _ in range(2):  # the second call reuses the pinned result buffers
        res = pipe.run([h(a) for a in f0], [h(a) for a in f1], h(ff), h(fb), h(mf), h(mb), [h(g) for g in gos])
    for g in range(len(f0)):
        assert not res["outs"][g].is_cuda
        assert relerr(res["outs"][g], ref[g][:, 0]) <= FWD_TOL
        assert relerr(res["grad_frames0"][g], rg["grad_srcs"][g][0][:, 0]) <= BWD_TOL
        assert relerr(res["grad_frames1"][g], rg["grad_srcs"][g][1][:, 0]) <= BWD_TOL
    assert relerr(res["grad_for_flow"], rg["grad_flows"][0][:, :, 0]) <= BWD_TOL
    assert relerr(res["grad_back_flow"], rg["grad_flows"][1][:, :, 0]) <= BWD_TOL
    assert relerr(res["grad_for_mask"], rg["grad_blends"][0]) <= BWD_TOL
    assert relerr(res["grad_back_mask"], rg["grad_blends"][1]) <= BWD_TOL
    assert pipe.h2d_bytes == sum(a.size * 4 for a in f0 + f1 + [ff, fb, mf, mb] + gos)
    with pytest.raises(RuntimeError):
        pipe.run([cu(a) for a in f0], [h(a) for a in f1], h(ff), h(fb), h(mf), h(mb), [h(g) for g in gos])


@pytest.mark.parametrize("chunk", [1, 2, 8])
def test_host_pipeline_arena_mode_matches_oracle(pkg, oracle, chunk):
    """HostWarpBlend.arena / fill_arena / run_arena: every chunk moves with ONE copy each way out of / into pinned arenas whose
    per-chunk views the producer writes directly; same results as the oracle (ragged last chunk: N=5)."""
    N, H, W = 5, 40, 64
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    ref = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode="border")
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode="border")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    pipe = pkg.HostWarpBlend("cuda:0", chunk=chunk)
    a = pipe.fill_arena([t(x) for x in f0], [t(x) for x in f1], t(ff), t(fb), t(mf), t(mb), [t(g) for g in gos])
    for _ in range(2):
        pipe.run_arena()
    G = len(f0)
    cat = lambda k: torch.cat([views[k] for views in a["results"]], 0)  # noqa: E731  (chunk views -> the whole batch)
    for g in range(G):
        assert relerr(cat(g), ref[g][:, 0]) <= FWD_TOL
        assert relerr(cat(G + g), rg["grad_srcs"][g][0][:, 0]) <= BWD_TOL
        assert relerr(cat(2 * G + g), rg["grad_srcs"][g][1][:, 0]) <= BWD_TOL
    assert relerr(cat(3 * G), rg["grad_flows"][0][:, :, 0]) <= BWD_TOL and relerr(cat(3 * G + 1), rg["grad_flows"][1][:, :, 0]) <= BWD_TOL
    assert relerr(cat(3 * G + 2), rg["grad_blends"][0]) <= BWD_TOL and relerr(cat(3 * G + 3), rg["grad_blends"][1]) <= BWD_TOL
    assert len(a["chunks"]) == (N + min(chunk, N) - 1) // min(chunk, N)


# ---------------------------------------------------------------- grad_src zero-fill fused into the forward
@pytest.mark.parametrize("shape,sigma,T", [((2, 48, 64), 8.0, 1), ((1, 37, 52), 8.0, 1), ((1, 33, 50), 8.0, 1),
                                           ((1, 64, 96), 300.0, 1), ((2, 40, 64), 6.0, 3)])
def test_forward_zero_clears_every_grad_src_plane(pkg, oracle, shape, sigma, T):
    """fwb_warp_blend_forward_zero: same outputs as the plain forward, and every grad_src plane is exactly zero
    afterwards whatever it held (NaN here) - tile path, ragged tiles, W % 4 != 0 (memset path), wild flows (per-tile
    generic path inside the tile kernel), T frames sharing one source (T-stride 0)."""
    import ctypes
    from deep_video_interpolation_extrapolation_b200 import _lib as L
    from deep_video_interpolation_extrapolation_b200._problem import fill_grads, fill_problem
    lib = L.load()
    N, H, W = shape
    Cs = (3, 5)
    five = lambda a: a.unsqueeze(1)  # noqa: E731
    if T == 1:
        srcs = [[five(cu(synth.rgb(10 * g + d, N, H, W, c))) for d in range(2)] for g, c in enumerate(Cs)]
        flows = [cu(synth.flow(3 + d, N, H, W, sigma)).unsqueeze(2) for d in range(2)]
        blends = [cu(synth.mask(2 + d, N, H, W)) for d in range(2)]
        gsrc = [[torch.full((N, 1, c, H, W), float("nan"), device="cuda") for _ in range(2)] for c in Cs]
    else:  # one source frame feeds all T flows: source and grad_src have T-stride 0
        srcs = [[five(cu(synth.rgb(10 * g + d, N, H, W, c))).expand(N, T, c, H, W) for d in range(2)] for g, c in enumerate(Cs)]
        flows = [cu(synth.flow(3 + d, N, H, W, sigma, T=T)) for d in range(2)]
        blends = [cu(synth.mask(2 + d, N, H, W, T=T)) for d in range(2)]
        gsrc = [[torch.full((N, 1, c, H, W), float("nan"), device="cuda").expand(N, T, c, H, W) for _ in range(2)] for c in Cs]
    outs_a = [torch.empty(N, T, c, H, W, device="cuda") for c in Cs]
    outs_b = [torch.empty(N, T, c, H, W, device="cuda") for c in Cs]
    ptr, st = (lambda t: t.data_ptr()), (lambda t: t.stride())
    mk = lambda outs: fill_problem(N=N, T=T, H=H, W=W, flows=flows, gates=[None, None], blends=blends, signs=[-1.0, 1.0],  # noqa: E731
                                   srcs=srcs, outs=outs, padding_mode=L.FWB_PAD_BORDER, align_corners=False,
                                   flags=L.FWB_FLAG_FUSED_BWD, ptr=ptr, strides=st)
    pa, pb = mk(outs_a), mk(outs_b)
    q = fill_grads(pb, grad_outs=[None, None], grad_srcs=gsrc, grad_flows=[None, None], grad_gates=[None, None],
                   grad_blends=[None, None], ptr=ptr, strides=st)
    s = torch.cuda.current_stream().cuda_stream
    L.check(lib.fwb_warp_blend_forward(ctypes.byref(pa), s), "forward")
    L.check(lib.fwb_warp_blend_forward_zero(ctypes.byref(pb), ctypes.byref(q), s), "forward_zero")
    torch.cuda.synchronize()
    for a, b in zip(outs_a, outs_b):
        assert torch.equal(a, b)
    for row in gsrc:
        for g in row:
            assert torch.count_nonzero(g[:, :1] if T > 1 else g).item() == 0 and not torch.isnan(g).any()


def test_backward_twice_and_prezero_switch(pkg, monkeypatch):
    """The zero-filled grad_src buffers of the forward serve ONE backward; a second backward (retain_graph) and the
    PREZERO_GRAD_SRC=False path must give the same gradients."""
    from deep_video_interpolation_extrapolation_b200 import ops
    N, H, W = 2, 48, 64
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    tg = [cu(g) for g in gos]

    def run(twice):
        t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
        outs = pkg.warp_blend(t0, t1, cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True), deterministic=False)
        if twice:
            torch.autograd.backward(outs, tg, retain_graph=True)
            for t in t0 + t1:
                t.grad = None
        torch.autograd.backward(outs, tg)
        return [t.grad.clone() for t in t0 + t1]

    a = run(False)
    b = run(True)
    monkeypatch.setattr(ops, "PREZERO_GRAD_SRC", False)
    c = run(False)
    for x, y, z in zip(a, b, c):
        assert relerr(y, x) <= BWD_TOL and relerr(z, x) <= BWD_TOL


# ---------------------------------------------------------------- refine's mask blend (utils/net_utils.py:131-143)
@pytest.mark.parametrize("shape,Cn", [((2, 3, 23, 32, 64), 3), ((1, 1, 3, 17, 30), 3), ((2, 6, 5, 16, 36), 0),
                                      ((1, 2, 23, 9, 7), 23)])
def test_mask_blend_bit_exact_vs_torch_and_oracle(pkg, oracle, shape, Cn):
    """out = input*mask + noise*(1-mask) per frame: bit-identical to the reference's torch expression (same device)
    and to the numpy oracle; gradients vs fp64 oracle at 1e-5.  Covers W % 4 != 0 (scalar path), T above the register
    accumulators (6 > 4: read-modify-write path), no noise, full-channel noise."""
    N, T, C, H, W = shape
    rng = np.random.default_rng(7)
    inp = rng.standard_normal(shape).astype(np.float32)
    mask = synth.mask(3, N, H, W, T=T)
    noise = rng.standard_normal((N, Cn, H, W)).astype(np.float32) if Cn else None
    go = rng.standard_normal(shape).astype(np.float32)
    ti, tm = cu(inp, True), cu(mask, True)
    tn = cu(noise, True) if Cn else None
    out = pkg.mask_blend(ti, tm, tn)
    # the reference expression, on the same device (utils/net_utils.py:134-136,141-142)
    full = torch.cat([tn.detach(), torch.zeros(N, C - Cn, H, W, device="cuda")], 1) if Cn else torch.zeros(N, C, H, W, device="cuda")
    ref = torch.stack([ti.detach()[:, i] * tm.detach()[:, i:i + 1] + full * (1. - tm.detach()[:, i:i + 1]) for i in range(T)], 1)
    assert torch.equal(out, ref)
    assert np.array_equal(out.detach().cpu().numpy(), oracle.mask_blend_forward(inp, mask, noise))
    out.backward(cu(go))
    gi, gm, gn = oracle.mask_blend_backward(inp, mask, noise, go)
    assert relerr(ti.grad, gi) <= BWD_TOL
    assert relerr(tm.grad, gm) <= BWD_TOL
    if Cn:
        assert relerr(tn.grad, gn) <= BWD_TOL
    # drop-in name of the reference's blend
    assert torch.equal(pkg.blend_with_noise(ti.detach(), tm.detach(), tn.detach() if Cn else None), ref)


def test_mask_blend_strided_views_and_errors(pkg):
    N, T, C, H, W = 2, 2, 4, 8, 16
    big = torch.randn(N, T, C + 2, H, W + 4, device="cuda")
    inp = big[:, :, 1:C + 1, :, 4:]            # strided view, 16-byte aligned rows
    mask = torch.rand(N, T + 1, H, W, device="cuda")  # more frames than input: the first T are used
    noise = torch.randn(N, 3, H, W, device="cuda")
    out = pkg.mask_blend(inp, mask, noise)
    full = torch.cat([noise, torch.zeros(N, C - 3, H, W, device="cuda")], 1)
    ref = torch.stack([inp[:, i] * mask[:, i:i + 1] + full * (1. - mask[:, i:i + 1]) for i in range(T)], 1)
    assert torch.equal(out, ref)
    with pytest.raises(RuntimeError):
        pkg.mask_blend(inp.cpu(), mask, noise)
    with pytest.raises(RuntimeError):
        pkg.mask_blend(inp, mask[:, :1], noise)
    with pytest.raises(RuntimeError):
        pkg.mask_blend(inp, mask, torch.randn(N, C + 1, H, W, device="cuda"))
    assert pkg.mask_blend(inp[:0], mask[:0], noise[:0]).shape[0] == 0


@pytest.mark.parametrize("name", ["refine_0.npz", "refine_1.npz"])
def test_mask_blend_vs_reference_refine_golden(pkg, golden_dir, name):
    """blend_with_noise against the committed outputs of the unmodified reference `refine` (identity refine_net)."""
    z = np.load(os.path.join(golden_dir, name))
    N, T, C, H, W = z["shape"]
    s = z["seeds"]
    ti = cu(synth.grad(s[0], (N, T, C, H, W)), True)
    tm = cu(synth.mask(s[1], N, H, W, T=T), True)
    tn = cu(synth.rgb(s[2], N, H, W, 3), True)
    out = pkg.blend_with_noise(ti, tm, tn)
    assert np.array_equal(out.detach().cpu().numpy(), z["out"])  # bit-exact
    out.backward(cu(synth.grad(s[3], (N, T, C, H, W))))
    assert relerr(ti.grad, z["grad_input"]) <= BWD_TOL
    assert relerr(tm.grad, z["grad_mask"]) <= BWD_TOL
    assert relerr(tn.grad, z["grad_noise"]) <= BWD_TOL


@pytest.mark.parametrize("name", ["refine_0.npz", "refine_1.npz"])
def test_refine_drop_in_signature_vs_reference_golden(pkg, golden_dir, name):
    """refine(input, flow, mask, refine_net, opt, noise_bg) — the reference's own signature (utils/net_utils.py:131-147) with
    an identity refine_net, against the committed outputs of the unmodified reference: bit-exact forward."""
    z = np.load(os.path.join(golden_dir, name))
    N, T, C, H, W = z["shape"]
    s = z["seeds"]
    ti = cu(synth.grad(s[0], (N, T, C, H, W)), True)
    tm = cu(synth.mask(s[1], N, H, W, T=T), True)
    tn = cu(synth.rgb(s[2], N, H, W, 3), True)
    opt = types.SimpleNamespace(vid_length=int(T), seg=bool(C > 3))
    seen = []

    def refine_net(x, flow):  # identity, as in tests/golden/make_golden_refine.py; records the per-frame calls
        seen.append((tuple(x.shape), tuple(flow.shape)))
        return x

    out = pkg.refine(ti, torch.zeros(N, 2, T, H, W, device="cuda"), tm, refine_net, opt, tn)
    assert seen == [((N, C, H, W), (N, 2, H, W))] * T
    assert np.array_equal(out.detach().cpu().numpy(), z["out"])
    out.backward(cu(synth.grad(s[3], (N, T, C, H, W))))
    assert relerr(ti.grad, z["grad_input"]) <= BWD_TOL and relerr(tm.grad, z["grad_mask"]) <= BWD_TOL
    assert relerr(tn.grad, z["grad_noise"]) <= BWD_TOL


@pytest.mark.parametrize("det", [False, True])
def test_warp_cat_equals_cat_of_warps(pkg, det):
    """warp_cat (one launch into one [N,T,3+20,H,W] buffer) == torch.cat([warp(rgb), warp(seg)], dim=2) (nets/VAE_S.py:134-141),
    forward bit for bit, gradients within the bar."""
    N, T, H, W = 2, 3, 40, 64
    rgb, seg = synth.rgb(0, N, H, W, 3), synth.seg(1, N, H, W, 20)
    fl, m = synth.flow(2, N, H, W, 5.0, T=T), synth.mask(3, N, H, W, T=T)
    go = synth.grad(4, (N, T, 23, H, W))
    opt = types.SimpleNamespace(vid_length=T)
    fw = pkg.FlowWrapper(deterministic=det)
    a = [cu(rgb, True), cu(seg, True), cu(fl, True), cu(m, True)]
    ref = torch.cat([pkg.warp(a[0], a[2], opt, fw, a[3]), pkg.warp(a[1], a[2], opt, fw, a[3])], dim=2)
    ref.backward(cu(go))
    b = [cu(rgb, True), cu(seg, True), cu(fl, True), cu(m, True)]
    out = pkg.warp_cat([b[0], b[1]], b[2], opt, fw, b[3])
    assert out.shape == (N, T, 23, H, W) and out.is_contiguous()
    assert torch.equal(out, ref)
    out.backward(cu(go))
    for x, y in zip(a, b):
        assert relerr(y.grad, x.grad) <= BWD_TOL


# ---------------------------------------------------------------- compact segmentation format: uint8 labels
def _labels(seed, N, H, W, K=20, T=None):
    rng = np.random.default_rng(seed)
    shape = (N, (H + 7) // 8, (W + 7) // 8) if T is None else (N, T, (H + 7) // 8, (W + 7) // 8)
    lab = rng.integers(0, K, shape).astype(np.uint8)
    return np.ascontiguousarray(lab.repeat(8, -2).repeat(8, -1)[..., :H, :W])  # blocky map, like synth.seg


@pytest.mark.parametrize("pad,align", [("border", False), ("zeros", False), ("zeros", True)])
@pytest.mark.parametrize("shape", [(2, 48, 80), (1, 37, 53)])
def test_label_warp_blend_equals_dense_on_one_hot(pkg, oracle, shape, pad, align):
    """label_warp_blend(labels) is bit-identical to the dense op on one_hot(labels) (which is pinned to the reference), and
    its flow / mask gradients agree with the dense op's and with the fp64 oracle at the backward bar."""
    N, H, W = shape
    K = 20
    l0, l1 = _labels(21, N, H, W, K), _labels(22, N, H, W, K)
    l0[0, :3, :5] = 200  # labels >= K belong to no class
    ff, fb = synth.flow(3, N, H, W, 8.0, oob_frac=0.05), synth.flow(4, N, H, W, 8.0, oob_frac=0.05)
    mf, mb = synth.mask(2, N, H, W), synth.mask(12, N, H, W)
    go = synth.grad(5, (N, K, H, W))
    oh = lambda l: (np.arange(K)[None, :, None, None] == l[:, None]).astype(np.float32)  # noqa: E731
    A = [cu(a, True) for a in (ff, fb, mf, mb)]
    B = [cu(a, True) for a in (ff, fb, mf, mb)]
    out = pkg.label_warp_blend([cu(l0), cu(l1)], K, A[:2], blends=A[2:], signs=[-1, 1], padding_mode=pad, align_corners=align)
    dense = pkg.flow_warp_blend([(cu(oh(l0)), cu(oh(l1)))], B[:2], blends=B[2:], signs=[-1, 1], padding_mode=pad, align_corners=align)[0]
    assert torch.equal(out, dense)
    ref = oracle.forward([(oh(l0), oh(l1))], [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)[0][:, 0]
    assert relerr(out, ref) <= FWD_TOL
    out.backward(cu(go))
    dense.backward(cu(go))
    rg = oracle.backward([(oh(l0), oh(l1))], [ff, fb], [go], blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    refs = [rg["grad_flows"][0][:, :, 0], rg["grad_flows"][1][:, :, 0], rg["grad_blends"][0], rg["grad_blends"][1]]
    for a, b, r in zip(A, B, refs):
        assert relerr(a.grad, b.grad) <= BWD_TOL
        assert relerr(a.grad, r) <= BWD_TOL


def test_label_warp_T_frames_gated_and_composite(pkg, oracle):
    """`warp`-style use: one label map, T gated flows (T-stride 0 labels), one direction; and warp_blend_labels = RGB through
    the dense kernels + seg through the label kernels with summed flow / mask gradients."""
    N, T, H, W, K = 2, 3, 40, 64, 20
    lab = _labels(31, N, H, W, K)
    fl, gate = synth.flow(41, N, H, W, 4.0, T=T, oob_frac=0.05), synth.mask(42, N, H, W, T=T)
    oh = (np.arange(K)[None, :, None, None] == lab[:, None]).astype(np.float32)
    tf, tg = cu(fl, True), cu(gate, True)
    out = pkg.label_warp_blend(cu(lab), K, tf, gates=tg, signs=-1.0)
    ref = oracle.forward([oh], [fl], gates=[gate])[0]
    assert tuple(out.shape) == (N, T, K, H, W) and relerr(out, ref) <= FWD_TOL
    go = synth.grad(43, (N, T, K, H, W))
    out.backward(cu(go))
    rg = oracle.backward([oh], [fl], [go], gates=[gate])
    assert relerr(tf.grad, rg["grad_flows"][0]) <= BWD_TOL
    assert relerr(tg.grad, rg["grad_gates"][0]) <= BWD_TOL
    # composite
    N, H, W = 2, 48, 64
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    l0, l1 = _labels(51, N, H, W, K), _labels(52, N, H, W, K)
    oh = lambda l: (np.arange(K)[None, :, None, None] == l[:, None]).astype(np.float32)  # noqa: E731
    A = [cu(a, True) for a in (f0[0], f1[0], ff, fb, mf, mb)]
    B = [cu(a, True) for a in (f0[0], f1[0], ff, fb, mf, mb)]
    oa = pkg.warp_blend_labels([A[0]], [A[1]], cu(l0), cu(l1), K, A[2], A[3], A[4], A[5])
    ob = pkg.warp_blend([B[0], cu(oh(l0))], [B[1], cu(oh(l1))], B[2], B[3], B[4], B[5])
    for x, y in zip(oa, ob):
        assert torch.equal(x, y)
    torch.autograd.backward(oa, [cu(g) for g in gos])
    torch.autograd.backward(ob, [cu(g) for g in gos])
    for a, b in zip(A, B):
        assert relerr(a.grad, b.grad) <= BWD_TOL
    with pytest.raises(RuntimeError):
        pkg.label_warp_blend(cu(l0).float(), K, A[2])
    with pytest.raises(RuntimeError):
        pkg.label_warp_blend(cu(l0).cpu(), K, A[2])


# ---------------------------------------------------------------- seeded random sweep of the whole option space
def _random_case(rng):
    N, T = int(rng.integers(1, 3)), int(rng.integers(1, 4))
    H, W = int(rng.integers(1, 70)), int(rng.choice([1, 3, 4, 8, 20, 36, 52, 64, 100, 132]))
    D = int(rng.integers(1, 3))
    Cs = [int(c) for c in rng.integers(1, 7, int(rng.integers(1, 4)))]
    return dict(N=N, T=T, H=H, W=W, D=D, Cs=Cs, sigma=float(rng.choice([0.5, 3.0, 8.0, 40.0])), pad=str(rng.choice(["zeros", "border"])),
                align=bool(rng.integers(0, 2)), gate=bool(rng.integers(0, 2)), blend=bool(rng.integers(0, 2)),
                shared=bool(rng.integers(0, 2)), det=bool(rng.integers(0, 2)), seed=int(rng.integers(0, 1 << 30)))


@pytest.mark.parametrize("k", range(int(os.environ.get("FWB_SWEEP", "24"))))
def test_random_option_sweep_vs_oracle(pkg, oracle, k):
    """Random (seeded) shapes and options through flow_warp_blend, forward and every gradient against the oracle: 1-2
    directions, 1-3 channel groups, T frames with shared or per-frame sources, gates / blends on or off, both padding and
    align modes, deterministic or fused backward, widths that take the tile path (W % 4 == 0) and widths that cannot."""
    c = _random_case(np.random.default_rng(1000 + k))
    N, T, H, W, D = c["N"], c["T"], c["H"], c["W"], c["D"]
    sd = c["seed"]
    signs = [-1.0, 1.0][:D]
    srcs = [[(synth.rgb(sd + 10 * g + d, N, H, W, C)[:, None] if c["shared"] else
              np.stack([synth.rgb(sd + 10 * g + d + 100 * t, N, H, W, C) for t in range(T)], 1)) for d in range(D)]
            for g, C in enumerate(c["Cs"])]
    flows = [synth.flow(sd + 3 + d, N, H, W, c["sigma"], T=T, oob_frac=0.03) for d in range(D)]
    gates = [synth.mask(sd + 5 + d, N, H, W, T=T) if c["gate"] else None for d in range(D)]
    blends = [synth.mask(sd + 7 + d, N, H, W, T=T) if c["blend"] else None for d in range(D)]
    gos = [synth.grad(sd + 20 + g, (N, T, C, H, W)) for g, C in enumerate(c["Cs"])]
    osrc = [tuple(np.broadcast_to(s, (N, T) + s.shape[2:]) for s in row) for row in srcs]
    ref = oracle.forward(osrc, flows, gates=gates, blends=blends, signs=signs, padding_mode=c["pad"], align_corners=c["align"])
    rg = oracle.backward(osrc, flows, gos, gates=gates, blends=blends, signs=signs, padding_mode=c["pad"], align_corners=c["align"])
    ts = [[cu(s, True) for s in row] for row in srcs]
    tf = [cu(f, True) for f in flows]
    tg = [cu(g, True) if g is not None else None for g in gates]
    tb = [cu(b, True) if b is not None else None for b in blends]
    outs = pkg.flow_warp_blend([tuple(r) for r in ts], tf, gates=tg, blends=tb, signs=signs, padding_mode=c["pad"],
                               align_corners=c["align"], deterministic=c["det"])
    torch.autograd.backward(outs, [cu(g) for g in gos])
    for g in range(len(c["Cs"])):
        assert relerr(outs[g], ref[g]) <= FWD_TOL, c
        for d in range(D):
            want = rg["grad_srcs"][g][d]
            if c["shared"] and want.shape[1] != 1:  # the oracle saw T distinct frames (a W == 1 broadcast view gets copied):
                want = want.sum(axis=1, keepdims=True)  # the shared source's gradient is their sum
            assert relerr(ts[g][d].grad, want) <= BWD_TOL, c
    for d in range(D):
        assert relerr(tf[d].grad, rg["grad_flows"][d]) <= BWD_TOL, c
        if c["gate"]:
            assert relerr(tg[d].grad, rg["grad_gates"][d]) <= BWD_TOL, c
        if c["blend"]:
            assert relerr(tb[d].grad, rg["grad_blends"][d]) <= BWD_TOL, c


@pytest.mark.parametrize("which", ["none", "dir0_only", "rgb_only"])
def test_backward_without_some_source_gradients(pkg, oracle, which):
    """Sources that are data (requires_grad=False) get no grad_src buffer: the fused backward then runs the tile kernel
    without scatter / flush for them.  Flow and mask gradients must not change; the requested source gradients neither."""
    N, H, W = 2, 64, 96
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode="border")
    need0 = {"none": [False, False], "dir0_only": [True, True], "rgb_only": [True, False]}[which]
    need1 = {"none": [False, False], "dir0_only": [False, False], "rgb_only": [True, False]}[which]
    t0 = [cu(a, n) for a, n in zip(f0, need0)]
    t1 = [cu(a, n) for a, n in zip(f1, need1)]
    tff, tfb, tmf, tmb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
    outs = pkg.warp_blend(t0, t1, tff, tfb, tmf, tmb)
    torch.autograd.backward(outs, [cu(g) for g in gos])
    assert relerr(tff.grad, rg["grad_flows"][0][:, :, 0]) <= BWD_TOL
    assert relerr(tfb.grad, rg["grad_flows"][1][:, :, 0]) <= BWD_TOL
    assert relerr(tmf.grad, rg["grad_blends"][0]) <= BWD_TOL
    assert relerr(tmb.grad, rg["grad_blends"][1]) <= BWD_TOL
    for g in range(2):
        for d, (t, need) in enumerate(((t0[g], need0[g]), (t1[g], need1[g]))):
            if need:
                assert relerr(t.grad, rg["grad_srcs"][g][d][:, 0]) <= BWD_TOL
            else:
                assert t.grad is None


@pytest.mark.parametrize("det", [True, False])
def test_c_abi_calls_are_cuda_graph_capturable(pkg, det):
    """The entry points only launch kernels / memset nodes on the caller's stream (no allocation, no synchronisation), so a
    forward+backward pair can be captured once and replayed: same bits as the eager launches in deterministic mode, within
    the backward bar for the fused (reduction-order dependent) path."""
    from deep_video_interpolation_extrapolation_b200.host_pipeline import _Slot
    N, H, W = 2, 48, 64
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    kw = dict(padding_mode="border", align_corners=False, deterministic=det, tail=None)
    slot = _Slot(torch.device("cuda:0"), N, (3, 20), H, W, kw)
    for dst, src in zip(slot.inputs(), [*f0, *f1, ff, fb, mf, mb, *gos]):
        dst.view(-1).copy_(cu(src).view(-1))
    slot.launch(N, torch.cuda.current_stream().cuda_stream)  # eager (also warms up cudaFuncSetAttribute)
    torch.cuda.synchronize()
    eager = [t.clone() for t in slot.results()]
    for t in slot.results():
        t.fill_(float("nan"))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        slot.launch(N, torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager, slot.results()):
        assert torch.equal(a, b) if det else relerr(b, a) <= BWD_TOL


def test_caller_expanded_source_keeps_per_frame_gradient(pkg, oracle):
    """A source the CALLER expanded over T ([N,1,C,H,W].expand(-1,T,...), T-stride 0) is an [N,T,...] autograd input: the op
    returns a per-frame gradient and autograd's expand backward sums it (the op used to return [N,1,...])."""
    N, T, H, W = 2, 3, 40, 64
    x = synth.seg(1, N, H, W, 4)
    fl, m = synth.flow(2, N, H, W, 5.0, T=T), synth.mask(3, N, H, W, T=T)
    go = synth.grad(4, (N, T, 4, H, W))
    for det in (False, True):
        xt = cu(x, True)
        xe = xt.unsqueeze(1).expand(-1, T, -1, -1, -1)
        out = pkg.flow_warp_blend([xe], [cu(fl)], gates=[cu(m)], deterministic=det)[0]
        out.backward(cu(go))
        rg = oracle.backward([x], [fl], [go], gates=[m])
        assert relerr(xt.grad, rg["grad_srcs"][0][0].sum(axis=1)) <= BWD_TOL


# ---------------------------------------------------------------- flow-regularisation losses
LOSS_TOL = 2e-6


@pytest.mark.parametrize("name", ["losses_0.npz", "losses_1.npz", "losses_2.npz"])
@pytest.mark.parametrize("det", [False, True])
def test_flow_losses_vs_reference_golden(pkg, golden_dir, name, det):
    """flowgradloss / flowconsist (the library's warp kernels + the fused reductions of csrc/fwb_loss.cuh) against goldens made
    with the reference's own FlowWrapper / gradientx / gradienty composed in the order of the surviving bytecode."""
    from test_oracle_golden import _loss_inputs
    z = np.load(os.path.join(golden_dir, name))
    T, flow, flowback, image, m_fw, m_bw = _loss_inputs(z)
    tf, tb, ti = cu(flow, True), cu(flowback, True), cu(image)
    assert np.array_equal(pkg.gradientx(ti[:, 0]).cpu().numpy(), z["gx"]) and np.array_equal(pkg.gradienty(ti[:, 0]).cpu().numpy(), z["gy"])
    lg = pkg.flowgradloss(tf, ti, T)
    lg.backward()
    assert abs(float(lg) - float(z["flowgrad"])) <= LOSS_TOL * abs(float(z["flowgrad"]))
    assert relerr(tf.grad, z["flowgrad_gflow"]) <= BWD_TOL
    # one frame through the 4-D entry point == the frame loop's term
    l1 = pkg.flow_gradient_loss(cu(flow[:, :, 0]), ti[:, 0])
    ref1 = pkg.flowgradloss(cu(flow[:, :, :1]), ti[:, :1], 1)
    assert float(l1) == float(ref1)
    tf.grad = None
    tm = [None if m is None else cu(m, True) for m in (m_fw, m_bw)]
    fw = pkg.FlowWrapper(deterministic=det)
    lc = pkg.flowconsist(fw, tf, tb, tm[0], tm[1], T)
    lc.backward()
    assert abs(float(lc) - float(z["flowcon"])) <= LOSS_TOL * abs(float(z["flowcon"]))
    assert relerr(tf.grad, z["flowcon_gflow"]) <= BWD_TOL and relerr(tb.grad, z["flowcon_gflowback"]) <= BWD_TOL
    if tm[0] is not None:
        assert relerr(tm[0].grad, z["flowcon_gmfw"]) <= BWD_TOL and relerr(tm[1].grad, z["flowcon_gmbw"]) <= BWD_TOL


def test_flow_losses_are_deterministic_and_reject_cpu(pkg):
    N, T, H, W = 2, 2, 33, 50
    f, img = cu(synth.flow(1, N, H, W, 3.0, T=T), True), cu(np.stack([synth.mask(5 + c, N, H, W, T=T) for c in range(3)], 2))
    a = [float(pkg.flowgradloss(f, img, T)) for _ in range(3)]
    assert a[0] == a[1] == a[2]
    with pytest.raises(RuntimeError):
        pkg.flowgradloss(f.cpu(), img.cpu(), T)


# ---------------------------------------------------------------- the channel-per-lane fused backward (csrc/fwb_cl.cuh)
@pytest.mark.parametrize("Cs,D,T,shape,sigma,pad", [
    ([12], 1, 1, (2, 40, 64), 8.0, "zeros"),       # the narrowest channel set that takes the kernel, one direction
    ([31], 2, 1, (1, 37, 52), 8.0, "border"),      # every lane but the tap counter's, ragged tiles
    ([3, 20], 2, 3, (1, 48, 96), 6.0, "zeros"),    # T frames, shared source (summed gradient), two groups
    ([3, 20], 2, 1, (1, 64, 128), 60.0, "border"),  # large displacements: many slow items (owner-thread scatter)
    ([23], 1, 2, (2, 33, 68), 300.0, "zeros"),     # nearly every tap outside the image
    ([5, 9], 2, 1, (2, 16, 32), 2.0, "border"),    # tiles smaller than 32x8 in both dimensions of the grid
])
def test_channel_lane_backward_vs_oracle(pkg, oracle, Cs, D, T, shape, sigma, pad):
    N, H, W = shape
    signs = [-1.0, 1.0][:D]
    srcs = [[synth.rgb(300 + 10 * g + d, N, H, W, C)[:, None] for d in range(D)] for g, C in enumerate(Cs)]
    flows = [synth.flow(310 + d, N, H, W, sigma, T=T, oob_frac=0.02) for d in range(D)]
    gates = [synth.mask(320 + d, N, H, W, T=T) for d in range(D)]
    blends = [synth.mask(330 + d, N, H, W, T=T) for d in range(D)]
    gos = [synth.grad(340 + g, (N, T, C, H, W)) for g, C in enumerate(Cs)]
    osrc = [tuple(np.broadcast_to(x, (N, T) + x.shape[2:]) for x in row) for row in srcs]
    ref = oracle.forward(osrc, flows, gates=gates, blends=blends, signs=signs, padding_mode=pad)
    rg = oracle.backward(osrc, flows, gos, gates=gates, blends=blends, signs=signs, padding_mode=pad)
    ts = [[cu(x, True) for x in row] for row in srcs]
    tf, tg, tb = [cu(f, True) for f in flows], [cu(g, True) for g in gates], [cu(b, True) for b in blends]
    outs = pkg.flow_warp_blend([tuple(r) for r in ts], tf, gates=tg, blends=tb, signs=signs, padding_mode=pad)
    torch.autograd.backward(outs, [cu(g) for g in gos])
    for g in range(len(Cs)):
        assert relerr(outs[g], ref[g]) <= FWD_TOL
        for d in range(D):
            want = rg["grad_srcs"][g][d]
            assert relerr(ts[g][d].grad, want.sum(axis=1, keepdims=True) if want.shape[1] != 1 else want) <= BWD_TOL
    for d in range(D):
        assert relerr(tf[d].grad, rg["grad_flows"][d]) <= BWD_TOL
        assert relerr(tg[d].grad, rg["grad_gates"][d]) <= BWD_TOL and relerr(tb[d].grad, rg["grad_blends"][d]) <= BWD_TOL


def test_channel_lane_backward_non_finite_and_partial_gradients(pkg, oracle):
    """inf / NaN in grad_out of two of 16 channels: those channels take the exact float path (the non-finite value reaches grad_src as
    in the reference), the others stay finite and correct; and a source gradient wanted for one direction only."""
    N, H, W, C = 1, 40, 64, 16
    x0, x1 = synth.rgb(1, N, H, W, C), synth.rgb(2, N, H, W, C)
    ff, fb = synth.flow(3, N, H, W, 4.0, oob_frac=0.0), synth.flow(4, N, H, W, 4.0, oob_frac=0.0)
    go = synth.grad(5, (N, C, H, W))
    t0, t1 = cu(x0, True), cu(x1, False)  # gradient for direction 0 only
    tff, tfb = cu(ff, True), cu(fb, True)
    gt = cu(go).clone()
    gt[0, 3, 10, 20] = float("inf")
    gt[0, 7, 5, 9] = float("nan")
    out = pkg.flow_warp_blend([(t0, t1)], [tff, tfb], signs=[-1.0, 1.0], padding_mode="border")[0]
    out.backward(gt)
    assert t1.grad is None
    g = t0.grad
    assert torch.isinf(g[0, 3]).any() and torch.isnan(g[0, 7]).any()
    rg = oracle.backward([(x0[:, None], x1[:, None])], [ff[:, :, None], fb[:, :, None]], [go[:, None]], signs=[-1.0, 1.0], padding_mode="border")
    want = rg["grad_srcs"][0][0][:, 0]
    ok = [c for c in range(C) if c not in (3, 7)]
    assert torch.isfinite(g[0, ok]).all()
    assert relerr(g[:, ok], want[:, ok]) <= BWD_TOL


def test_fused_backward_grad_out_views_not_16_byte_aligned(pkg, oracle):
    """The channel-per-lane backward stages the grad_out tile with 16-byte copies when every plane row is 16-byte aligned and
    with 4-byte copies otherwise: upstream gradients handed over as column-shifted views of wider buffers (row stride W + 4,
    first element 4 bytes off) must give the gradients of the dense tensors, ragged last tile row included."""
    N, H, W = 2, 60, 96
    f0, f1, ff, fb, mf, mb, gos = _blend_inputs(N, H, W)
    rg = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode="border")
    for shifted in (False, True):
        t0, t1 = [cu(a, True) for a in f0], [cu(a, True) for a in f1]
        tff, tfb, tmf, tmb = cu(ff, True), cu(fb, True), cu(mf, True), cu(mb, True)
        outs = pkg.warp_blend(t0, t1, tff, tfb, tmf, tmb)
        if shifted:
            tg = []
            for g in gos:
                wide = torch.full(g.shape[:-1] + (W + 4,), float("nan"), device="cuda")
                wide[..., 1:W + 1] = cu(g)
                tg.append(wide[..., 1:W + 1])
                assert tg[-1].data_ptr() % 16 != 0 and tg[-1].stride(-1) == 1
        else:
            tg = [cu(g) for g in gos]
        torch.autograd.backward(outs, tg)
        for g in range(2):
            assert relerr(t0[g].grad, rg["grad_srcs"][g][0][:, 0]) <= BWD_TOL, shifted
            assert relerr(t1[g].grad, rg["grad_srcs"][g][1][:, 0]) <= BWD_TOL, shifted
        assert relerr(tff.grad, rg["grad_flows"][0][:, :, 0]) <= BWD_TOL, shifted
        assert relerr(tmb.grad, rg["grad_blends"][1]) <= BWD_TOL, shifted
