"""CPU-only tests: the C-ABI library loads and exports every symbol include/flowwarp_b200.h declares, struct
layouts agree between ctypes and the C compiler, host-side validation behaves like the reference's torch
checks, and the oracle satisfies hand-checkable known answers.  No compute call touches a GPU here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from deep_video_interpolation_extrapolation_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "flowwarp_b200.h")).read()
    declared = set(re.findall(r"\b(fwb_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    lib = _lib.load()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.fwb_version() == 0x00010000
    assert lib.fwb_strerror(0) == b"ok"
    assert b"NULL" in lib.fwb_strerror(-1)


def test_struct_layout_matches_c_compiler():
    """sizeof/offsetof as gcc sees the header == what ctypes computes."""
    from deep_video_interpolation_extrapolation_b200 import _lib as L
    src = r'''
#include <stdio.h>
#include "flowwarp_b200.h"
int main(void){
  printf("%zu %zu %zu %zu ", sizeof(fwb_dir), sizeof(fwb_group), sizeof(fwb_problem), sizeof(fwb_grads));
  printf("%zu %zu %zu %zu %zu %zu ", sizeof(fwb_blend), sizeof(fwb_label_problem), offsetof(fwb_blend, grad_noise),
         offsetof(fwb_blend, out), offsetof(fwb_label_problem, labels), offsetof(fwb_label_problem, accumulate));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(fwb_dir, sign), offsetof(fwb_group, out), offsetof(fwb_problem, dir),
         offsetof(fwb_problem, grp), offsetof(fwb_grads, grad_src), offsetof(fwb_grads, grad_blend));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "a.c")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "a")], check=True)
        got = [int(v) for v in subprocess.run([os.path.join(d, "a")], capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(L.fwb_dir), ctypes.sizeof(L.fwb_group), ctypes.sizeof(L.fwb_problem), ctypes.sizeof(L.fwb_grads),
            ctypes.sizeof(L.fwb_blend), ctypes.sizeof(L.fwb_label_problem), L.fwb_blend.grad_noise.offset, L.fwb_blend.out.offset,
            L.fwb_label_problem.labels.offset, L.fwb_label_problem.accumulate.offset,
            L.fwb_dir.sign.offset, L.fwb_group.out.offset, L.fwb_problem.dir.offset, L.fwb_problem.grp.offset,
            L.fwb_grads.grad_src.offset, L.fwb_grads.grad_blend.offset]
    assert got == want


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so the error codes can be checked on a CPU box."""
    from deep_video_interpolation_extrapolation_b200 import _lib as L
    lib = L.load()
    assert lib.fwb_warp_blend_forward(None, None) == -1
    p = L.fwb_problem()
    p.N, p.T, p.H, p.W, p.n_dirs, p.n_groups = 1, 1, 0, 8, 1, 1
    assert lib.fwb_warp_blend_forward(ctypes.byref(p), None) == -2  # empty spatial dim, as torch rejects
    p.H = 8
    p.n_dirs = 3
    assert lib.fwb_warp_blend_forward(ctypes.byref(p), None) == -3
    p.n_dirs, p.n_groups = 1, 9
    assert lib.fwb_warp_blend_forward(ctypes.byref(p), None) == -4
    p.n_groups, p.padding_mode = 1, 7
    assert lib.fwb_warp_blend_forward(ctypes.byref(p), None) == -5
    p.padding_mode = 0
    assert lib.fwb_warp_blend_forward(ctypes.byref(p), None) == -1  # flow pointer NULL
    p.W = 40000
    assert lib.fwb_warp_blend_forward(ctypes.byref(p), None) == -8
    with pytest.raises(ValueError):
        L.check(-2, "x")
    with pytest.raises(RuntimeError):
        L.check(700, "x")


def test_python_wrapper_rejects_cpu_tensors_and_bad_shapes():
    import deep_video_interpolation_extrapolation_b200 as P
    x, f = torch.zeros(2, 3, 8, 8), torch.zeros(2, 2, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors required"):
        P.FlowWrapper()(x, f)
    with pytest.raises(ValueError):
        P.flow_warp_blend([x], [f, f, f])
    with pytest.raises(ValueError):
        P.flow_warp_blend([x], [f], padding_mode="reflection")
    assert list(P.FlowWrapper().parameters()) == [] and list(P.FlowWrapper().buffers()) == []
    assert P.FlowWrapper().state_dict() == {}


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deep_video_interpolation_extrapolation_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("test oracle", "").lower() or f == "_problem.py", f


# ------------------------------------------------------------------ oracle known answers (SURVEY §8c ii)
def test_oracle_known_answers(oracle):
    N, C, H, W = 1, 2, 9, 11
    x = synth.rgb(0, N, H, W, C)
    zero = np.zeros((N, 2, H, W), np.float32)
    assert np.abs(oracle.forward([x], [zero], align_corners=True)[0][:, 0] - x).max() <= 1e-6
    k = 3
    fl = zero.copy()
    fl[:, 0] = 2.0 * k / (W - 1)
    out = oracle.forward([x], [fl], align_corners=True)[0][:, 0]
    assert np.abs(out[..., k:] - x[..., : W - k]).max() <= 1e-5
    assert np.abs(out[..., : k - 1]).max() <= 1e-5
    fl = zero.copy()
    fl[:, 0] = 5.0
    assert np.abs(oracle.forward([x], [fl])[0]).max() == 0.0
    assert np.array_equal(oracle.forward([x], [fl], padding_mode="border", align_corners=True)[0][:, 0], np.broadcast_to(x[..., :1], x.shape))
    fl = zero.copy()
    fl[:, 0] = 1.0 / (W - 1)
    out = oracle.forward([x], [fl], align_corners=True)[0][:, 0]
    assert np.abs(out[..., 1:] - 0.5 * (x[..., 1:] + x[..., :-1])).max() <= 1e-5
    # linspace restatement is bit-equal to torch.linspace on the CPU (utils/net_utils.py:100)
    for n in (2, 3, 7, 128, 150, 257, 1000, 2048):
        ref = torch.linspace(-1, 1, n).numpy()
        mine = np.array([oracle.base_coord(i, n) for i in range(n)], np.float32)
        assert np.array_equal(ref, mine), n
    assert oracle.base_coord(0, 1) == -1.0


@pytest.mark.parametrize("pad,align", [("border", False), ("zeros", False), ("border", True), ("zeros", True)])
def test_oracle_blend_vs_torch_autograd(oracle, pad, align):
    """Bidirectional warp + blend and ALL its gradients against torch autograd of the stock composition."""
    from oracle import torch_ref
    N, H, W = 2, 24, 40
    f0, f1 = [synth.rgb(0, N, H, W), synth.seg(1, N, H, W, 6)], [synth.rgb(10, N, H, W), synth.seg(11, N, H, W, 6)]
    ff, fb, mf, mb = synth.flow(3, N, H, W, 4.0), synth.flow(4, N, H, W, 4.0), synth.mask(2, N, H, W), synth.mask(12, N, H, W)
    gos = [synth.grad(5 + i, a.shape) for i, a in enumerate(f0)]
    T = lambda a: torch.from_numpy(a).clone().requires_grad_()
    a0, a1, aff, afb, amf, amb = [T(a) for a in f0], [T(a) for a in f1], T(ff), T(fb), T(mf), T(mb)
    ref = torch_ref.ref_warp_blend(a0, a1, aff, afb, amf, amb, padding_mode=pad, align_corners=align)
    torch.autograd.backward(ref, [torch.from_numpy(g) for g in gos])
    out = oracle.forward(list(zip(f0, f1)), [ff, fb], blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    g = oracle.backward(list(zip(f0, f1)), [ff, fb], gos, blends=[mf, mb], signs=[-1, 1], padding_mode=pad, align_corners=align)
    rel = lambda a, r: float(np.abs(a - r.detach().numpy()).max() / max(np.abs(r.detach().numpy()).max(), 1e-30))
    for i in range(2):
        assert rel(out[i][:, 0], ref[i]) <= 1e-6
        assert rel(g["grad_srcs"][i][0][:, 0], a0[i].grad) <= 1e-5
        assert rel(g["grad_srcs"][i][1][:, 0], a1[i].grad) <= 1e-5
    assert rel(g["grad_flows"][0][:, :, 0], aff.grad) <= 1e-5
    assert rel(g["grad_flows"][1][:, :, 0], afb.grad) <= 1e-5
    assert rel(g["grad_blends"][0], amf.grad) <= 1e-5
    assert rel(g["grad_blends"][1], amb.grad) <= 1e-5


def test_oracle_threads_agree(oracle):
    N, H, W = 3, 20, 30
    x, fl, go = synth.rgb(0, N, H, W, 4), synth.flow(1, N, H, W, 3.0), synth.grad(2, (N, 4, H, W))
    oracle.set_num_threads(1)
    a, ga = oracle.forward([x], [fl])[0], oracle.backward([x], [fl], [go])
    oracle.set_num_threads(4)
    b, gb = oracle.forward([x], [fl])[0], oracle.backward([x], [fl], [go])
    oracle.set_num_threads(1)
    assert np.array_equal(a, b)
    assert np.array_equal(ga["grad_srcs"][0][0], gb["grad_srcs"][0][0])
