"""Full-size parity gate: the CUDA library against the C oracle (all host cores) at the BASELINE.json config shapes, on the
very inputs bench.py times (bench.make_inputs, including the 1 % out-of-bounds kick), fused and deterministic modes, and
config 4's full batch against the stock torch CUDA composition.  The oracle's problem description is packed in C
(oracle.bidir_contig -> fwo_bidir_contig), not by the product's struct packer.  Max errors are printed (pytest -s).

Bars: forward max|a-b| <= 1e-6 max|ref|; backward <= 1e-5 max|ref| (utils/net_utils.py:93-129 through ATen grid_sampler_2d).
"""
import os

import numpy as np
import pytest
import torch

import bench

pytestmark = pytest.mark.gpu
FWD_TOL, BWD_TOL = 1e-6, 1e-5


def rel(a, ref):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref
    return float(np.abs(a.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def pkg():
    import deep_video_interpolation_extrapolation_b200 as P
    from deep_video_interpolation_extrapolation_b200 import _lib
    _lib.load()
    return P


def _run_cuda(pkg, inp, det):
    leaves = [t.detach().clone().requires_grad_() for t in inp["f0"] + inp["f1"] + [inp["ff"], inp["fb"], inp["mf"], inp["mb"]]]
    G = len(inp["f0"])
    f0, f1, (ff, fb, mf, mb) = leaves[:G], leaves[G:2 * G], leaves[2 * G:]
    outs = pkg.warp_blend(f0, f1, ff, fb, mf, mb, padding_mode="border", deterministic=det)
    torch.autograd.backward(outs, inp["gos"])
    torch.cuda.synchronize()
    return outs, leaves


def _check_vs_oracle(pkg, oracle, cfg_id, n_clips, label):
    cfg = dict(bench.CONFIGS[cfg_id], N=n_clips)
    inp = bench.make_inputs(cfg, torch.device("cuda"), seed=0)
    oracle.set_num_threads(os.cpu_count() or 8)
    h = lambda t: t.detach().cpu().numpy()
    ref = oracle.bidir_contig([h(a) for a in inp["f0"]], [h(a) for a in inp["f1"]], h(inp["ff"]), h(inp["fb"]),
                              h(inp["mf"])[:, 0], h(inp["mb"])[:, 0], grad_outs=[h(g) for g in inp["gos"]], padding_mode="border")
    G = len(inp["f0"])
    worst = {}
    for det in (False, True):
        outs, leaves = _run_cuda(pkg, inp, det)
        e = {"fwd": max(rel(outs[g], ref["out"][g]) for g in range(G)),
             "gsrc": max(max(rel(leaves[g].grad, ref["gsrc0"][g]), rel(leaves[G + g].grad, ref["gsrc1"][g])) for g in range(G)),
             "gflow": max(rel(leaves[2 * G].grad, ref["gflow0"]), rel(leaves[2 * G + 1].grad, ref["gflow1"])),
             "gmask": max(rel(leaves[2 * G + 2].grad[:, 0], ref["gblend0"]), rel(leaves[2 * G + 3].grad[:, 0], ref["gblend1"]))}
        print(f"[full-size parity] {label} {'deterministic' if det else 'fused'}: " + ", ".join(f"{k} {v:.2e}" for k, v in e.items()))
        worst[det] = e
        del outs, leaves
    for det, e in worst.items():
        assert e["fwd"] <= FWD_TOL, (det, e)
        assert e["gsrc"] <= BWD_TOL and e["gflow"] <= BWD_TOL and e["gmask"] <= BWD_TOL, (det, e)


def test_config2_full_batch_vs_oracle(pkg, oracle):
    """BASELINE configs[1]: 16 x (3+20) x 256 x 512, sigma 8 px — the exact tensors the headline bench line times."""
    _check_vs_oracle(pkg, oracle, 2, 16, "config2 16x23x256x512")


def test_config3_one_step_vs_oracle(pkg, oracle):
    """BASELINE configs[2]: one step of the chain, 8 x (3+20) x 512 x 1024."""
    _check_vs_oracle(pkg, oracle, 3, 8, "config3 8x23x512x1024")


def test_config4_one_clip_vs_oracle(pkg, oracle):
    """BASELINE configs[3]: full resolution 1024 x 2048, sigma 32 px (slow pixels, generic tiles, large offsets): 1 clip."""
    _check_vs_oracle(pkg, oracle, 4, 1, "config4 1x23x1024x2048 sigma32")


def test_config4_full_batch_vs_stock_torch_cuda(pkg):
    """Config 4 at its full per-GPU batch (4 clips) against the reference's own torch composition on the same GPU."""
    cfg = bench.CONFIGS[4]
    inp = bench.make_inputs(cfg, torch.device("cuda"), seed=0)
    ref_outs, ref_leaves = bench.stock_torch_gpu_step(inp, return_leaves=True)
    G = len(inp["f0"])
    for det in (False, True):
        outs, leaves = _run_cuda(pkg, inp, det)
        e = {"fwd": max(rel(outs[g], ref_outs[g]) for g in range(G)),
             "gsrc": max(rel(leaves[i].grad, ref_leaves[i].grad) for i in range(2 * G)),
             "gflow": max(rel(leaves[2 * G + i].grad, ref_leaves[2 * G + i].grad) for i in range(2)),
             "gmask": max(rel(leaves[2 * G + 2 + i].grad, ref_leaves[2 * G + 2 + i].grad) for i in range(2))}
        print(f"[full-size parity] config4 4x23x1024x2048 vs torch CUDA, {'deterministic' if det else 'fused'}: "
              + ", ".join(f"{k} {v:.2e}" for k, v in e.items()))
        assert e["fwd"] <= FWD_TOL, e
        assert e["gsrc"] <= BWD_TOL and e["gflow"] <= BWD_TOL and e["gmask"] <= BWD_TOL, e
        del outs, leaves
