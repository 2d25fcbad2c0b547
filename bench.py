#!/usr/bin/env python
"""bench.py — warp+blend forward+backward throughput (Gpix/s) and HBM-roofline fraction on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of synthetic Cityscapes-shaped clips: the fused
bidirectional warp + mask-weighted blend forward (kernel 1), the flow/mask gradient (kernel 2) and the
source-gradient (kernel 3), C = 3 RGB + 20 seg channels, per-GPU batch fixed (weak scaling, batch-sharded,
no collective in the op).  Default workload = BASELINE.json configs[1]: 256x512, batch 16, 1xB200.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (N per GPU, H, W, sigma_px, chained steps, allreduce params)
    1: dict(name="config1: InterNet flow-warp forward, one 3-frame clip 128x256, batch 1 (the reference's CPU-runnable case)", N=1, H=128,
            W=256, sigma=8.0, chain=1, allreduce=0, fwd_only=True),
    2: dict(name="config2: InterNet warp+blend fwd+bwd 256x512 batch 16", N=16, H=256, W=512, sigma=8.0, chain=1, allreduce=0),
    3: dict(name="config3: ExtraNet 3-step chained warp 512x1024 batch 8", N=8, H=512, W=1024, sigma=8.0, chain=3, allreduce=0),
    4: dict(name="config4: int_9 full-res 1024x2048 batch 4 per GPU", N=4, H=1024, W=2048, sigma=32.0, chain=1, allreduce=0),
    5: dict(name="config5: InterGAN/refine step 256x512 batch 8 per GPU + NCCL grad all-reduce", N=8, H=256, W=512, sigma=8.0, chain=1,
            allreduce=3_821_891),
}
CH = (3, 20)  # RGB + 20-class seg
C_TOTAL = sum(CH)
BYTES_PER_PIX = 32 * C_TOTAL + 72  # bidirectional warp+blend fwd+bwd, fp32
# per-kernel algorithmic bytes/pixel when the three kernels run as separate launches
KERNEL_BYTES = {"forward": 12 * C_TOTAL + 24, "backward_flow": 12 * C_TOTAL + 48, "backward_src": 12 * C_TOTAL + 24,
                "backward_fused": 20 * C_TOTAL + 48}


def config_dict(cfg, args, world):
    """The `config` object of the JSON line: identical keys in the b200 and the reference arm."""
    d = {"workload": cfg["name"], "per_gpu_batch": cfg["N"], "H": cfg["H"], "W": cfg["W"], "channels": list(CH),
            "flow_sigma_px": cfg["sigma"], "chained_steps": cfg["chain"], "padding_mode": "border", "align_corners": False,
            "forward_only": bool(cfg.get("fwd_only", False)),
            "parallelism": f"batch-sharded x{world}, no collective in the op",
            "l2": "inputs+outputs per step (~1.2 GB at config 2) exceed the 126 MB L2; no explicit flush"}
    fused = not (args.deterministic or getattr(args, "atomic_src", False))
    d.update({"deterministic": bool(args.deterministic),
              "grad_src_zeroing": args.zero if (fused and not cfg.get("fwd_only", False)) else "memset"})
    if cfg["chain"] > 1:
        d["chain"] = ("true K-step chain through the autograd op: prediction k is a source of step k+1, seg re-one-hotted "
                      "by argmax, gradient through the RGB prediction, one backward (runners/ExtraTrainer.py:254-310)")
    if cfg["allreduce"]:
        d["allreduce"] = f"{cfg['allreduce']} fp32 parameters, dist.all_reduce(async_op=True) issued before the op, joined after it"
    return d


def ncu_traffic(kernel, cfg_id):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum); None when no capture exists for this kernel and workload."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(f"config{cfg_id}", {}).get(kernel)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------- inputs
def make_inputs(cfg, device, seed=0):
    """Seeded synthetic clip batch generated with torch on `device`."""
    import torch
    import torch.nn.functional as F
    N, H, W, sig = cfg["N"], cfg["H"], cfg["W"], cfg["sigma"]
    g = torch.Generator(device="cpu").manual_seed(seed)

    def smooth(ch):
        coarse = torch.randn(N, ch, H // 16 + 2, W // 16 + 2, generator=g).to(device)
        return F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True)

    def flow():
        f = smooth(2)
        f[:, 0] *= 2.0 * sig / W
        f[:, 1] *= 2.0 * sig / H
        kick = (torch.rand(N, H, W, generator=g) < 0.01).to(device)  # >= 1% of pixels out of bounds
        f[:, 0] = torch.where(kick, f[:, 0] + 2.5, f[:, 0])
        return f.contiguous()

    def seg():
        lab = torch.randint(0, CH[1], (N, 1, (H + 7) // 8, (W + 7) // 8), generator=g).float().to(device)
        lab = F.interpolate(lab, size=(H, W), mode="nearest").long()
        return torch.zeros(N, CH[1], H, W, device=device).scatter_(1, lab, 1.0)

    def rgb():
        return (torch.rand(N, CH[0], H, W, generator=g) * 2 - 1).to(device)

    return dict(f0=[rgb(), seg()], f1=[rgb(), seg()], ff=flow(), fb=flow(), mf=torch.sigmoid(smooth(1)),
                mb=torch.sigmoid(smooth(1)), gos=[torch.randn(N, c, H, W, generator=g).to(device) for c in CH])


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/fwb_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.marks = {}

    def mark(self, name):
        self.marks[name] = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        import datetime
        rows = []
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(p[1]), float(p[2]), float(p[3]), p[4], p[5], p[6], p[7], p[8]))
            except Exception:
                continue
        os.unlink(self.path)
        t0, t1 = self.marks.get("timed_start", 0), self.marks.get("timed_end", 1e18)
        sel = [r for r in rows if t0 <= r[0] <= t1]
        window = "timed region"
        if len(sel) < 3:  # short timed region: widen to every sample taken while this bench was launching kernels
            t0, t1 = self.marks.get("load_start", 0), self.marks.get("load_end", 1e18)
            sel = [r for r in rows if t0 <= r[0] <= t1]
            window = "warm-up + timed + per-kernel passes (timed region shorter than 3 samples)"
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        reasons = set()
        for r in sel:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[1] for r in sel), "sm_max_mhz": sel[0][2],
                "power_w_max": max(r[3] for r in sel), "reasons": sorted(reasons), "samples": len(sel), "window": window}


# --------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_reference_step(inp, pad="border", fwd_only=False):
    """The reference's own composition of torch ops (oracle/torch_ref.py restates utils/net_utils.py:93-114 and
    nets/OpticalUnet.py:123-146), forward + backward, on the host cores."""
    import torch
    from oracle import torch_ref
    leaves = [t.detach().clone().requires_grad_() for t in inp["f0"] + inp["f1"] + [inp["ff"], inp["fb"], inp["mf"], inp["mb"]]]
    f0, f1, (ff, fb, mf, mb) = leaves[:2], leaves[2:4], leaves[4:]
    outs = torch_ref.ref_warp_blend(f0, f1, ff, fb, mf, mb, padding_mode=pad, align_corners=False)
    if not fwd_only:
        torch.autograd.backward(outs, inp["gos"])
    return outs


def time_cpu_reference(cfg, n_sample, min_seconds, min_reps, max_reps):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sub = dict(cfg, N=n_sample)
    inp = make_inputs(sub, "cpu")
    fo = bool(cfg.get("fwd_only", False))
    cpu_reference_step(inp, fwd_only=fo)  # warm-up
    times = []
    t_all = time.perf_counter()
    while len(times) < max_reps and (len(times) < min_reps or time.perf_counter() - t_all < min_seconds):
        t = time.perf_counter()
        cpu_reference_step(inp, fwd_only=fo)
        times.append(time.perf_counter() - t)
    pix = n_sample * cfg["H"] * cfg["W"]
    return pix / statistics.median(times) / 1e9, cores, len(times)


def run_reference_arm(args, cfg, rank, world):
    """--impl reference: the reference path on the host cores (torch CPU, all threads), bounded sample per step."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fo = bool(cfg.get("fwd_only", False))
    probe = make_inputs(dict(cfg, N=1), "cpu")
    cpu_reference_step(probe, fwd_only=fo)
    t = time.perf_counter()
    cpu_reference_step(probe, fwd_only=fo)
    t1 = time.perf_counter() - t
    n_s = int(max(1, min(cfg["N"], 90.0 / max((args.steps + args.warmup) * t1, 1e-9))))
    inp = make_inputs(dict(cfg, N=n_s), "cpu")
    for _ in range(args.warmup):
        cpu_reference_step(inp, fwd_only=fo)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(inp, fwd_only=fo)
    el = time.perf_counter() - t
    pix = n_s * cfg["H"] * cfg["W"] * cfg["chain"]
    val = pix * args.steps / el / 1e9
    sample = f"{n_s} of {cfg['N']} clips per step ({cfg['H']}x{cfg['W']}, C={C_TOTAL}), torch {torch.__version__} CPU, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "warp+blend fwd+bwd Gpix/s", "value": val, "unit": "Gpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, args, args.gpus),  # the same keys and values as the b200 arm
        "cpu_baseline": {"value": val, "unit": "Gpix/s", "cores": cores, "kind": "port", "sample": sample,
                         "what": "oracle/torch_ref.py: the reference's own torch-op sequence (utils/net_utils.py:93-114, "
                                 "nets/OpticalUnet.py:123-146) restated; the reference has no installable package (no setup.py)"},
        "e2e": {"value": val, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def stock_torch_gpu_step(inp, return_leaves=False):
    """The composition the reference runs on its GPU today, restated inline (nothing imported from oracle/): base grid
    built on the CPU and sent to the device on every call (utils/net_utils.py:96-107), `grid = base -/+ flow` (:109-111),
    one F.grid_sample per modality and direction (:113, nets/VAE_S.py:134-135), mask weighting and sum
    (nets/OpticalUnet.py:141-146), autograd backward.  The GPU baseline SURVEY 8d asks to time next to ours."""
    import torch
    import torch.nn.functional as F
    leaves = [t.detach().clone().requires_grad_() for t in inp["f0"] + inp["f1"] + [inp["ff"], inp["fb"], inp["mf"], inp["mb"]]]
    f0, f1, (ff, fb, mf, mb) = leaves[:2], leaves[2:4], leaves[4:]
    N, _, H, W = ff.shape
    outs = []
    for a, b in zip(f0, f1):
        warped = []
        for x, fl, sign in ((a, ff, -1.0), (b, fb, 1.0)):
            base = torch.zeros(N, H, W, 2)
            base[..., 0] = torch.ger(torch.ones(H), torch.linspace(-1, 1, W))
            base[..., 1] = torch.ger(torch.linspace(-1, 1, H), torch.ones(W))
            grid = base.to(x.device) + sign * fl.transpose(1, 2).transpose(2, 3)
            warped.append(F.grid_sample(x, grid, padding_mode="border", align_corners=False))
        outs.append(mf * warped[0] + mb * warped[1])
    torch.autograd.backward(outs, inp["gos"])
    return (outs, leaves) if return_leaves else outs


def parity_check(step, inp):
    """--check: the buffers the timed loop just wrote (outputs and all gradients of the last step) against the stock torch
    CUDA composition on the same inputs; max|a-b| / max|ref| per quantity."""
    import torch
    torch.cuda.synchronize()
    ref_outs, leaves = stock_torch_gpu_step(inp, return_leaves=True)
    G = len(inp["f0"])
    rel = lambda a, r: float((a.reshape(r.shape).double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))
    res = {"fwd": max(rel(step.outs[g], ref_outs[g]) for g in range(G)),
           "gsrc": max(rel(step.g_srcs[g][d], leaves[d * G + g].grad) for g in range(G) for d in range(2)),
           "gflow": max(rel(step.g_flows[d], leaves[2 * G + d].grad) for d in range(2)),
           "gmask": max(rel(step.g_blends[d], leaves[2 * G + 2 + d].grad) for d in range(2)),
           "against": "stock torch CUDA composition (F.grid_sample + autograd) on the timed inputs",
           "bars": {"fwd": 1e-6, "bwd": 1e-5}}
    res["ok"] = bool(res["fwd"] <= 1e-6 and max(res["gsrc"], res["gflow"], res["gmask"]) <= 1e-5)
    return res


# --------------------------------------------------------------------------------------------- B200 arm
class CabiStep:
    """The three C-ABI entry points on preallocated device buffers (what the autograd op calls)."""

    def __init__(self, inp, deterministic, pad="border", atomic_src=False, zero="fwd"):
        import torch
        from deep_video_interpolation_extrapolation_b200 import _lib as L
        from deep_video_interpolation_extrapolation_b200._problem import fill_grads, fill_problem
        self.lib = L.load()
        self.L = L
        N, _, H, W = inp["ff"].shape
        dev = inp["ff"].device
        u5 = lambda t: t.unsqueeze(1)  # [N,C,H,W] -> [N,1,C,H,W]
        self.keep = []
        flows = [inp["ff"].unsqueeze(2), inp["fb"].unsqueeze(2)]
        blends = [inp["mf"], inp["mb"]]
        srcs = [[u5(a), u5(b)] for a, b in zip(inp["f0"], inp["f1"])]
        self.outs = [torch.empty(N, 1, c, H, W, device=dev) for c in CH]
        gos = [u5(g) for g in inp["gos"]]
        self.g_srcs = [[torch.empty(N, 1, c, H, W, device=dev) for _ in range(2)] for c in CH]
        self.g_flows = [torch.empty(N, 2, 1, H, W, device=dev) for _ in range(2)]
        self.g_blends = [torch.empty(N, 1, H, W, device=dev) for _ in range(2)]
        ptr, st = (lambda t: t.data_ptr()), (lambda t: t.stride())
        self.p = fill_problem(N=N, T=1, H=H, W=W, flows=flows, gates=[None, None], blends=blends, signs=[-1.0, 1.0], srcs=srcs,
                              outs=self.outs, padding_mode=L.FWB_PAD_BORDER if pad == "border" else L.FWB_PAD_ZEROS,
                              align_corners=False,
                              flags=(L.FWB_FLAG_DETERMINISTIC if deterministic else
                                     (L.FWB_FLAG_ATOMIC_SRC if atomic_src else L.FWB_FLAG_FUSED_BWD)), ptr=ptr, strides=st)
        self.fused = not deterministic and not atomic_src
        # who zeroes grad_src before the fused backward accumulates into it:
        #   "fwd"    the forward kernel's tiles (fwb_warp_blend_forward_zero), backward told FWB_FLAG_GRAD_SRC_ZEROED (default:
        #            what the autograd op does); "memset": the backward entry point's own memsets; "side": torch fills on a
        #            side stream while the forward runs (A/B only)
        self.zero = zero if self.fused else "memset"
        if self.zero in ("fwd", "side"):
            self.p.flags |= L.FWB_FLAG_GRAD_SRC_ZEROED
        if self.zero == "side":
            self.side = torch.cuda.Stream(dev)
            self.main = torch.cuda.current_stream(dev)
        self.q = fill_grads(self.p, grad_outs=gos, grad_srcs=self.g_srcs, grad_flows=self.g_flows, grad_gates=[None, None],
                            grad_blends=self.g_blends, ptr=ptr, strides=st)
        self.keep += [flows, blends, srcs, gos]
        self.ws_bytes = int(self.lib.fwb_workspace_bytes(ctypes.byref(self.p)))
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.current_stream(dev).cuda_stream

    def forward(self):
        if self.zero == "fwd":
            self.L.check(self.lib.fwb_warp_blend_forward_zero(ctypes.byref(self.p), ctypes.byref(self.q), self.stream), "forward_zero")
        else:
            self.L.check(self.lib.fwb_warp_blend_forward(ctypes.byref(self.p), self.stream), "forward")

    def backward_flow(self):
        self.L.check(self.lib.fwb_warp_blend_backward_flow(ctypes.byref(self.p), ctypes.byref(self.q), self.ws.data_ptr(),
                                                           self.ws_bytes, self.stream), "backward_flow")

    def backward_src(self):
        self.L.check(self.lib.fwb_warp_blend_backward_src(ctypes.byref(self.p), ctypes.byref(self.q), self.ws.data_ptr(),
                                                          self.ws_bytes, self.stream), "backward_src")

    def step(self):
        if self.zero == "side":
            import torch
            self.side.wait_stream(self.main)  # the previous step's backward has consumed grad_src
            with torch.cuda.stream(self.side):
                for row in self.g_srcs:
                    for g in row:
                        g.zero_()
            self.forward()
            self.main.wait_stream(self.side)
        else:
            self.forward()
        self.backward_flow()
        self.backward_src()


class AutogradStep:
    """The headline op through the PUBLIC autograd API (warp_blend + torch.autograd.backward): what a trainer calls.  Includes
    the op's own host cost (struct fill, ctypes, output / gradient allocation by torch's caching allocator)."""

    def __init__(self, inp):
        import deep_video_interpolation_extrapolation_b200 as P
        self.P, self.gos = P, inp["gos"]
        G = len(inp["f0"])
        self.leaves = [t.detach().clone().requires_grad_() for t in inp["f0"] + inp["f1"] + [inp["ff"], inp["fb"], inp["mf"], inp["mb"]]]
        self.G = G

    def step(self):
        import torch
        G, lv = self.G, self.leaves
        outs = self.P.warp_blend(lv[:G], lv[G:2 * G], lv[2 * G], lv[2 * G + 1], lv[2 * G + 2], lv[2 * G + 3])
        torch.autograd.backward(outs, self.gos)
        for t in lv:
            t.grad = None


class ChainStep:
    """config 3: K chained invocations per training step (runners/ExtraTrainer.py:254-310): step k+1 warps [the newer source
    frame of step k, prediction k]; the predicted seg is re-one-hotted by argmax (:308-310, no gradient through it), the RGB
    prediction carries the gradient; ONE backward over the whole chain (every step has its own loss terms, :284-288).
    Through the public autograd op.  Flows / masks of every step are network outputs (requires_grad); the frames are data."""

    def __init__(self, inp, K, cfg, dev, seed):
        import deep_video_interpolation_extrapolation_b200 as P
        self.P, self.K, self.gos = P, K, inp["gos"]
        self.f0 = [t.detach() for t in inp["f0"]]
        self.f1 = [t.detach() for t in inp["f1"]]
        self.nets = []  # per step: ff, fb, mf, mb
        for k in range(K):
            src = inp if k == 0 else make_inputs(cfg, dev, seed=seed + 1000 * k)
            self.nets.append([src[n].detach().clone().requires_grad_() for n in ("ff", "fb", "mf", "mb")])
            if k:
                del src

    def step(self):
        import torch
        A, B = self.f0, self.f1
        outs_all = []
        for k in range(self.K):
            ff, fb, mf, mb = self.nets[k]
            o = self.P.warp_blend(A, B, ff, fb, mf, mb)
            outs_all += o
            if k + 1 < self.K:
                lab = o[1].argmax(dim=1, keepdim=True)
                A, B = B, [o[0], torch.zeros_like(o[1]).scatter_(1, lab, 1.0)]
        torch.autograd.backward(outs_all, self.gos * self.K)
        for net in self.nets:
            for t in net:
                t.grad = None


def timed(fn, steps, warmup, sync):
    import torch
    for _ in range(warmup):
        fn()
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    sync()
    return a.elapsed_time(b) / 1e3  # seconds


def run_b200_arm(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import deep_video_interpolation_extrapolation_b200 as P
    from deep_video_interpolation_extrapolation_b200 import _lib
    _lib.load()  # no fallback: fail loudly when the CUDA library is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from deep_video_interpolation_extrapolation_b200 import sharding
    numa_cores = sharding.bind_to_gpu_numa(local_rank) if (world > 1 and not args.no_numa_bind) else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    inp = make_inputs(cfg, dev, seed=rank)
    ar_buf = torch.zeros(cfg["allreduce"], device=dev) if cfg["allreduce"] else None
    chain = cfg["chain"]
    fwd_only = bool(cfg.get("fwd_only", False))
    step = CabiStep(inp, args.deterministic, atomic_src=args.atomic_src, zero="memset" if fwd_only else args.zero)
    chain_step = ChainStep(inp, chain, cfg, dev, seed=rank) if chain > 1 else None

    def one_step():
        work = None
        if ar_buf is not None and world > 1:
            # config 5: the parameter-gradient all-reduce DDP issues while the rest of the backward still runs: async on
            # NCCL's stream, joined at the end of the step
            work = dist.all_reduce(ar_buf, async_op=True)
        if chain_step is not None:
            chain_step.step()  # config 3: a true K-step chain through the autograd op
        elif fwd_only:
            step.forward()     # config 1: the forward warp only
        else:
            step.step()
        if work is not None:
            work.wait()

    pix_step = cfg["N"] * cfg["H"] * cfg["W"] * chain
    if sampler:
        sampler.mark("load_start")
    # ---- headline: device-resident inputs, direct C-ABI launches, CUDA events, max over ranks
    for _ in range(args.warmup):
        one_step()
    sync()
    if sampler:
        sampler.mark("timed_start")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        one_step()
    b.record()
    sync()
    if sampler:
        sampler.mark("timed_end")
    from deep_video_interpolation_extrapolation_b200 import sharding
    el = sharding.max_over_ranks(a.elapsed_time(b) / 1e3, dev)  # the job's step time is the slowest rank's device time
    value = sharding.job_throughput(pix_step, args.steps, el, world) / 1e9

    # ---- --check (default on): the buffers the timed loop just wrote against the stock torch CUDA composition
    pcheck = None
    if not args.no_check and not args.profile and chain == 1 and not fwd_only:
        pcheck = parity_check(step, inp)
    if args.profile:
        if sampler:
            sampler.stop()
        print(f"profile run: {value:.3f} Gpix/s, {el / args.steps * 1e3:.3f} ms/step")
        return
    # ---- per-kernel durations (same buffers, CUDA events on the launching stream)
    kt = {}
    ksteps = max(10, min(args.steps, 100))
    if fwd_only:
        klist = (("forward", step.forward),)
    elif step.fused:  # kernels 2+3 run as one fused launch inside the backward_flow entry point
        klist = (("forward", step.forward), ("backward_fused", step.backward_flow))
    else:
        klist = (("forward", step.forward), ("backward_flow", step.backward_flow), ("backward_src", step.backward_src))
    for name, fn in klist:
        kt[name] = timed(fn, ksteps, 3, sync) / ksteps

    # ---- aux: `refine`'s mask blend (utils/net_utils.py:141-143) as one streaming kernel each way, T = 3 frames
    aux = None
    if args.aux:
        Ta, Ca = 3, sum(CH)
        a_in = torch.randn(cfg["N"], Ta, Ca, cfg["H"], cfg["W"], device=dev, requires_grad=True)
        a_m = torch.rand(cfg["N"], Ta, cfg["H"], cfg["W"], device=dev, requires_grad=True)
        a_nz = torch.randn(cfg["N"], 3, cfg["H"], cfg["W"], device=dev, requires_grad=True)
        a_go = torch.randn_like(a_in)
        a_out = P.mask_blend(a_in, a_m, a_nz)
        pixf = cfg["N"] * Ta * cfg["H"] * cfg["W"]
        t_f = timed(lambda: P.mask_blend(a_in, a_m, a_nz), 20, 3, sync) / 20
        t_b = timed(lambda: torch.autograd.grad(a_out, (a_in, a_m, a_nz), a_go, retain_graph=True), 20, 3, sync) / 20
        bf = pixf * (8 * Ca + 4) + cfg["N"] * cfg["H"] * cfg["W"] * 12          # in + out + mask (+ noise once per clip)
        bb = pixf * (12 * Ca + 8) + cfg["N"] * cfg["H"] * cfg["W"] * 24        # go + in + gi, mask + gm (+ noise, gnoise)
        pk = peaks()[0]
        aux = {"mask_blend_forward": {"ms": t_f * 1e3, "GBps": bf / t_f / 1e9, "frac": bf / t_f / 1e9 / pk, "bytes": bf},
               "mask_blend_backward": {"ms": t_b * 1e3, "GBps": bb / t_b / 1e9, "frac": bb / t_b / 1e9 / pk, "bytes": bb,
                                       "note": "autograd.grad through the op (includes torch's allocation of the 3 gradients)"},
               "shape": [cfg["N"], Ta, Ca, cfg["H"], cfg["W"]]}
        del a_in, a_m, a_nz, a_go, a_out

    # ---- variant (reported separately, different algorithmic bytes): segmentation given as uint8 label maps
    variants = None
    if args.aux:
        K = CH[1]
        lab0 = inp["f0"][1].argmax(1).to(torch.uint8)
        lab1 = inp["f1"][1].argmax(1).to(torch.uint8)
        lv = [t.detach().clone().requires_grad_() for t in (inp["f0"][0], inp["f1"][0], inp["ff"], inp["fb"], inp["mf"], inp["mb"])]

        def label_step():
            outs = P.warp_blend_labels([lv[0]], [lv[1]], lab0, lab1, K, lv[2], lv[3], lv[4], lv[5])
            torch.autograd.backward(outs, inp["gos"])
            for t in lv:
                t.grad = None

        # flow-only backward: the sources are data (requires_grad=False), the usual training case; no grad_src traffic
        fo = [t.detach() for t in inp["f0"] + inp["f1"]] + [t.detach().clone().requires_grad_() for t in (inp["ff"], inp["fb"], inp["mf"], inp["mb"])]

        def flow_only_step():
            outs = P.warp_blend(fo[0:2], fo[2:4], fo[4], fo[5], fo[6], fo[7])
            torch.autograd.backward(outs, inp["gos"])
            for t in fo[4:]:
                t.grad = None

        t_fo = timed(flow_only_step, 20, 3, sync) / 20
        t_l = timed(label_step, 20, 3, sync) / 20
        pix = cfg["N"] * cfg["H"] * cfg["W"]
        c3 = CH[0]
        # fwd: R 2 rgb frames 8*c3, 2 label maps 2, flows 16, masks 8, W out 4*(c3+K); bwd: R gout 4*(c3+K), rgb 8*c3, labels 2,
        # flows 16, masks 8, W grad rgb 8*c3, gflows 16, gmasks 8
        lb = (8 * c3 + 2 + 24 + 4 * (c3 + K)) + (4 * (c3 + K) + 8 * c3 + 2 + 24 + 8 * c3 + 24)
        fob = (12 * C_TOTAL + 24) + (12 * C_TOTAL + 48)  # fwd + (R gout 4C, R x0,x1 8C, flows 16, masks 8, W gflows 16, gmasks 8)
        variants = {"no_source_gradient": {"ms_per_step": t_fo * 1e3, "Gpix_per_s": cfg["N"] * cfg["H"] * cfg["W"] / t_fo / 1e9,
                                           "bytes_per_pixel": fob,
                                           "frac": cfg["N"] * cfg["H"] * cfg["W"] * fob / t_fo / 1e9 / peaks()[0],
                                           "api": "warp_blend + autograd.backward with requires_grad=False frames",
                                           "note": "SURVEY 8a: reduced byte count when grad w.r.t. the sources is not needed; "
                                                   "NOT the headline metric"},
                    "compact_seg_labels": {"ms_per_step": t_l * 1e3, "Gpix_per_s": pix / t_l / 1e9, "bytes_per_pixel": lb,
                                           "GBps": pix * lb / t_l / 1e9, "frac": pix * lb / t_l / 1e9 / peaks()[0],
                                           "api": "warp_blend_labels + autograd.backward (RGB dense kernels + label kernels; "
                                                  "autograd sums the two flow / mask gradients)",
                                           "note": "seg as uint8 labels (SURVEY 8f row 4): same outputs as the dense op on one-hot "
                                                   "maps, no seg source gradient; NOT the headline metric"}}

    # ---- aux: the stock torch composition on this GPU, and other flow regimes, reported separately
    regimes = None
    if args.aux:
        pixs = cfg["N"] * cfg["H"] * cfg["W"]
        t_s = timed(lambda: stock_torch_gpu_step(inp), 5, 2, sync) / 5
        regimes = {"gpu_stock_torch": {"ms_per_step": t_s * 1e3, "Gpix_per_s": pixs / t_s / 1e9,
                                       "what": "reference composition on this GPU: CPU base grid + H2D per call, 4 x F.grid_sample, "
                                               "mask weighting, autograd (torch " + torch.__version__ + ")"}}
        for name, mk in (("sigma2_small", lambda f, k: f * (2.0 / cfg["sigma"])),
                         ("adversarial_uniform_pm0.5", lambda f, k: (torch.rand(f.shape, generator=torch.Generator().manual_seed(k)) - 0.5).to(dev))):
            alt = dict(inp, ff=mk(inp["ff"], 1).contiguous(), fb=mk(inp["fb"], 2).contiguous())
            st2 = CabiStep(alt, args.deterministic, atomic_src=args.atomic_src, zero=args.zero)
            t_r = timed(st2.step, 10, 3, sync) / 10
            regimes[name] = {"ms_per_step": t_r * 1e3, "Gpix_per_s": pixs / t_r / 1e9, "frac_of_hbm_roofline": pixs * BYTES_PER_PIX / t_r / 1e9 / peaks()[0]}
            del st2, alt

    # ---- the headline op through the public autograd API (same tensors, same kernels + the op's host-side cost)
    via_autograd = None
    if not fwd_only and chain == 1:
        ag = AutogradStep(inp)
        t_ag = sharding.max_over_ranks(timed(ag.step, max(10, min(args.steps, 50)), 3, sync), dev) / max(10, min(args.steps, 50))
        via_autograd = {"ms_per_step": t_ag * 1e3, "Gpix_per_s": world * cfg["N"] * cfg["H"] * cfg["W"] / t_ag / 1e9,
                        "api": "deep_video_interpolation_extrapolation_b200.warp_blend + torch.autograd.backward (allocations included)"}
        del ag

    # ---- e2e: public API for HOST buffers (HostWarpBlend: the autograd op per batch chunk, pinned host in -> pinned host
    # out, H2D of every input and D2H of every output / gradient inside the timed region, copies overlapped with compute)
    host_in = [t.cpu().pin_memory() for t in inp["f0"] + inp["f1"] + [inp["ff"], inp["fb"], inp["mf"], inp["mb"]] + inp["gos"]]
    pipe = P.HostWarpBlend(dev, chunk=args.e2e_chunk, padding_mode="border", deterministic=args.deterministic)
    G = len(inp["f0"])

    # the producer (data loader / previous stage) writes its tensors straight into the pinned per-chunk ARENAS, so every chunk
    # moves with ONE cudaMemcpyAsync each way (HostWarpBlend.arena / run_arena); --e2e-per-tensor times the 13 + 11 copies path
    if args.e2e_per_tensor:
        def e2e_step():
            pipe.run(host_in[:G], host_in[G:2 * G], host_in[2 * G], host_in[2 * G + 1], host_in[2 * G + 2], host_in[2 * G + 3],
                     host_in[2 * G + 4:], synchronize=False)
    else:
        pipe.fill_arena(host_in[:G], host_in[G:2 * G], host_in[2 * G], host_in[2 * G + 1], host_in[2 * G + 2], host_in[2 * G + 3],
                        host_in[2 * G + 4:])

        def e2e_step():
            pipe.run_arena(synchronize=False)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_el = sharding.max_over_ranks(timed(e2e_step, e2e_steps, 2, sync), dev)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_val = sharding.job_throughput(cfg["N"] * cfg["H"] * cfg["W"], e2e_steps, e2e_el, world) / 1e9
    if sampler:
        sampler.mark("load_end")
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = peaks()
        pix_launch = cfg["N"] * cfg["H"] * cfg["W"]
        kernels = {k: {"ms": v * 1e3, "bytes_per_pixel": KERNEL_BYTES[k], "GBps": pix_launch * KERNEL_BYTES[k] / v / 1e9,
                       "frac": pix_launch * KERNEL_BYTES[k] / v / 1e9 / peak} for k, v in kt.items()}
        dom = max(kt, key=kt.get)
        # algorithmic bytes per pixel of the timed step: the headline op (808), the forward alone (config 1: 300), or the chain of
        # config 3 (frames are data: no source gradient at step 0, only the RGB prediction's gradient (12 B) afterwards)
        C = C_TOTAL
        if fwd_only:
            step_bytes = 12 * C + 24
        elif chain > 1:
            step_bytes = (12 * C + 24) + (12 * C + 48) + 12.0 * (chain - 1) / chain
        else:
            step_bytes = BYTES_PER_PIX
        step_gbs = value / world * step_bytes  # per-GPU Gpix/s * B/pix = GB/s
        cfgd = config_dict(cfg, args, world)
        launches = (1 if fwd_only else (LAUNCHES_PER_STEP["fused"] if step.fused else LAUNCHES_PER_STEP["split"]))
        if chain > 1:  # per chain: step 0 (frames are data) forward + kernel 2; later steps forward (zero-fill) + fused backward
            launches = 2 * chain
        out = {
            "metric": "warp+blend fwd+bwd Gpix/s", "value": value, "unit": "Gpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac"], "traffic": ncu_traffic(dom, args.config), "peak_source": peak_src,
                         "algorithmic_bytes_per_pixel": KERNEL_BYTES[dom]},
            "roofline_step": {"bytes_per_pixel": step_bytes, "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                              "frac": step_gbs / peak, "frac_of_nominal_8TBps": step_gbs / 8000.0},
            "kernels": kernels,
            "e2e": {"value": e2e_val, "unit": "Gpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "chunk": args.e2e_chunk, "numa_local_cores": numa_cores,
                    "h2d_GBps": h2d * e2e_steps / e2e_el / 1e9, "d2h_GBps": d2h * e2e_steps / e2e_el / 1e9,
                    "copies_per_chunk": "13 + 11 (per tensor)" if args.e2e_per_tensor else "1 + 1 (pinned arenas)",
                    "api": ("deep_video_interpolation_extrapolation_b200.HostWarpBlend." + ("run" if args.e2e_per_tensor else "run_arena")
                            + " (forward + backward C-ABI calls per batch chunk; pinned host in/out, H2D | compute | D2H on three streams)")},
            "gpu_launches": args.steps * launches,
            "clocks": clocks,
        }
        if via_autograd is not None:
            out["via_autograd"] = via_autograd
        if pcheck is not None:
            out["parity_check"] = pcheck
        if aux:
            out["aux_kernels"] = aux
        if variants:
            out["variants"] = variants
        if regimes:
            out["other_measurements"] = regimes
        if world == 1 and not args.no_cpu:
            v, cores, reps = time_cpu_reference(cfg, cfg["N"] if cfg["H"] * cfg["W"] <= 256 * 512 else 1, 10.0, 3, 30)
            ns = cfg["N"] if cfg["H"] * cfg["W"] <= 256 * 512 else 1
            out["cpu_baseline"] = {"value": v, "unit": "Gpix/s", "cores": cores, "kind": "port",
                                   "sample": f"{reps} reps of {ns} of {cfg['N']} clips, torch {torch.__version__} CPU restatement "
                                             f"of the reference composition (oracle/torch_ref.py), fwd+bwd"}
        line = json.dumps(out)
        print(line)
        if args.record:  # tracked record of non-default configs (the driver only runs the default one)
            with open(os.path.join(ROOT, "profiles", "r2_bench_lines.jsonl"), "a") as f:
                f.write(json.dumps({"label": args.record, **out}) + "\n")
    if world > 1:
        dist.destroy_process_group()


# launches of OUR kernels per step (see csrc/flowwarp_b200.cu)
# fused: fwd_tex_kernel 1 (also zero-fills grad_src) + bwd_tile_kernel<TEX> 1 (--zero memset: 4 cudaMemsetAsync nodes, not counted);
# split (deterministic): forward 1 + (table init 1 + emit 1 + kernel 2) + kernel 3 x2
LAUNCHES_PER_STEP = {"fused": 1 + 1, "split": 1 + 3 + 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--atomic-src", action="store_true", help="A/B: grad_src via the global-atomic scatter kernel")
    ap.add_argument("--sigma", type=float, default=None, help="override flow sigma in pixels")
    ap.add_argument("--zero", default="fwd", choices=["fwd", "memset", "side"],
                    help="who zero-fills grad_src before the fused backward: the forward kernel (default), the backward's "
                         "memsets, or a side stream (A/B)")
    ap.add_argument("--e2e-chunk", type=int, default=1, help="clips per chunk of the host pipeline (e2e leg)")
    ap.add_argument("--e2e-per-tensor", action="store_true", help="e2e leg with one copy per tensor instead of the pinned arenas")
    ap.add_argument("--aux", action="store_true", help="also time the mask-blend (refine) kernels")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not pin each rank to its GPU's NUMA-local cores")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--check", action="store_true", help="(default) compare the timed buffers with the stock torch CUDA composition")
    ap.add_argument("--no-check", action="store_true", help="skip the parity_check of the timed buffers")
    ap.add_argument("--record", default=None, help="append the JSON line to profiles/r2_bench_lines.jsonl under this label")
    ap.add_argument("--profile", action="store_true", help="only warm-up + timed steps (for ncu); prints no JSON")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = dict(CONFIGS[args.config])
    if args.sigma is not None:
        cfg["sigma"] = args.sigma
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
    else:
        run_b200_arm(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
