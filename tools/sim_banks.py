"""Simulate shared-memory wavefronts of the gather (LDS) and scatter (ATOMS) of the tile kernels on the benchmark's flows.
numpy only (no GPU).  Usage: python tools/sim_banks.py"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

def coords(cfg, n_use=2, seed=0):
    cfg = dict(cfg); cfg["N"] = n_use
    inp = bench.make_inputs(cfg, "cpu", seed)
    H, W = cfg["H"], cfg["W"]
    out = []
    for f, sign in ((inp["ff"], -1.0), (inp["fb"], 1.0)):
        f = f.numpy().astype(np.float64)
        bx = np.linspace(-1, 1, W)[None, None, :]
        by = np.linspace(-1, 1, H)[None, :, None]
        gx = bx + sign * f[:, 0]; gy = by + sign * f[:, 1]
        ix = ((gx + 1) * W - 1) / 2; iy = ((gy + 1) * H - 1) / 2
        ix = np.clip(ix, 0, W - 1); iy = np.clip(iy, 0, H - 1)
        out.append((np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)))
    return out, H, W

def wavefronts(addr, atomic):
    """addr: [ninstr, 32] word addresses (or -1 inactive). returns wavefronts per instruction."""
    n = addr.shape[0]
    res = np.zeros(n, np.int64)
    for i in range(n):
        a = addr[i]; a = a[a >= 0]
        if a.size == 0: continue
        if not atomic: a = np.unique(a)
        res[i] = np.bincount(a % 32, minlength=32).max()
    return res

def patch_lanes(H, W, pw, ph):
    """yield index arrays (ys, xs) [npatch, 32] for pw x ph patches (lane = px + pw*py)"""
    ly, lx = np.divmod(np.arange(32), pw)
    ys = (np.arange(0, H, ph)[:, None, None] + ly[None, None, :])
    xs = (np.arange(0, W, pw)[None, :, None] + lx[None, None, :])
    ys, xs = np.broadcast_arrays(ys, xs)
    return ys.reshape(-1, 32), xs.reshape(-1, 32)

def main():
    cfg = bench.CONFIGS[int(os.environ.get("CFG", "2"))]
    (dirs, H, W) = coords(cfg)
    rng = np.random.default_rng(0)
    for pw, ph in ((8, 4), (16, 2), (32, 1), (4, 8)):
        ys, xs = patch_lanes(H, W, pw, ph)
        sel = rng.choice(ys.shape[0], 3000, replace=False)
        ys_s, xs_s = ys[sel], xs[sel]
        for skew in (8, 12, 20, 33, 36, 40, 44, 48, 52):
            tot = {}
            for (x0, y0) in dirs[:1]:
                X = x0[0][ys_s, xs_s]; Y = y0[0][ys_s, xs_s]
                lane = np.arange(32)[None, :]
                lpx, lpy = lane % pw, lane // pw
                base = X + skew * Y
                taps = {"nw": base, "ne": base + 1, "sw": base + skew, "se": base + skew + 1}
                ld = sum(wavefronts(t, False).mean() for t in taps.values())
                at = sum(wavefronts(t, True).mean() for t in taps.values())
                # rotated: instruction k: lane does tap (ex ^ kx, ey ^ ky) with ex = px parity, ey = py parity
                ex, ey = lpx & 1, lpy & 1
                atr = 0.0
                for kx in (0, 1):
                    for ky in (0, 1):
                        a = base + (ex ^ kx) + skew * (ey ^ ky)
                        atr += wavefronts(a, True).mean()
                # rotation on x only
                atx = 0.0
                for kx in (0, 1):
                    for ky in (0, 1):
                        a = base + (ex ^ kx) + skew * ky
                        atx += wavefronts(a, True).mean()
                tot = (ld, at, atr, atx)
            print(f"patch {pw}x{ph} skew {skew:2d}: LDS {tot[0]:.2f}  ATOMS {tot[1]:.2f}  ATOMS rot-xy {tot[2]:.2f}  rot-x {tot[3]:.2f}   (sum over 4 taps)")
main()
