"""Run the compact-seg variant (warp_blend_labels fwd+bwd, config 2 shapes) a few times: for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import deep_video_interpolation_extrapolation_b200 as P
cfg = dict(bench.CONFIGS[2])
dev = torch.device("cuda:0")
inp = bench.make_inputs(cfg, dev)
K = 20
lab0 = inp["f0"][1].argmax(1).to(torch.uint8)
lab1 = inp["f1"][1].argmax(1).to(torch.uint8)
lv = [t.detach().clone().requires_grad_() for t in (inp["f0"][0], inp["f1"][0], inp["ff"], inp["fb"], inp["mf"], inp["mb"])]
for _ in range(4):
    outs = P.warp_blend_labels([lv[0]], [lv[1]], lab0, lab1, K, lv[2], lv[3], lv[4], lv[5])
    torch.autograd.backward(outs, inp["gos"])
    for t in lv:
        t.grad = None
torch.cuda.synchronize()
print("ok")
