"""Names and device times of the kernels one forward+backward step launches (torch profiler / CUPTI; no ncu needed).
usage: python tools/kernel_list.py [config]   (env FWB_KERNELS selects variants)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

cfg = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
inp = bench.make_inputs(cfg, torch.device("cuda"), 0)
st = bench.CabiStep(inp, False)
if len(sys.argv) > 2 and sys.argv[2] == "flowonly":  # sources without gradient: kernel 2 only
    import deep_video_interpolation_extrapolation_b200 as P
    fo = [t.detach() for t in inp["f0"] + inp["f1"]] + [t.detach().clone().requires_grad_() for t in (inp["ff"], inp["fb"], inp["mf"], inp["mb"])]

    class _S:
        def step(self):
            outs = P.warp_blend(fo[0:2], fo[2:4], fo[4], fo[5], fo[6], fo[7])
            torch.autograd.backward(outs, inp["gos"])
    st = _S()
for _ in range(3):
    st.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        st.step()
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"{e.device_time_total / e.count:10.1f} us x {e.count:3d}  {e.key[:150]}")
