"""Phase timeline of bwd_cl_kernel on the config-2 inputs (needs the -DCL_PROF build: FWB_LIB=.../libflowwarp_b200_ab.so)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

cfg = dict(bench.CONFIGS[int(os.environ.get("CFG", "2"))])
inp = bench.make_inputs(cfg, torch.device("cuda:0"))
st = bench.CabiStep(inp, deterministic=False, zero=os.environ.get("ZERO", "fwd"))
buf = (ctypes.c_uint64 * 64)()
for _ in range(3):
    st.step()
st.lib.fwb_debug_cl_prof(buf)
reps = 5
for _ in range(reps):
    st.step()
st.lib.fwb_debug_cl_prof(buf)
tiles = cfg["N"] * ((cfg["H"] + 7) // 8) * ((cfg["W"] + 31) // 32) * reps
names = {0: "px taps", 1: "px tables+descriptors", 2: "px wait syncthreads", 3: "px kernel 2 loop", 4: "px epilogue",
         5: "px own SLOW scatter", 16: "ch staging issue+chan table", 17: "ch cp.async wait", 18: "ch wait syncthreads", 19: "ch amax",
         20: "ch d0 clear", 21: "ch d0 bar", 22: "ch d0 scatter", 23: "ch d0 bar", 24: "ch d0 flush setup", 25: "ch d0 flush", 26: "ch d0 bar",
         28: "ch d1 clear", 29: "ch d1 bar", 30: "ch d1 scatter", 31: "ch d1 bar", 32: "ch d1 flush setup", 33: "ch d1 flush", 34: "ch d1 bar",
         44: "ch slow items", 45: "ch initial fill"}
tot = {"px": 0.0, "ch": 0.0}
for k in sorted(names):
    v = buf[k] / tiles
    tot[names[k][:2]] += v
    print(f"{names[k]:32s} {v:9.0f} clk/tile")
print(tot)
