for c in 3 4; do
python bench.py --config $c --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg$c.log 2> gpurun_out/bench_cfg$c.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_cfg$c.log").read().strip().splitlines()[-1])
print("cfg $c", round(d["value"],3), "Gpix/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "frac", round(d["roofline_step"]["frac"],3))
EOP
done
FWB_KERNELS=notile python bench.py --config 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_cfg4_notile.log 2>&1
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_cfg4_notile.log").read().strip().splitlines()[-1])
print("cfg 4 notile", round(d["value"],3), "Gpix/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms"],4) for k,v in d["kernels"].items()})
EOP
