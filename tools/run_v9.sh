python -m pytest tests -m gpu -x -q -k "tile or warp_blend_vs or golden or full_size or large" 2>&1 | tail -6 > gpurun_out/pytest_v9.log; tail -3 gpurun_out/pytest_v9.log
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_v9.log 2> gpurun_out/bench_v9.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v9.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["value"])
EOP
tail -3 gpurun_out/bench_v9.err
