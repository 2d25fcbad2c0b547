python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_v5.log; tail -3 gpurun_out/pytest_v5.log
python bench.py --steps 50 --warmup 5 --no-cpu --aux > gpurun_out/bench_v5.log 2> gpurun_out/bench_v5.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v5.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["value"])
print(d.get("aux_kernels"))
EOP
tail -3 gpurun_out/bench_v5.err
