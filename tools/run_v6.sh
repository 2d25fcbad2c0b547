python -m pytest tests -m gpu -x -q -k "mask_blend" 2>&1 | tail -3 > gpurun_out/pytest_v6.log; tail -2 gpurun_out/pytest_v6.log
python bench.py --steps 30 --warmup 5 --no-cpu --aux > gpurun_out/bench_v6.log 2> gpurun_out/bench_v6.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v6.log").read().strip().splitlines()[-1])
print({k:(round(v["ms"],4),round(v["frac"],3)) for k,v in d["aux_kernels"].items() if isinstance(v,dict)})
EOP
tail -3 gpurun_out/bench_v6.err
