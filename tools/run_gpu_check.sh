python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_t13.log; tail -3 gpurun_out/pytest_t13.log
python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_t13.log 2> gpurun_out/bench_t13.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_t13.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
EOP
ncu --set full --import-source on --clock-control none -k regex:tile_kernel -s 2 -c 2 -o gpurun_out/prof_t13 python bench.py --profile --steps 1 --warmup 3 > gpurun_out/ncu_t13.log 2>&1; tail -2 gpurun_out/ncu_t13.log
