python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_v1.log; tail -3 gpurun_out/pytest_v1.log
for ch in 1 2 4; do
python bench.py --steps 50 --warmup 5 --no-cpu --e2e-chunk $ch > gpurun_out/bench_v1_$ch.log 2> gpurun_out/bench_v1_$ch.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v1_$ch.log").read().strip().splitlines()[-1])
print("chunk $ch", d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["value"])
EOP
done
tail -3 gpurun_out/bench_v1_2.err
