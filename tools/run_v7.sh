python -m pytest tests -m gpu -x -q -k "label" 2>&1 | tail -12 > gpurun_out/pytest_v7.log; tail -12 gpurun_out/pytest_v7.log
python bench.py --steps 30 --warmup 5 --no-cpu --aux > gpurun_out/bench_v7.log 2> gpurun_out/bench_v7.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v7.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d.get("variants"))
EOP
tail -3 gpurun_out/bench_v7.err
