T=$1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --aux > gpurun_out/bench_$T.log 2> gpurun_out/bench_$T.err
python tools/kt.py gpurun_out/bench_$T.log
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_$T.log").read().strip().splitlines()[-1])
print("parity", d.get("parity_check"))
print("variants", {k:(round(v["ms_per_step"],4)) for k,v in (d.get("variants") or {}).items()})
print("other", {k:(round(v["ms_per_step"],4)) for k,v in (d.get("other_measurements") or {}).items()})
print("via_autograd", d.get("via_autograd",{}).get("ms_per_step"), "e2e", d.get("e2e",{}).get("value"))
EOP
timeout 300 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu --deterministic > gpurun_out/bench_${T}_det.log 2> gpurun_out/bench_${T}_det.err
python tools/kt.py gpurun_out/bench_${T}_det.log
