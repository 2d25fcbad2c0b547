#!/bin/bash
# quick check under gpurun: parity tests, then the config-2 bench line with per-kernel times.  usage: tools/run_r2b.sh <tag> [pytest -k expr]
T=${1:-x}
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" 2>&1 | tail -15 > gpurun_out/pytest_$T.log
else
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$T.log
fi
tail -3 gpurun_out/pytest_$T.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_$T.log 2> gpurun_out/bench_$T.err
python tools/kt.py gpurun_out/bench_$T.log
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_$T.log").read().strip().splitlines()[-1])
print("parity", d.get("parity_check"), "e2e", d.get("e2e",{}).get("value"))
EOP
