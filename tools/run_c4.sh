T=$1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_$T.log; tail -2 gpurun_out/pytest_$T.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${T}_c2.log 2> gpurun_out/bench_${T}_c2.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${T}_c4.log 2> gpurun_out/bench_${T}_c4.err
python tools/kt.py gpurun_out/bench_${T}_c2.log gpurun_out/bench_${T}_c4.log
python - <<EOP
import json
for c in ("c2","c4"):
    d=json.loads(open("gpurun_out/bench_${T}_"+c+".log").read().strip().splitlines()[-1]); print(c, d.get("parity_check"))
EOP
