python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final.log; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -3 gpurun_out/smoke_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.log 2> gpurun_out/bench_final_ref.err; cut -c1-400 gpurun_out/bench_final_ref.log
python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_final.log").read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ("value","ms_per_step","roofline","e2e","cpu_baseline","gpu_launches","clocks")}, indent=1)[:2500])
EOP
tail -3 gpurun_out/bench_final.err
