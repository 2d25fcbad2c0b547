python -m pytest tests -m gpu -x -q -k "host_pipeline" 2>&1 | tail -5 > gpurun_out/pytest_v2.log; tail -3 gpurun_out/pytest_v2.log
for v in "--e2e-chunk 1" "--e2e-chunk 2 --prezero"; do
tag=$(echo $v | tr -d ' -')
python bench.py --steps 50 --warmup 5 --no-cpu $v > gpurun_out/bench_v2_$tag.log 2> gpurun_out/bench_v2_$tag.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v2_$tag.log").read().strip().splitlines()[-1])
print("$v", d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["value"])
EOP
tail -3 gpurun_out/bench_v2_$tag.err
done
