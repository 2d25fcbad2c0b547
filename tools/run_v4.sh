python -m pytest tests -m gpu -x -q -k "forward_zero or backward_twice or warp_blend_vs_oracle or host_pipeline" 2>&1 | tail -4 > gpurun_out/pytest_v4.log; tail -2 gpurun_out/pytest_v4.log
for v in "--zero fwd" "--zero memset"; do
tag=$(echo $v | tr -d ' -')
python bench.py --steps 100 --warmup 5 --no-cpu $v > gpurun_out/bench_v4_$tag.log 2> gpurun_out/bench_v4_$tag.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v4_$tag.log").read().strip().splitlines()[-1])
print("$v", d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["value"])
EOP
tail -3 gpurun_out/bench_v4_$tag.err
done
