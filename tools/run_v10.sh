for ch in 1 2 4 8; do
python bench.py --steps 20 --warmup 3 --no-cpu --e2e-chunk $ch > gpurun_out/bench_v10_$ch.log 2> gpurun_out/bench_v10_$ch.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v10_$ch.log").read().strip().splitlines()[-1])
print("chunk $ch", round(d["ms_per_step"],4), {k:round(v["ms"],4) for k,v in d["kernels"].items()}, "e2e", round(d["e2e"]["value"],4))
EOP
done
