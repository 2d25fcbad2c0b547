for kb in 52 48 40; do
FWB_TILE_FWD_KB=$kb python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_v11_$kb.log 2> gpurun_out/bench_v11_$kb.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_v11_$kb.log").read().strip().splitlines()[-1])
print("fwd 4cta kb $kb", round(d["ms_per_step"],4), {k:round(v["ms"],4) for k,v in d["kernels"].items()})
EOP
done
