for sg in 0.01 2 8; do
python bench.py --steps 30 --warmup 3 --no-cpu --sigma $sg > gpurun_out/bench_sig_$sg.log 2> gpurun_out/bench_sig_$sg.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_sig_$sg.log").read().strip().splitlines()[-1])
print("sigma $sg", round(d["ms_per_step"],4), {k:round(v["ms"],4) for k,v in d["kernels"].items()})
EOP
done
