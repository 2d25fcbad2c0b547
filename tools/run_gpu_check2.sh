python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_u1.log; tail -2 gpurun_out/pytest_u1.log
for ppt in 1 2; do
FWB_TILE_BWD_PPT=$ppt python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_u1_$ppt.log 2> gpurun_out/bench_u1_$ppt.err
python - <<EOP
import json
d=json.loads(open("gpurun_out/bench_u1_$ppt.log").read().strip().splitlines()[-1])
print("ppt $ppt", d["ms_per_step"], {k:round(v["ms"],4) for k,v in d["kernels"].items()})
EOP
done
ncu --set full --import-source on --clock-control none -k regex:bwd_tile_kernel -s 1 -c 1 -o gpurun_out/prof_u1 python bench.py --profile --steps 1 --warmup 3 > gpurun_out/ncu_u1.log 2>&1; tail -1 gpurun_out/ncu_u1.log
