N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_mg${N}_c2.log 2> gpurun_out/bench_mg${N}_c2.err
$TR --master-port 29522 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --config 5 > gpurun_out/bench_mg${N}_c5.log 2> gpurun_out/bench_mg${N}_c5.err
$TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --config 4 > gpurun_out/bench_mg${N}_c4.log 2> gpurun_out/bench_mg${N}_c4.err
python - <<EOP
import json
for c in ("c2","c5","c4"):
    try:
        d=json.loads(open(f"gpurun_out/bench_mg${N}_{c}.log").read().strip().splitlines()[-1])
        print(c, d.get("n_gpus"), round(d["value"],4), d["unit"], round(d.get("ms_per_step",0),3), "e2e", round(d["e2e"]["value"],4))
    except Exception as e:
        print(c, "ERR", e)
EOP
tail -n 3 gpurun_out/bench_mg${N}_c2.err
