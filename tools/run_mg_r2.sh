TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --record mg2_config2 > gpurun_out/r2_mg2_c2.log 2> gpurun_out/r2_mg2_c2.err
$TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --config 4 --no-cpu --record mg2_config4 > gpurun_out/r2_mg2_c4.log 2> gpurun_out/r2_mg2_c4.err
$TR --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --config 5 --no-cpu --record mg2_config5 > gpurun_out/r2_mg2_c5.log 2> gpurun_out/r2_mg2_c5.err
$TR --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/r2_mg2_ref.log 2> gpurun_out/r2_mg2_ref.err
cp profiles/r2_bench_lines.jsonl gpurun_out/r2_mg2_lines.jsonl
python - <<EOP
import json
for c in ("c2","c4","c5","ref"):
    try:
        d=json.loads(open(f"gpurun_out/r2_mg2_{c}.log").read().strip().splitlines()[-1])
        print(c, d.get("n_gpus"), round(d["value"],4), d["unit"], round(d.get("ms_per_step",0),3), "e2e", d["e2e"]["value"])
    except Exception as e:
        print(c, "ERR", e)
EOP
tail -n 3 gpurun_out/r2_mg2_c2.err gpurun_out/r2_mg2_c5.err gpurun_out/r2_mg2_ref.err
