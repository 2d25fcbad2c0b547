# 8 GPUs of one box (gpurun --gpus 8): config 2 and config 5 lines of round 2, recorded as mg8_* in profiles/r2_bench_lines.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu --record mg8_config2 > gpurun_out/r2_mg8_c2.log 2> gpurun_out/r2_mg8_c2.err
$TR --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 10 --config 5 --no-cpu --record mg8_config5 > gpurun_out/r2_mg8_c5.log 2> gpurun_out/r2_mg8_c5.err
cp profiles/r2_bench_lines.jsonl gpurun_out/r2_mg8_lines.jsonl
python - <<EOP
import json
for c in ("c2","c5"):
    try:
        d=json.loads(open(f"gpurun_out/r2_mg8_{c}.log").read().strip().splitlines()[-1])
        print(c, d.get("n_gpus"), round(d["value"],4), d["unit"], round(d.get("ms_per_step",0),4), "e2e", d["e2e"]["value"])
    except Exception as e:
        print(c, "ERR", e)
EOP
tail -n 3 gpurun_out/r2_mg8_c2.err gpurun_out/r2_mg8_c5.err
