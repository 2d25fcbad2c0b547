#!/bin/bash
# Evidence run for profiles/ (under gpurun): launch list, full ncu capture of the two default kernels, bench lines of every config.
set -x
T=${1:-r2}
python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_ncu_l.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"bwd_cl|fwd_tex" -s 6 -c 2 -o gpurun_out/prof_${T}_head python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
rm -f profiles/r2_bench_lines.jsonl
python bench.py --steps 50 --warmup 5 --aux --record config2 > gpurun_out/${T}_bench2.log 2> gpurun_out/${T}_bench2.err
python bench.py --config 1 --steps 50 --warmup 5 --record config1 > gpurun_out/${T}_bench1.log 2> gpurun_out/${T}_bench1.err
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --record config3 > gpurun_out/${T}_bench3.log 2> gpurun_out/${T}_bench3.err
python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --record config4 > gpurun_out/${T}_bench4.log 2> gpurun_out/${T}_bench4.err
python bench.py --config 2 --steps 20 --warmup 5 --no-cpu --deterministic --record config2_deterministic > gpurun_out/${T}_bench2d.log 2> gpurun_out/${T}_bench2d.err
cp profiles/r2_bench_lines.jsonl gpurun_out/${T}_bench_lines.jsonl
python tools/kt.py gpurun_out/${T}_bench2.log gpurun_out/${T}_bench1.log gpurun_out/${T}_bench3.log gpurun_out/${T}_bench4.log gpurun_out/${T}_bench2d.log
