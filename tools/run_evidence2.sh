#!/bin/bash
# Final evidence run of round 2 (under gpurun): tests, launch list, full ncu capture of the two default kernels, phase timeline,
# bench lines of every single-GPU config (recorded in profiles/r2_bench_lines.jsonl).
set -x
T=${1:-r2b}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_ncu_l.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"bwd_cl|fwd_tex" -s 6 -c 2 -o gpurun_out/prof_${T}_head python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log
FWB_LIB=$PWD/deep_video_interpolation_extrapolation_b200/csrc/libflowwarp_b200_ab.so timeout 300 python tools/cl_prof.py > gpurun_out/${T}_cl_timeline.txt 2>&1
grep -v '"label": "mg2_' profiles/r2_bench_lines.jsonl > /dev/null  # (the 2-GPU lines are re-recorded by tools/run_mg_r2.sh)
rm -f profiles/r2_bench_lines.jsonl
timeout 900 python bench.py --steps 50 --warmup 5 --aux --record config2 > gpurun_out/${T}_bench2.log 2> gpurun_out/${T}_bench2.err
timeout 300 python bench.py --config 1 --steps 50 --warmup 5 --record config1 > gpurun_out/${T}_bench1.log 2> gpurun_out/${T}_bench1.err
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu --record config3 > gpurun_out/${T}_bench3.log 2> gpurun_out/${T}_bench3.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu --record config4 > gpurun_out/${T}_bench4.log 2> gpurun_out/${T}_bench4.err
timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu --deterministic --record config2_deterministic > gpurun_out/${T}_bench2d.log 2> gpurun_out/${T}_bench2d.err
cp profiles/r2_bench_lines.jsonl gpurun_out/${T}_bench_lines.jsonl
python tools/kt.py gpurun_out/${T}_bench2.log gpurun_out/${T}_bench1.log gpurun_out/${T}_bench3.log gpurun_out/${T}_bench4.log gpurun_out/${T}_bench2d.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.log 2> gpurun_out/${T}_ref.err; tail -c 600 gpurun_out/${T}_ref.log
