set -x
python bench.py --profile --steps 2 --warmup 3 > gpurun_out/plain_r1b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_l_r1b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tile_kernel -s 6 -c 2 -o gpurun_out/prof_r1b python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_r1b.log 2>&1; tail -2 gpurun_out/ncu_r1b.log
