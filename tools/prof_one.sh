#!/bin/bash
# usage: tools/prof_one.sh <tag> <kernel regex> [skip]   (run under gpurun; writes gpurun_out/prof_<tag>.ncu-rep)
ncu --set full --import-source on --clock-control none -k regex:"$2" -s ${3:-3} -c 1 -o gpurun_out/prof_$1 python bench.py --profile --steps 2 --warmup 3 > gpurun_out/ncu_$1.log 2>&1; tail -2 gpurun_out/ncu_$1.log
