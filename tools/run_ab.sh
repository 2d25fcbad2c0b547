#!/bin/bash
# A/B under gpurun: the config-2 bench line with the default library and with csrc/libflowwarp_b200_ab.so.  usage: tools/run_ab.sh <tag> [bench args]
T=$1; shift
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu "$@" > gpurun_out/bench_${T}_main.log 2> gpurun_out/bench_${T}_main.err
FWB_LIB=$PWD/deep_video_interpolation_extrapolation_b200/csrc/libflowwarp_b200_ab.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu "$@" > gpurun_out/bench_${T}_ab.log 2> gpurun_out/bench_${T}_ab.err
python tools/kt.py gpurun_out/bench_${T}_main.log gpurun_out/bench_${T}_ab.log
