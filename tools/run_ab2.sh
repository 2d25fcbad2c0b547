#!/bin/bash
# A/B with parity tests on the A/B library first.  usage: tools/run_ab2.sh <tag>
T=$1
FWB_LIB=$PWD/deep_video_interpolation_extrapolation_b200/csrc/libflowwarp_b200_ab.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
tools/run_ab.sh $T
