#!/bin/bash
# ncu --set full capture of the step kernel (default library, or $1) -> gpurun_out/prof_$2.ncu-rep
LIB=${1:-starfish_b200/libstarfish_gpu.so}; TAG=${2:-cur}
SFGPU_LIB_PATH=$PWD/$LIB timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_fast_step -s 6 -c 3 -f -o gpurun_out/prof_$TAG python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
