#!/bin/bash
mkdir -p gpurun_out
BVH_CUDA_LIB=voidin_b200/variants/libbvh_cuda_tctime.so timeout 300 python scripts/tc_timing.py dragon > gpurun_out/r02d_tc_timing.log 2>&1
echo "tc_timing rc=$?"; head -50 gpurun_out/r02d_tc_timing.log | cut -c1-200
