#!/bin/bash
mkdir -p gpurun_out
BVH_CUDA_LIB=voidin_b200/variants/libbvh_cuda_tctime.so timeout 300 python scripts/tc_timing.py dragon > gpurun_out/r02h_tc_timing.log 2>&1
echo "tc_timing rc=$?"; head -14 gpurun_out/r02h_tc_timing.log | cut -c1-250
timeout 420 python scripts/variants.py r02h > gpurun_out/r02h_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited" gpurun_out/r02h_variants.log | cut -c1-500
