#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/micro/dsmem_bench > gpurun_out/r02e_dsmem.log 2>&1; echo "dsmem rc=$?"
timeout 420 python scripts/variants.py r02e r02e:BVH_CUDA_TC=global > gpurun_out/r02e_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited" gpurun_out/r02e_variants.log | cut -c1-700
BVH_CUDA_LIB=voidin_b200/variants/libbvh_cuda_tctime.so timeout 300 python scripts/tc_timing.py dragon > gpurun_out/r02e_tc_timing.log 2>&1
echo "tc_timing rc=$?"; head -12 gpurun_out/r02e_tc_timing.log | cut -c1-200
