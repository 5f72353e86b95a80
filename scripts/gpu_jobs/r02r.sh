#!/bin/bash
mkdir -p gpurun_out
BVH_CUDA_LIB=$PWD/voidin_b200/variants/libbvh_cuda_r02p_timing.so timeout 300 python scripts/t1_timing.py > gpurun_out/r02r_t1_timing.log 2>&1; tail -26 gpurun_out/r02r_t1_timing.log | cut -c1-400
