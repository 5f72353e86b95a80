#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/variants.py r02p r02p:BVH_CUDA_T1_PULL=0 > gpurun_out/r02q_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited|rror" gpurun_out/r02q_variants.log | cut -c1-600
BVH_CUDA_LIB=$PWD/voidin_b200/variants/libbvh_cuda_r02p_timing.so timeout 300 python scripts/t1_timing.py > gpurun_out/r02q_t1_timing.log 2>&1; tail -22 gpurun_out/r02q_t1_timing.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blas" > gpurun_out/r02q_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r02q_pytest.log
