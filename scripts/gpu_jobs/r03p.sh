#!/bin/bash
mkdir -p gpurun_out
( time BVH_CUDA_INSTANCE_CULL=1 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "culling or two_level or animated or any_hit_order or staged" > gpurun_out/r03p_pytest.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/r03p_pytest.log
