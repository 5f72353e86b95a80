#!/bin/bash
mkdir -p gpurun_out
PARITY=1 SIZES=1 timeout 900 python scripts/any_check.py 24 > gpurun_out/r02v_any_steal1.log 2>&1; grep -E "ANY-HIT|mismatch|rays:|half|dragon any" gpurun_out/r02v_any_steal1.log
BVH_CUDA_TRACE_STEAL=0 PARITY=0 SIZES=1 timeout 900 python scripts/any_check.py 24 > gpurun_out/r02v_any_steal0.log 2>&1; grep -E "rays:|half|dragon any" gpurun_out/r02v_any_steal0.log
RAYS=incoherent PARITY=0 timeout 900 python scripts/any_check.py 24 > gpurun_out/r02v_any_inc_steal1.log 2>&1; grep -E "dragon any" gpurun_out/r02v_any_inc_steal1.log
BVH_CUDA_TRACE_STEAL=0 RAYS=incoherent PARITY=0 timeout 900 python scripts/any_check.py 24 > gpurun_out/r02v_any_inc_steal0.log 2>&1; grep -E "dragon any" gpurun_out/r02v_any_inc_steal0.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trace or any or smoke" > gpurun_out/r02v_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r02v_pytest.log
