#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/variants.py r03d:BVH_CUDA_T1_PULL=0 r03d:BVH_CUDA_T1_PULL=1 r03d:BVH_CUDA_T1_PULL=2 > gpurun_out/r03d_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited|rror" gpurun_out/r03d_variants.log | cut -c1-330
for p in 0 1; do BVH_CUDA_T1_PULL=$p BVH_CUDA_LIB=$PWD/voidin_b200/variants/libbvh_cuda_r03d_timing.so timeout 300 python scripts/t1_timing.py > gpurun_out/r03d_t1_timing_pull$p.log 2>&1; grep "per level" gpurun_out/r03d_t1_timing_pull$p.log | tail -1; done
for p in 0 1; do BVH_CUDA_T1_PULL=$p timeout 300 python bench.py --workload bunny --steps 10 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bunny pull=$p', d['phase_ms'])"; done
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blas" > gpurun_out/r03d_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r03d_pytest.log
