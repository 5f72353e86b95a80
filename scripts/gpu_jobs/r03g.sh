#!/bin/bash
mkdir -p gpurun_out
run() { for I in 32767 100000; do timeout 600 python bench.py --workload instances --instances $I --steps 2 --warmup 1 --no-cpu-baseline --rays 65536 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 I=$I tlas', round(d['phase_ms']['tlas_build'],1))"; done; }
run base512x16
BVH_CUDA_TLAS_CL=8 run base512x8
BVH_CUDA_LIB=$PWD/voidin_b200/variants/libbvh_cuda_cl1024.so run cl1024x16
BVH_CUDA_TLAS_CL=8 BVH_CUDA_LIB=$PWD/voidin_b200/variants/libbvh_cuda_cl1024.so run cl1024x8
BVH_CUDA_LIB=$PWD/voidin_b200/variants/libbvh_cuda_cl1024.so timeout 900 python -m pytest tests -m gpu -x -q -k "tlas" 2>&1 | tail -2
