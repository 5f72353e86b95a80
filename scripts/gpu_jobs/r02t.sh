#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/variants.py r02t_g1 r02t_g8 r02t r02t_g37 r02t:BVH_CUDA_T1_PULL=1 > gpurun_out/r02t_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited|rror" gpurun_out/r02t_variants.log | cut -c1-330
