#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/variants.py r02h:BVH_CUDA_NO_CLUSTER=1 r02h:BVH_CUDA_TC=global r02h > gpurun_out/r02i_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited" gpurun_out/r02i_variants.log | cut -c1-420
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02i_pytest.log
tail -6 gpurun_out/r02i_pytest.log
