#!/bin/bash
# cluster tier v2 (cp.async staging, batched look-ups) + in-place suffix-only shuffles in the block / warp-node tiers
mkdir -p gpurun_out
timeout 600 python scripts/variants.py base:BVH_CUDA_NO_CLUSTER=1 base r02c > gpurun_out/r02c_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited" gpurun_out/r02c_variants.log | cut -c1-1200
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02c_pytest.log
tail -6 gpurun_out/r02c_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02c_smoke.log 2>&1; tail -2 gpurun_out/r02c_smoke.log
