#!/bin/bash
# first contact of the cluster tier with the hardware: micro-benchmark, A/B against the grid-only build, parity
mkdir -p gpurun_out
timeout 120 scripts/micro/dsmem_bench > gpurun_out/r02b_dsmem.log 2>&1
echo "dsmem rc=$?"
timeout 420 python scripts/variants.py base base:BVH_CUDA_NO_CLUSTER=1 > gpurun_out/r02b_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited" gpurun_out/r02b_variants.log | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02b_fullsize.log
tail -5 gpurun_out/r02b_fullsize.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
