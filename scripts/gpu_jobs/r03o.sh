#!/bin/bash
mkdir -p gpurun_out
( time BVH_CUDA_INSTANCE_CULL=1 timeout 900 python -m pytest tests -m gpu -x -q -k "trace or culling or animated or any_hit or smoke or above_32767" > gpurun_out/r03o_pytest.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/r03o_pytest.log
for c in 0 1; do
BVH_CUDA_INSTANCE_CULL=$c timeout 300 python bench.py --workload instances --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cull=$c', d['phase_ms'], d['rays']['value'], d['animated_frame']['ms'], d['rays']['cpu_baseline'])"
done
