#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload scene1024 --meshes 64 --mesh-res 61 --c5-rays 4194304 --steps 2 --warmup 1 > gpurun_out/r03e_scene64.json 2> gpurun_out/r03e_scene64.err; tail -3 gpurun_out/r03e_scene64.err; head -c 900 gpurun_out/r03e_scene64.json; echo
timeout 900 python bench.py --workload scene1024 --steps 2 --warmup 1 > gpurun_out/r03e_scene1024.json 2> gpurun_out/r03e_scene1024.err; tail -3 gpurun_out/r03e_scene1024.err; python -c "
import json; d=json.loads(open('gpurun_out/r03e_scene1024.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline'], d['cpu_baseline'])"
out=gpurun_out/r03e_sanitizer.log; : > $out
run() { echo "== $*" >> $out; ( timeout 900 "$@" 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|rror" | head -20 ) >> $out; }
for m in 1 2; do
  export BVH_CUDA_T1_PULL=$m BVH_CUDA_TLAS_TOP=1
  echo "-- with BVH_CUDA_T1_PULL=$m BVH_CUDA_TLAS_TOP=1" >> $out
  run compute-sanitizer --tool memcheck python scripts/sanitize_small.py 130000
  run compute-sanitizer --tool racecheck python scripts/sanitize_small.py 60000
done
cat $out
