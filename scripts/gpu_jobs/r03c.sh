#!/bin/bash
# compute-sanitizer over the default build and the opt-in paths added in round 2
mkdir -p gpurun_out
out=gpurun_out/r03c_sanitizer.log; : > $out
run() { echo "== $*" >> $out; ( timeout 900 "$@" 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|rror" | head -20 ) >> $out; }
run compute-sanitizer --tool memcheck python scripts/sanitize_small.py 130000
run compute-sanitizer --tool racecheck python scripts/sanitize_small.py
run compute-sanitizer --tool synccheck python scripts/sanitize_small.py
export BVH_CUDA_T1_PULL=1 BVH_CUDA_TLAS_TOP=1
echo "-- with BVH_CUDA_T1_PULL=1 BVH_CUDA_TLAS_TOP=1" >> $out
run compute-sanitizer --tool memcheck python scripts/sanitize_small.py 130000
run compute-sanitizer --tool racecheck python scripts/sanitize_small.py 60000
run compute-sanitizer --tool synccheck python scripts/sanitize_small.py 60000
cat $out
