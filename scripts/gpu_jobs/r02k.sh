#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -3 gpurun_out/r02k_bench_n1.err; head -c 600 gpurun_out/r02k_bench_n1.json
