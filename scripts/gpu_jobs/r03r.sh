#!/bin/bash
mkdir -p gpurun_out
( time timeout 200 python bench.py --workload instances --instances 32767 --rays 67108864 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r03r_bench_instances_64Mrays.json 2> gpurun_out/r03r.err ) 2>&1 | grep real
tail -2 gpurun_out/r03r.err; python -c "
import json; d=json.loads(open('gpurun_out/r03r_bench_instances_64Mrays.json').read().strip().splitlines()[-1]); print(d['phase_ms'], d['rays'], d['animated_frame'])"
