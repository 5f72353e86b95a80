#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "tlas or smoke or animated or instances" > gpurun_out/r03h_pytest.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/r03h_pytest.log
for I in 14001 32767 100000; do timeout 600 python bench.py --workload instances --instances $I --steps 2 --warmup 1 --no-cpu-baseline --rays 65536 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('I=$I tlas', round(d['phase_ms']['tlas_build'],1))"; done
