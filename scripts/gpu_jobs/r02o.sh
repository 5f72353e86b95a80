#!/bin/bash
# L2 peak, ncu captures (trace any-hit, grid tier on an 8 Mi soup), TLAS-top A/B on configs 3 and 5
mkdir -p gpurun_out
./scripts/micro/l2_bw 32 > gpurun_out/r02o_l2_peak.json; cat gpurun_out/r02o_l2_peak.json
./scripts/micro/l2_bw 64 | tee gpurun_out/r02o_l2_peak_64.json
for top in 0 1; do
  BVH_CUDA_TLAS_TOP=$top timeout 300 python bench.py --workload instances --steps 3 --warmup 1 > gpurun_out/r02o_inst_top$top.json 2> gpurun_out/r02o_inst_top$top.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02o_inst_top$top.json').read().strip().splitlines()[-1]); print('inst top=$top', d['phase_ms'], d['rays']['value'])"
  BVH_CUDA_TLAS_TOP=$top timeout 600 python bench.py --workload scene1024 --steps 2 --warmup 1 --c5-rays 134217728 > gpurun_out/r02o_c5_top$top.json 2> gpurun_out/r02o_c5_top$top.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02o_c5_top$top.json').read().strip().splitlines()[-1]); print('c5 top=$top', d.get('rays'), d.get('value'))"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_any -c 1 -o gpurun_out/r02o_k_trace_any_16M python bench.py --steps 1 --warmup 0 --no-config5 --no-cpu-baseline > gpurun_out/r02o_ncu_trace.log 2>&1
tail -2 gpurun_out/r02o_ncu_trace.log
timeout 1200 ncu --set full --clock-control none -k regex:k_t1_coop -c 1 -o gpurun_out/r02o_k_t1_coop_soup8M python bench.py --workload soup --meshes 8388608 --steps 1 --warmup 0 > gpurun_out/r02o_ncu_soup.log 2>&1
tail -2 gpurun_out/r02o_ncu_soup.log
timeout 600 python bench.py --workload soup --meshes 8388608 --steps 3 --warmup 1 > gpurun_out/r02o_soup8M.json 2> gpurun_out/r02o_soup8M.err; head -c 1200 gpurun_out/r02o_soup8M.json
ls -la gpurun_out/*.ncu-rep
