#!/bin/bash
# last validation of the round + a fresh ncu capture of the any-hit kernel (with tail stealing)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03m_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r03m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03m_smoke.log 2>&1; tail -1 gpurun_out/r03m_smoke.log
( time timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r03m_bench_n1.json 2> gpurun_out/r03m_bench_n1.err ) 2>&1 | grep real
tail -2 gpurun_out/r03m_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_any -c 1 -o gpurun_out/r03m_k_trace_any_16M python bench.py --steps 1 --warmup 0 --no-config5 --no-cpu-baseline > gpurun_out/r03m_ncu_trace.log 2>&1; tail -1 gpurun_out/r03m_ncu_trace.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03m_bench_n1.json').read().strip().splitlines()[-1]); r=d['rays']
print('N1', round(d['value'],1), d['phase_ms'], 'e2e', round(d['e2e']['value'],1), 'rays', round(r['value'],1), round(r['e2e']['value'],1), 'issue' in d['roofline'])
PY
