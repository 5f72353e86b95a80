#!/bin/bash
# last validation of the round on the final code (strict timeouts)
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r03q_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r03q_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03q_smoke.log 2>&1; tail -1 gpurun_out/r03q_smoke.log
( time timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r03q_bench_n1.json 2> gpurun_out/r03q_bench_n1.err ) 2>&1 | grep real
timeout 120 python bench.py --workload instances --steps 3 --warmup 1 > gpurun_out/r03q_bench_instances.json 2>/dev/null
python - <<'PY'
import json
def last(fn): return json.loads(open(fn).read().strip().splitlines()[-1])
d=last('gpurun_out/r03q_bench_n1.json'); r=d['rays']
print('N1', round(d['value'],1), d['phase_ms'], 'e2e', round(d['e2e']['value'],1), 'rays', round(r['value'],1), round(r['e2e']['value'],1))
d=last('gpurun_out/r03q_bench_instances.json'); print('inst', d['phase_ms'], d['rays']['value'], d['animated_frame']['ms'], d['rays']['cpu_baseline'])
PY
