#!/bin/bash
# final validation of the round on the final code
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03j_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r03j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03j_smoke.log 2>&1; tail -1 gpurun_out/r03j_smoke.log
( time timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r03j_bench_n1.json 2> gpurun_out/r03j_bench_n1.err ) 2>&1 | grep real
tail -2 gpurun_out/r03j_bench_n1.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03j_bench_reference.json 2> gpurun_out/r03j_bench_reference.err ) 2>&1 | grep real
timeout 300 python bench.py --workload instances --steps 3 --warmup 1 > gpurun_out/r03j_bench_instances.json 2>/dev/null
timeout 300 python bench.py --workload instances --instances 4096 --steps 3 --warmup 1 > gpurun_out/r03j_bench_instances4096.json 2>/dev/null
python - <<'PY'
import json
def last(fn): return json.loads(open(fn).read().strip().splitlines()[-1])
d=last('gpurun_out/r03j_bench_n1.json'); r=d['rays']
print('N1', round(d['value'],1), d['phase_ms'], 'e2e', round(d['e2e']['value'],1), 'rays', round(r['value'],1), round(r['e2e']['value'],1), round(r['incoherent']['value'],1))
print('REF', last('gpurun_out/r03j_bench_reference.json')['value'])
for w in ('instances','instances4096'):
    d=last(f'gpurun_out/r03j_bench_{w}.json'); print(w, d['phase_ms'], d['animated_frame']['ms'], d['cpu_baseline'])
PY
