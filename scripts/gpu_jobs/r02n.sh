#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r02n_pytest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -3 gpurun_out/r02n_bench_n1.err
( time timeout 600 python bench.py --workload instances --steps 3 --warmup 1 > gpurun_out/r02n_bench_inst.json 2> gpurun_out/r02n_bench_inst.err ) 2>&1 | grep real
tail -3 gpurun_out/r02n_bench_inst.err; head -c 1500 gpurun_out/r02n_bench_inst.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02n_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['phase_ms'])
r=d['rays']; print(r['value'], r['ms'], r['e2e']['value'], r['incoherent']['value'])
print(d.get('config5'))
PY
