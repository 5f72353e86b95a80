#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/r02u_pytest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02u_bench_n1.json 2> gpurun_out/r02u_bench_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/r02u_bench_n1.err
for w in bunny instances; do
  ( time timeout 600 python bench.py --workload $w --steps 5 --warmup 1 > gpurun_out/r02u_bench_$w.json 2> gpurun_out/r02u_bench_$w.err ) 2>&1 | grep real
  tail -3 gpurun_out/r02u_bench_$w.err
done
( time timeout 900 python bench.py --workload soup --meshes 8388608 --steps 3 --warmup 1 > gpurun_out/r02u_bench_soup8M.json 2> gpurun_out/r02u_bench_soup8M.err ) 2>&1 | grep real
tail -3 gpurun_out/r02u_bench_soup8M.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['phase_ms'], d['phase_ms_dragon']['ms_grid'])
r=d['rays']; print(r['value'], r['ms'], r['e2e']['value'], r['incoherent']['value'], r['roofline']['l2_model'], r['roofline']['issue'] is not None)
for w in ('bunny','instances','soup8M'):
    try:
        d=json.loads(open(f'gpurun_out/r02u_bench_{w}.json').read().strip().splitlines()[-1])
        print(w, d['value'], d.get('phase_ms'), d.get('cpu_baseline'), d.get('rays',{}).get('cpu_baseline'), {k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic_over_algorithmic','pair_evaluations_per_s')})
    except Exception as e: print(w, 'ERR', e)
PY
