#!/bin/bash
# final validation of the round: full GPU suite, all bench lines, launch list + ncu captures, fuzz
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r03b_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r03b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03b_smoke.log 2>&1; tail -1 gpurun_out/r03b_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r03b_bench_n1.json 2> gpurun_out/r03b_bench_n1.err ) 2>&1 | grep real
tail -2 gpurun_out/r03b_bench_n1.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03b_bench_reference.json 2> gpurun_out/r03b_bench_reference.err ) 2>&1 | grep real
for w in bunny instances; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 1 > gpurun_out/r03b_bench_$w.json 2> gpurun_out/r03b_bench_$w.err; tail -2 gpurun_out/r03b_bench_$w.err
done
timeout 900 python bench.py --workload instances --instances 100000 --steps 2 --warmup 1 > gpurun_out/r03b_bench_instances100k.json 2> gpurun_out/r03b_bench_instances100k.err; tail -2 gpurun_out/r03b_bench_instances100k.err
timeout 900 python bench.py --workload soup --steps 2 --warmup 1 > gpurun_out/r03b_bench_soup64M.json 2> gpurun_out/r03b_bench_soup64M.err; tail -2 gpurun_out/r03b_bench_soup64M.err
timeout 900 python bench.py --workload scene1024 --steps 2 --warmup 1 > gpurun_out/r03b_bench_scene1024.json 2> gpurun_out/r03b_bench_scene1024.err; tail -2 gpurun_out/r03b_bench_scene1024.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-config5 --no-cpu-baseline > gpurun_out/r03b_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_t1_coop -c 1 -o gpurun_out/r03b_k_t1_coop python scripts/one_build.py 1 0 > gpurun_out/r03b_ncu_t1.log 2>&1; tail -1 gpurun_out/r03b_ncu_t1.log
timeout 900 ncu --set full --clock-control none -k regex:k_t2 -c 1 -o gpurun_out/r03b_k_t2_big python scripts/one_build.py 1 0 > gpurun_out/r03b_ncu_t2.log 2>&1; tail -1 gpurun_out/r03b_ncu_t2.log
timeout 400 python scripts/fuzz_gpu.py 180 7 > gpurun_out/r03b_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r03b_fuzz_gpu.log
python - <<'PY'
import json
def last(fn):
    return json.loads(open(fn).read().strip().splitlines()[-1])
d=last('gpurun_out/r03b_bench_n1.json'); r=d['rays']
print('N1', round(d['value'],1), d['phase_ms'], 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms'],2), 'frac', d['roofline']['frac'], d['roofline']['whole_build']['frac'])
print('   rays', round(r['value'],1), r['ms'], 'e2e', round(r['e2e']['value'],1), 'inc', round(r['incoherent']['value'],1))
print('   c5', json.dumps(d.get('config5'))[:900])
print('REF', open('gpurun_out/r03b_bench_reference.json').read()[:1200])
for w in ('bunny','instances','instances100k','soup64M','scene1024'):
    try:
        d=last(f'gpurun_out/r03b_bench_{w}.json'); print(w, round(d['value'],2), d.get('unit'), d.get('phase_ms'), json.dumps(d.get('roofline'))[:300], d.get('cpu_baseline'))
    except Exception as e: print(w,'ERR',e)
PY
ls -la gpurun_out/*.ncu-rep | tail -3
