#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4 2; do
  extra=""; [ $n != 8 ] && extra="--no-config5"
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 $extra > gpurun_out/r03l_bench_n$n.json 2> gpurun_out/r03l_bench_n$n.err ) 2>&1 | grep real
  grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r03l_bench_n$n.err | tail -3 | cut -c1-300
done
( time timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-config5 > gpurun_out/r03l_bench_n1.json 2> gpurun_out/r03l_bench_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r03l_bench_n{n}.json').read().strip().splitlines()[-1])
        r=d['rays']
        print(n, 'build', round(d['value'],1), d['phase_ms'], 'e2e', round(d['e2e']['value'],1), '| rays strong', round(r['value'],1), 'ms', round(r['ms'],3), 'weak', (r['weak'] or {}).get('value'), 'e2e', round(r['e2e']['value'],1), 'inc', round(r['incoherent']['value'],1))
        print('   multi_gpu', d.get('multi_gpu'))
        if d.get('config5'): print('   c5', json.dumps(d['config5'])[:1200])
    except Exception as e: print(n,'ERR',e)
PY
