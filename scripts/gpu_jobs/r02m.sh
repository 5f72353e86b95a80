#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-config5 > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err ) 2>&1 | grep real
echo "bench rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02m_bench_n2.err | tail -5 | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['phase_ms'], d['multi_gpu'])
r=d['rays']; print(r['value'], r['ms'], r['weak']['value'], r['e2e']['value'], r['incoherent']['value'])
PY
