#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "two_gpus" 2>&1 | tail -5
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -5 gpurun_out/r02l_bench_n2.err | cut -c1-300; head -c 300 gpurun_out/r02l_bench_n2.json
