#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_distributed.py 2>&1 | tail -5 | tee gpurun_out/r02_dist8_test.log
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-second-leg > gpurun_out/r02_bench_n${n}_a.json 2> gpurun_out/r02_bench_n${n}_a.err
tail -1 gpurun_out/r02_bench_n${n}_a.json | cut -c1-400
done
