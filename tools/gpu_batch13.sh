#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dt in f64 f32; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --cells-per-axis 64 --dtype $dt --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-second-leg > gpurun_out/r02_check_dist_$dt.json 2> gpurun_out/r02_check_dist_$dt.err
tail -3 gpurun_out/r02_check_dist_$dt.err | cut -c1-300
tail -1 gpurun_out/r02_check_dist_$dt.json | cut -c1-200
done
