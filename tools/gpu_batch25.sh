#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x -m gpu tests/test_gpu_distributed.py -k "2x2x2 and f64" 2>&1 | tail -3 | tee gpurun_out/r02_dist8_test_b.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-second-leg > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -1 gpurun_out/r02_bench_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=8 ms', d['ms_per_step'], 'value', '%.4g'%d['value'], 'K1', d['roofline']['avg_launch_ms_by_species'], 'check', d['check']['continuity_relative'], d['check']['gauss_drift_relative']); print(d['legs']['f32']['phases_ms'])"
