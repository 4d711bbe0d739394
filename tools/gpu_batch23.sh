#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-second-leg > gpurun_out/r02_bench_n${n}.json 2> gpurun_out/r02_bench_n${n}.err
tail -1 gpurun_out/r02_bench_n${n}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$n ms', d['ms_per_step'], 'value', '%.4g'%d['value'], 'K1', d['roofline']['avg_launch_ms_by_species'], 'check', d['check']['continuity_relative'], d['check']['gauss_drift_relative']); print(d['legs']['f32']['phases_ms'])"
done
