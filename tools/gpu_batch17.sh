#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_distributed.py -k "2-1-1 or 1-1-2 or mesh0 or mesh1 or tile_kernel" 2>&1 | tail -4 | tee gpurun_out/r02_dist2_test.log
for mrg in 1 0; do
PIC_J_MERGED=$mrg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-second-leg 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['check']; print('merged=$mrg', 'cont_rel', c['continuity_relative'], 'gauss_rel', c['gauss_drift_relative'], 'ms', d['ms_per_step'], 'K1', d['roofline']['avg_launch_ms_by_species'])" | tee -a gpurun_out/r02_dist2_test.log
done
