#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-second-leg "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['check']; print('$tag', 'cont_rel', c['continuity_relative'], 'gauss_rel', c['gauss_drift_relative'], 'ms', d['ms_per_step'], 'ovf', d['overflow'])"; }
run "n128 f32" --dtype f32 --cells-per-axis 128 | tee gpurun_out/r02_dbg_check5.log
run "n256 f32" --dtype f32 | tee -a gpurun_out/r02_dbg_check5.log
