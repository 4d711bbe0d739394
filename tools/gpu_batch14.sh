#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dbg_check.py 256 25 f32 2>&1 | tail -8 | tee gpurun_out/r02_dbg_check2.log
for extra in "--no-e2e" "--e2e-steps 1"; do
timeout 400 python bench.py --dtype f64 --steps 5 --warmup 3 --no-cpu-baseline --no-second-leg $extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['check']; print('bench f64 $extra', 'cont_res', c['continuity_residual_max'], 'cont_rel', c['continuity_relative'], 'gauss_rel', c['gauss_drift_relative'])" | tee -a gpurun_out/r02_dbg_check2.log
done
