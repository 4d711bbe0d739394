#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --n 128 --steps 6 --warmup 1 --no-cpu-baseline --no-e2e --no-check --no-second-leg"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -s 16 -c 4 -o gpurun_out/r02_prof_v11e_n128 $B > gpurun_out/r02_prof_v11e_n128.log 2>&1
tail -2 gpurun_out/r02_prof_v11e_n128.log | cut -c1-200
