#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_all.csv python bench.py --steps 11 --warmup 1 --no-cpu-baseline --no-e2e --no-check --no-second-leg > gpurun_out/r02_launches_all.log 2>&1
